"""One eager training step of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`):
every kernel of the step appears once, in issue order.   python scripts/one_step.py [workload] [precision] [n_steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import ops, synthetic  # noqa: E402
from get_b200.ddp import trainable_named_parameters  # noqa: E402
from get_b200.model import Graph_basedSemantiStructure  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "snopes"
ops.set_precision(sys.argv[2] if len(sys.argv) > 2 else "fp32")
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
w = synthetic.get_workload(wl)
torch.manual_seed(123756)
model = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev).train()
params = [p for _, p in trainable_named_parameters(model)]
opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-3, fused=True)
batches = [synthetic.batch_to_torch(synthetic.make_batch(w, seed=123756 + 1000 * i), device=dev) for i in range(2)]


def step(i):
    q, d, l, kw = batches[i % 2]
    opt.zero_grad(set_to_none=True)
    loss = ops.cross_entropy(model(q, d, **kw), l)
    loss.backward()
    opt.step()
    return loss


for i in range(3):
    step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(nsteps):
    step(3 + i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("pairs", batches[1][3]["doc_content_without_padding_evidences"].shape[0])
