"""Single-GPU repro of the per-rank shards of a 2-rank run through the captured training step (reducer attached, FlatAdam)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from get_b200 import ops, synthetic  # noqa: E402
from get_b200.ddp import shard_claims  # noqa: E402
from get_b200.keywords import KeyWordSettings as K  # noqa: E402
from get_b200.step_graph import pad_batch, slice_batch  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
graph = (sys.argv[2] != "eager") if len(sys.argv) > 2 else True
dev = torch.device("cuda", 0)
w = synthetic.get_workload("snopes")
model, reducer, opt, stepper = bench.build_trainer(w, dev, "fp32", use_graph=True)
glob = bench.make_batches(w, 16, 123756, n_claims=w.batch_claims * world)
sync = os.environ.get("REPRO_SYNC", "0") == "1"
order = [(0, bi) for bi in range(5)] + [(0, bi) for bi in range(16)] + [(0, bi % 16) for bi in range(30)]
for rank, bi in order:
    for g in [glob[bi]]:
        lo, hi = shard_claims(g[K.EvidenceCountPerQuery], world)[rank]
        b = pad_batch(slice_batch(g, lo, hi), 16)
        q, d, l, kw = synthetic.batch_to_torch(b, device=dev)
        n = b.get("n_real_claims", b["query"].shape[0])
        try:
            if graph:
                loss = stepper.step(q, d, l, kw, n, global_claims=g["query"].shape[0])
            else:
                reducer.zero()
                loss = ops.cross_entropy(model(q, d, **kw)[:n], l[:n])
                loss.backward()
                reducer.reduce()
                opt.step()
            if sync:
                torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("FAIL batch %d rank %d claims %d (real %d) pairs %d: %s" % (bi, rank, q.shape[0], n, d.shape[0], str(e)[:300]), flush=True)
            raise
        print("ok batch %d rank %d claims %d (real %d) pairs %d" % (bi, rank, q.shape[0], n, d.shape[0]), flush=True)
torch.cuda.synchronize()
print("all done", flush=True)
