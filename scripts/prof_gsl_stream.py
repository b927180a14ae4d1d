"""Three train-mode launches of the fused GSL kernel at streaming size (7 680 Snopes graphs, BASELINE.json configs[4] regime)
for `ncu -k regex:gather_row_kernel -s 2 -c 1`; inputs alternate between two sets (2 x 0.9 GB > L2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import ops, synthetic  # noqa: E402

DEV = "cuda"
w = synthetic.get_workload("snopes")
N, H, G = 100, 300, int(sys.argv[1]) if len(sys.argv) > 1 else 7680
rng = np.random.default_rng(7)
base = []
for g in range(48):
    pool = rng.integers(2, w.vocab, size=140)
    toks = pool[rng.integers(0, 140, size=N)]
    base.append(synthetic.word_graph(toks, N, 3)[1].astype(np.float32))
adj = ops.NeighborLists(torch.from_numpy(np.stack(base)).to(DEV).repeat((G + 47) // 48, 1, 1)[:G].contiguous())
wp, gate = torch.randn(H, device=DEV) * 0.1, torch.randn(12, device=DEV)
feats = [torch.randn(G, N, H, device=DEV) for _ in range(2)]
sps = [ops.rowdot(f.view(G * N, H), wp, 0.2, 1) for f in feats]
for i in range(3):
    ops.gsl_fused(adj, feats[i % 2], wp, gate, 60, drop_p=0.2, seed_scorer=1, seed_layer2=2, planes_n=2, sp_parts=sps[i % 2])
torch.cuda.synchronize()
print("done", G)
