"""Timeline / timing of the fused GSL graph kernel (build with GETB_EXTRA_NVCC_FLAGS=-DGETB_GRAPH_TIMELINE for the timeline)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import ops, synthetic  # noqa: E402

DEV = "cuda"
w = synthetic.get_workload("snopes")
N, H = 100, 300
rng = np.random.default_rng(7)
base = []
for g in range(48):
    pool = rng.integers(2, w.vocab, size=140)
    toks = pool[rng.integers(0, 140, size=N)]
    base.append(synthetic.word_graph(toks, N, 3)[1].astype(np.float32))
a1 = torch.from_numpy(np.stack(base)).to(DEV)
wp, gate = torch.randn(H, device=DEV) * 0.1, torch.randn(12, device=DEV)
for graphs in (220, 7680):
    adj = a1.repeat((graphs + 47) // 48, 1, 1)[:graphs].contiguous()
    feats = [torch.randn(graphs, N, H, device=DEV) for _ in range(2)]
    for p in (0.0, 0.2):
        sps = [ops.rowdot(f.view(graphs * N, H), wp, p, 1) for f in feats]
        for planes in (2, 0):
            for i in range(3):
                ops.gsl_fused(adj, feats[i % 2], wp, gate, 60, drop_p=p, seed_scorer=1, seed_layer2=2, planes_n=planes, sp_parts=sps[i % 2])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(10):
                ops.gsl_fused(adj, feats[i % 2], wp, gate, 60, drop_p=p, seed_scorer=1, seed_layer2=2, planes_n=planes, sp_parts=sps[i % 2])
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print("TIME graphs %5d p %.1f planes %d: %8.1f us  %.0f GB/s (%.3f of 6540)" % (
                graphs, p, planes, ms * 1e3, graphs * 280000 / ms / 1e6, graphs * 280000 / ms / 1e6 / 6540.2), flush=True)
