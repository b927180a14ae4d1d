"""Fused GSL kernel at B=32 size with COLD inputs: 16 distinct input sets (1 GB > L2) replayed from one CUDA graph."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import ops, synthetic  # noqa: E402

DEV = "cuda"
w = synthetic.get_workload("snopes")
N, H, G, NSETS = 100, 300, 220, 16
rng = np.random.default_rng(7)
wp, gate = torch.randn(H, device=DEV) * 0.1, torch.randn(12, device=DEV)
sets = []
for s in range(NSETS):
    base = []
    for g in range(G):
        pool = rng.integers(2, w.vocab, size=140)
        toks = pool[rng.integers(0, 140, size=N)]
        base.append(synthetic.word_graph(toks, N, 3)[1].astype(np.float32))
    adj = ops.NeighborLists(torch.from_numpy(np.stack(base)).to(DEV))
    feat = torch.randn(G, N, H, device=DEV)
    sets.append((adj, feat))
for p in (0.0, 0.2):
    for planes in (2, 0):
        recs = []
        ops.PROFILE_GSL_ARGS = recs
        for adj, feat in sets:
            sp = ops.rowdot(feat.view(G * N, H), wp, p, 1)
            ops.gsl_fused(adj, feat, wp, gate, 60, drop_p=p, seed_scorer=1, seed_layer2=2, planes_n=planes, sp_parts=sp)
        ops.PROFILE_GSL_ARGS = None
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for r in recs:
                    ops.gsl_fused_replay(r)
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
        us = a.elapsed_time(b) / len(recs) * 1e3
        print("TIME cold graphs %d p %.1f planes %d: %.1f us" % (G, p, planes, us), flush=True)
