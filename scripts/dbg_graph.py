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
    dense = adj

    def timed(fn, n=10):
        """n launches captured into a CUDA graph: device time without the ~10 us of host time per Python launch"""
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(n):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    if os.environ.get("GET_B200_GRAPH_LISTS", "1") != "0":
        adj = ops.NeighborLists(dense)
        print("TIME graphs %5d build lists: %8.1f us" % (graphs, timed(adj.rebuild)), flush=True)
        xs = torch.randn(graphs, N, H, device=DEV)
        from get_b200.planes import alloc_planes
        plx = alloc_planes(2, graphs * N, H, DEV, ld=304)
        outx = torch.empty_like(xs)
        kp = (torch.rand(graphs, N, device=DEV) < 0.8).to(torch.uint8)
        print("TIME graphs %5d aggregate lists -> planes: %8.1f us   dense: %8.1f us" % (
            graphs, timed(lambda: ops.graph_aggregate(adj, xs, None, planes_out=plx, pad_one=True, want_f32=False)),
            timed(lambda: ops.graph_aggregate(dense, xs, None, planes_out=plx, pad_one=True, want_f32=False))), flush=True)
        print("TIME graphs %5d aggregate^T masked accumulate lists -> f32+planes: %8.1f us   dense: %8.1f us" % (
            graphs, timed(lambda: ops.graph_aggregate(adj, xs, kp, out=outx, transpose=True, accumulate=True, planes_out=plx)),
            timed(lambda: ops.graph_aggregate(dense, xs, kp, out=outx, transpose=True, accumulate=True, planes_out=plx))), flush=True)
    feats = [torch.randn(graphs, N, H, device=DEV) for _ in range(2)]
    for p in (0.0, 0.2):
        sps = [ops.rowdot(f.view(graphs * N, H), wp, p, 1) for f in feats]
        for planes in (2, 0):
            recs = []
            ops.PROFILE_GSL_ARGS = recs
            for i in range(2):
                ops.gsl_fused(adj, feats[i], wp, gate, 60, drop_p=p, seed_scorer=1, seed_layer2=2, planes_n=planes, sp_parts=sps[i])
            ops.PROFILE_GSL_ARGS = None
            cnt = [0]

            def one():
                ops.gsl_fused_replay(recs[cnt[0] % 2])
                cnt[0] += 1
            us = timed(one)
            print("TIME graphs %5d p %.1f planes %d: %8.1f us  %.0f GB/s (%.3f of 6540) [280 KB/graph]" % (
                graphs, p, planes, us, graphs * 280000 / us / 1e3, graphs * 280000 / us / 1e3 / 6540.2), flush=True)
