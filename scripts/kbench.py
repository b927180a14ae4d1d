#!/usr/bin/env python
"""Kernel micro-benchmarks on cuda:0 (CUDA events on the launch stream, rotating buffer sets larger than L2).

  python scripts/kbench.py graph [--G 216] [--N 100] [--H 300] [--p 0.2]
  python scripts/kbench.py gemm  [--M 21600] [--N 300] [--K 300] [--seg 2]

Prints one JSON line per measurement. Development tool; bench.py is the judged harness.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from get_b200 import _lib, ops, synthetic  # noqa: E402

PEAK = 6540.2
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_fn(fn, nsets, iters=20, warmup=5):
    for i in range(warmup):
        fn(i % nsets)
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i % nsets); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in evs)
    return float(np.median(t)), t[0]


def make_graphs(G, N, H, window=3, seed=0):
    """Reference-shaped graphs: adjacency from the restated convert_text, features N(0,1)."""
    rng = np.random.default_rng(seed)
    adj = np.zeros((G, N, N), np.float32)
    base = []
    for g in range(min(G, 64)):
        pool = rng.integers(2, 5000, size=int(N * 1.4))
        toks = pool[rng.integers(0, pool.shape[0], size=N)]
        _, a, _ = synthetic.word_graph(toks, N, window)
        base.append(a.astype(np.float32))
    for g in range(G):
        adj[g] = base[g % len(base)]
    return torch.from_numpy(adj)


def bench_graph(a):
    G, N, H = a.G, a.N, a.H
    dev = "cuda"
    adj0 = make_graphs(G, N, H, a.window)
    per_set = G * (2 * N * H + N * N) * 4
    nsets = max(2, int(300e6 // per_set) + 1)
    adjs = [adj0.to(dev) for _ in range(nsets)]
    feats = [torch.randn(G, N, H, device=dev) for _ in range(nsets)]
    wp = torch.randn(H, device=dev) * 0.1
    gate = torch.randn(12, device=dev)
    k = int(a.rate * N)
    res = {}
    for p in ([0.0, a.p] if a.p > 0 else [0.0]):
        med, best = time_fn(lambda i: ops.gsl_fused(adjs[i], feats[i], wp, gate, k, drop_p=p, seed_scorer=1, seed_layer2=2,
                                                    want_score=True), nsets, a.iters)
        by = G * 4 * N * (2 * H + N)
        res["fused_p%.1f" % p] = {"ms": med, "best_ms": best, "GBs": by / med / 1e6, "frac": by / med / 1e6 / PEAK}
    med, best = time_fn(lambda i: ops.graph_aggregate(adjs[i], feats[i]), nsets, a.iters)
    by = G * 4 * N * (2 * H + N)
    res["aggregate"] = {"ms": med, "best_ms": best, "GBs": by / med / 1e6, "frac": by / med / 1e6 / PEAK}
    keep = (torch.rand(G, N, device=dev) < 0.6).to(torch.uint8)
    outs = [torch.randn(G, N, H, device=dev) for _ in range(nsets)]
    med, best = time_fn(lambda i: ops.graph_aggregate(adjs[i], feats[i], keep, out=outs[i], transpose=True, accumulate=True),
                        nsets, a.iters)
    res["aggregate_T_acc_keep"] = {"ms": med, "best_ms": best, "GBs": (by + G * N * H * 4) / med / 1e6}
    print(json.dumps({"kernel": "graph", "G": G, "N": N, "H": H, "nsets": nsets, "peak": PEAK, **res}))


def bench_gemm(a):
    M, N, K, dev = a.M, a.N, a.K, "cuda"
    nsets = max(2, int(300e6 // (M * (a.seg * K + 2 * N) * 4)) + 1)
    As = [[torch.randn(M, K, device=dev) for _ in range(a.seg)] for _ in range(nsets)]
    Ws = [torch.randn(N, K, device=dev) * K ** -0.5 for _ in range(a.seg)]
    outs = [torch.empty(M, N, device=dev) for _ in range(nsets)]
    b0, b1 = torch.randn(N, device=dev), torch.randn(N, device=dev)
    aux = [torch.randn(M, N, device=dev) for _ in range(nsets)]
    out1 = [torch.empty(M, N, device=dev) for _ in range(nsets)]
    res = {}
    flops = 2.0 * M * N * K * a.seg
    for tc in (True, False):
        ops.TC_ENABLED = tc
        med, best = time_fn(lambda i: ops.gemm([(As[i][s], Ws[s]) for s in range(a.seg)], outs[i], tc=True), nsets, a.iters)
        res["store_tc%d" % tc] = {"ms": med, "TFLOPs_fp32": flops / med / 1e9}
        med, best = time_fn(lambda i: ops.gemm([(As[i][s], Ws[s]) for s in range(a.seg)], outs[i], epilogue=_lib.EPI_SIGMOID,
                                               bias0=b0, bias1=b1, aux0=aux[i], out1=out1[i], tc=True), nsets, a.iters)
        res["sigmoid_tc%d" % tc] = {"ms": med, "TFLOPs_fp32": flops / med / 1e9}
    ops.TC_ENABLED = True
    for nt in (2, 3, 4, 5):
        med, best = time_fn(lambda i: ops.gemm([(As[i][s], Ws[s]) for s in range(a.seg)], outs[i], tc=True, tc_n_tiles=nt),
                            nsets, a.iters)
        res["store_nt%d" % nt] = {"ms": med, "TFLOPs_fp32": flops / med / 1e9}
    med, best = time_fn(lambda i: ops.gemm([(As[i][s], Ws[s]) for s in range(a.seg)], outs[i], tc=True, presplit=False, split_k=1),
                        nsets, a.iters)
    res["store_rawB"] = {"ms": med, "TFLOPs_fp32": flops / med / 1e9}
    # weight-gradient shape: (N x N) = dg^T (N x M) @ act (M x N)
    dg = [torch.randn(M, N, device=dev) for _ in range(nsets)]
    act = [torch.randn(M, N, device=dev) for _ in range(nsets)]
    w = torch.empty(N, N, device=dev)
    med, best = time_fn(lambda i: ops.gemm([(dg[i].t(), act[i].t())], w, tc=True, presplit=False), nsets, a.iters)
    res["wgrad"] = {"ms": med, "TFLOPs_fp32": 2.0 * M * N * N / med / 1e9}
    print(json.dumps({"kernel": "gemm", "M": M, "N": N, "K": K, "seg": a.seg, "nsets": nsets, **res}))


def bench_accuracy(a):
    """max |err| / max sum|a||b| of the tcgen05 3xTF32 path and the SIMT fp32 path against fp64."""
    dev = "cuda"
    res = {}
    for (M, N, K) in [(21600, 300, 300), (21600, 300, 600), (5000, 96, 1628)]:
        g = torch.Generator().manual_seed(1)
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) * K ** -0.5
        ref = A.double() @ W.double().t()
        mag = float((A.double().abs() @ W.double().abs().t()).max())
        out = torch.empty(M, N, device=dev)
        for tc in (True, False):
            ops.TC_ENABLED = tc
            ops.gemm([(A.to(dev), W.to(dev))], out, tc=True)
            err = (out.cpu().double() - ref).abs()
            res["%dx%dx%d_tc%d" % (M, N, K, tc)] = {"max_rel_mag": float(err.max()) / mag, "rms": float(err.pow(2).mean().sqrt()),
                                                   "mag": mag}
        t32 = (A.to(dev) @ W.to(dev).t()).cpu().double()
        res["%dx%dx%d_torch" % (M, N, K)] = {"max_rel_mag": float((t32 - ref).abs().max()) / mag}
    ops.TC_ENABLED = True
    print(json.dumps({"kernel": "gemm_accuracy", **res}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["graph", "gemm", "accuracy"])
    ap.add_argument("--G", type=int, default=216)
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--H", type=int, default=300)
    ap.add_argument("--M", type=int, default=21600)
    ap.add_argument("--K", type=int, default=300)
    ap.add_argument("--seg", type=int, default=2)
    ap.add_argument("--p", type=float, default=0.2)
    ap.add_argument("--rate", type=float, default=0.6)
    ap.add_argument("--window", type=int, default=3)
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    if a.what == "accuracy":
        torch.backends.cuda.matmul.allow_tf32 = False
        bench_accuracy(a)
    elif a.what == "graph":
        bench_graph(a)
    else:
        bench_gemm(a)


if __name__ == "__main__":
    main()
