"""Bring-up probe of the plane GEMM (run on the GPU box): each variant in its own subprocess-free pass, errors printed.
Descriptor experiments are read from the environment on every launch (gemm_bp.cu: bp_plan)."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import planes as P  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def kmajor(M, N, K, mode, kblock):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    ap = P.to_planes(a.to(DEV), 3)
    out = torch.empty(M, N, device=DEV)
    P.gemm_bp([(ap, P.pack_of(w.to(DEV)).planes, K)], M, N, mode=mode, C=out, kblock=kblock)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    mag = float((a.double().abs() @ w.double().abs().t()).max())
    return float((out.cpu().double() - ref).abs().max()) / mag


def mnmajor(Kr, Ma, Nb, mode):
    dg, x = rnd(Kr, Ma, seed=1, scale=0.1), rnd(Kr, Nb, seed=2, scale=0.5)
    out = torch.zeros(Ma, Nb, device=DEV)
    P.wgrad_bp(P.to_planes(dg.to(DEV), 2).T(), P.to_planes(x.to(DEV), 2).T(), Ma, Nb, Kr, mode, [(out, 0, Ma, 0, Nb)], False)
    torch.cuda.synchronize()
    ref = dg.double().t() @ x.double()
    mag = float((dg.double().abs().t() @ x.double().abs()).max())
    return float((out.cpu().double() - ref).abs().max()) / mag


def attempt(name, fn):
    try:
        print("%-60s rel err %.3e" % (name, fn()), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-60s FAILED: %s" % (name, str(e).splitlines()[-1][:200]), flush=True)
        traceback.print_exc()


if __name__ == "__main__":
    for kb in (32, 64):
        for mode in (1, 2, 3):
            attempt("K-major M=512 N=160 K=256 mode %d kblock %d" % (mode, kb), lambda: kmajor(512, 160, 256, mode, kb))
    attempt("K-major M=21600 N=300 K=300 mode 2", lambda: kmajor(21600, 300, 300, 2, 0))
    for env in ({}, {"GET_B200_BP_MN_SWAP": "1"}, {"GET_B200_BP_MN_SBO": "512"}, {"GET_B200_BP_KB": "64"}):
        for k, v in env.items():
            os.environ[k] = v
        attempt("MN-major K=1024 M=128 N=64 mode 1 env %s" % env, lambda: mnmajor(1024, 128, 64, 1))
        attempt("MN-major K=4100 M=912 N=304 mode 2 env %s" % env, lambda: mnmajor(4100, 912, 304, 2))
        for k in env:
            del os.environ[k]
