"""Timeline of the plane GEMM's roles (build with GETB_EXTRA_NVCC_FLAGS=-DGETB_BP_TIMELINE, see gemm_bp.cu) and
micro-benchmarks of its configurations:   python scripts/dbg_bp.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from get_b200 import _lib as L  # noqa: E402
from get_b200 import planes as P  # noqa: E402

DEV = "cuda"
M, H = 21600, 300


def bench(name, fn, n=20):
    """GPU time per call: the calls are captured into a CUDA graph (the Python launch path costs more than these kernels)."""
    try:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        print("TIME %-50s %8.1f us" % (name, 1e3 * a.elapsed_time(b) / n), flush=True)
    except RuntimeError as e:
        print("TIME %-50s skipped: %s" % (name, str(e).splitlines()[-1][:80]), flush=True)


x = torch.randn(M, H, device=DEV)
w = torch.randn(H, H, device=DEV) * H ** -0.5
xp = P.to_planes(x, 3)
pk = P.pack_of(w)
out = torch.empty(M, H, device=DEV)
po = P.alloc_planes(2, M, H, DEV)
xar = P.alloc_planes(3, M, 912, DEV, zero=True)
wz = torch.randn(640, 608, device=DEV) * 0.05
pkz = P.pack_of(wz)
z, r = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
rxp = P.alloc_planes(2, M, H, DEV)

for kb in (32, 64):
    for mode in (1, 2, 3):
        bench("x-proj store C only     mode %d kb %d" % (mode, kb), lambda: P.gemm_bp([(xp, pk.planes, H)], M, H, mode=mode, C=out, kblock=kb))
    bench("x-proj store C+planes   mode 2 kb %d" % kb, lambda: P.gemm_bp([(xp, pk.planes, H)], M, H, mode=2, C=out, planes_out=po, kblock=kb))
    bench("x-proj planes only      mode 2 kb %d" % kb, lambda: P.gemm_bp([(xp, pk.planes, H)], M, H, mode=2, planes_out=po, kblock=kb))
    bn = P.tile_n(M, H, 2)
    bench("zr fused K=608 N=640    mode 2 kb %d" % kb,
          lambda: P.gemm_bp([(xar.view_cols(0, 608), pkz.planes, 608)], M, 640, mode=2, epilogue=L.BPE_ZR, C=z, out1=r, aux0=x,
                            planes_out=rxp, zr=(320, H), tn=bn, kblock=kb))
    for tn in (80, 112, 160):
        bench("x-proj C only mode 2 kb %d tile_n %d" % (kb, tn), lambda: P.gemm_bp([(xp, pk.planes, H)], M, H, mode=2, C=out, kblock=kb, tn=tn))
dg = P.alloc_planes(2, M, 912, DEV, zero=True)
dst = torch.empty(912, 304, device=DEV)
bench("wgrad 912x304 K=21600 mode 2", lambda: P.wgrad_bp(dg.T(), xar.view_cols(304, 304).T(), 912, 304, M, 2, [(dst, 0, 912, 0, 304)], False))
dst2 = torch.empty(300, 300, device=DEV)
bench("wgrad 300x300 K=21600 mode 2", lambda: P.wgrad_bp(dg.view_cols(0, 300).T(), xar.view_cols(0, 300).T(), 300, 300, M, 2, [(dst2, 0, 300, 0, 300)], False))
os.environ["GET_B200_BP_DEBUG"] = "9"
for kb in (32, 64):
    print("---- timeline x-proj mode 2 C+planes kb %d" % kb, flush=True)
    P.gemm_bp([(xp, pk.planes, H)], M, H, mode=2, C=out, planes_out=po, kblock=kb)
    torch.cuda.synchronize()
    print("---- timeline zr mode 2 kb %d" % kb, flush=True)
    P.gemm_bp([(xar.view_cols(0, 608), pkz.planes, 608)], M, 640, mode=2, epilogue=L.BPE_ZR, C=z, out1=r, aux0=x, planes_out=rxp,
              zr=(320, H), tn=160, kblock=kb)
    torch.cuda.synchronize()
print("---- timeline wgrad 912x304", flush=True)
P.wgrad_bp(dg.T(), xar.view_cols(304, 304).T(), 912, 304, M, 2, [(dst, 0, 912, 0, 304)], False)
torch.cuda.synchronize()
