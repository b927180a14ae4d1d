"""tcgen05 bf16-plane GEMM (get_gemm_bp) vs fp64 on the B200 (`-m gpu`).

mode 3 (three planes per operand, small terms in their own accumulator) must deliver fp32-level accuracy: the bound
(2e-6 of the |a|.|b| magnitude) is far tighter than any wrong descriptor / missing plane product would pass.
mode 2 (two planes) is the 16-bit operand class (bound 2e-5), mode 1 plain bf16 (bound 1e-2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
BOUND = {1: 1.2e-2, 2: 2e-5, 3: 2e-6}


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def _err(out, ref):
    return float((out.detach().cpu().double() - ref).abs().max())


def test_to_planes_roundtrip_and_padding():
    from get_b200 import planes as P
    x = _rand(257, 300, seed=1).to(DEV)
    for n, tol in ((1, 2.0 ** -8), (2, 2.0 ** -16), (3, 2.0 ** -23)):
        pl = P.to_planes(x, n, pad_one=True)
        assert pl.ld == 304 and pl.nplanes == n
        rel = float(((pl.to_float() - x).abs() / x.abs().clamp_min(1e-20)).max())
        assert rel <= tol, (n, rel)
        pad = pl.t[:, :, 300:304].float()
        assert torch.equal(pad[0, :, 0], torch.ones(257, device=DEV)) and float(pad[0, :, 1:].abs().max()) == 0.0
        if n > 1:
            assert float(pad[1:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(128, 160, 64), (256, 300, 300), (21600, 300, 300), (1000, 300, 600), (333, 512, 512),
                                   (960, 300, 1628), (5000, 96, 300)])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_bp_gemm_kmajor_plain(M, N, K, mode):
    from get_b200 import planes as P
    a, w, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    ref = a.double() @ w.double().t() + bias.double()
    mag = float((a.double().abs() @ w.double().abs().t()).max())
    ap = P.to_planes(a.to(DEV), 3)
    wd, bd = w.to(DEV), bias.to(DEV)
    pk = P.pack_of(wd, bd)
    out = torch.empty(M, N, device=DEV)
    po = P.alloc_planes(3, M, N, DEV, ld=P.round_up(N + 1, 8))          # room for the ones column
    P.gemm_bp([(ap, pk.planes, K)], M, N, mode=mode, C=out, bias=pk.bias, planes_out=po, planes_out_n=3, pad_one=True)
    assert _err(out, ref) <= BOUND[mode] * mag * max(1.0, K / 512.0), (mode, _err(out, ref), mag)
    # the planes written by the epilogue carry the same values (3 planes = fp32) and the padding contract
    assert float((po.to_float() - out).abs().max()) <= 2.0 ** -22 * float(out.abs().max())
    assert torch.equal(po.t[0, :, N].float(), torch.ones(M, device=DEV))          # the ones column
    assert float(po.t[1:, :, N:].float().abs().max()) == 0.0 and float(po.t[0, :, N + 1:].float().abs().max()) == 0.0


@pytest.mark.parametrize("kblock", [32, 64])
def test_bp_gemm_segments_accumulate_kblock(kblock):
    from get_b200 import planes as P
    M, H = 2100, 300
    a, x, g = (_rand(M, H, seed=s, scale=0.5) for s in (1, 2, 3))
    w0, w1, w2 = (_rand(H, H, seed=s, scale=H ** -0.5) for s in (4, 5, 6))
    ap, xp, gp = (P.to_planes(t.to(DEV), 3) for t in (a, x, g))
    w0d, w1d, w2d = (t.to(DEV) for t in (w0, w1, w2))
    c0 = _rand(M, H, seed=10)
    c = c0.clone().to(DEV)
    # backward style: three segments through TRANSPOSED weight views, accumulate into C
    segs = [(ap, P.pack_of(w0d.t()).planes, H), (xp, P.pack_of(w1d.t()).planes, H), (gp, P.pack_of(w2d.t()).planes, H)]
    P.gemm_bp(segs, M, H, mode=3, C=c, accumulate=True, kblock=kblock)
    ref = c0.double() + a.double() @ w0.double() + x.double() @ w1.double() + g.double() @ w2.double()
    mag = float((a.double().abs() @ w0.double().abs()).max()) * 3
    assert _err(c, ref) <= 2e-6 * mag, _err(c, ref)
    # two planes of a 3-plane tensor (mode 2 reads planes 0, 1 only)
    c2 = torch.empty(M, H, device=DEV)
    P.gemm_bp(segs[:2], M, H, mode=2, C=c2, kblock=kblock)
    ref2 = a.double() @ w0.double() + x.double() @ w1.double()
    assert _err(c2, ref2) <= 2e-5 * mag, _err(c2, ref2)


def test_bp_gemm_epilogues():
    from get_b200 import _lib as L
    from get_b200 import planes as P
    M, H = 3000, 300
    a, x, rx = (_rand(M, H, seed=s, scale=0.5) for s in (1, 2, 3))
    wz0, wz1, wr0, wr1, wh0, wh1 = (_rand(H, H, seed=s, scale=H ** -0.5) for s in range(4, 10))
    bz, br, bh = (_rand(H, seed=s) for s in (11, 12, 13))
    dev = lambda t: t.to(DEV)
    # [a | x] side by side in one plane buffer (columns 0..303 and 304..607), as the GGNN layer keeps them
    ax = P.alloc_planes(3, M, 608, DEV, zero=True)
    P.to_planes(dev(a), 3, pad_one=True, out=ax.view_cols(0, H))
    P.to_planes(dev(x), 3, out=ax.view_cols(304, H))
    for mode in (2, 3):
        bn = P.tile_n(M, H, mode)
        gs = P.round_up(304, bn)
        # fused z|r: weights stacked with group stride gs; K runs over the whole [a | x] buffer (608 columns)
        wd = [dev(t) for t in (wz0, wz1, wr0, wr1)]
        bd = [dev(t) for t in (bz, br)]

        def build():
            blocks = [(wd[0], 0, 0), (wd[1], 0, 304), (wd[2], gs, 0), (wd[3], gs, 304)]
            return 2 * gs, 608, blocks, 2 * gs, [(bd[0], None, 0), (bd[1], None, gs)]
        pk = P.get_pack(("test_zr", mode), build)
        z, r = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
        rxp = P.alloc_planes(3, M, H, DEV)
        xd = dev(x)
        P.gemm_bp([(ax.view_cols(0, 608), pk.planes, 608)], M, 2 * gs, mode=mode, epilogue=L.BPE_ZR, C=z, out1=r, bias=pk.bias,
                  aux0=xd, planes_out=rxp, planes_out_n=3, zr=(gs, H), tn=bn)
        vz = torch.sigmoid(a.double() @ wz0.double().t() + x.double() @ wz1.double().t() + bz.double())
        vr = torch.sigmoid(a.double() @ wr0.double().t() + x.double() @ wr1.double().t() + br.double())
        tol = 8e-6 if mode == 3 else 4e-5
        assert _err(z, vz) < tol and _err(r, vr) < tol, (mode, _err(z, vz), _err(r, vr))
        assert _err(rxp.to_float(), vr * x.double()) < tol
        assert float(rxp.t[:, :, 300:304].float().abs().max()) == 0.0
        # h gate: two segments ([a] columns of the shared buffer, r*x planes), tanh + GRU blend, planes of the output
        pk0, pk1 = P.pack_of(dev(wh0), dev(bh)), P.pack_of(dev(wh1))
        out, hh = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
        op = P.alloc_planes(2, M, H, DEV)
        rxp2 = P.to_planes(dev(rx), 3)
        P.gemm_bp([(ax.view_cols(0, H), pk0.planes, H), (rxp2, pk1.planes, H)], M, H, mode=mode, epilogue=L.BPE_TANH_BLEND,
                  C=out, out1=hh, bias=pk0.bias, aux0=z, aux1=xd, planes_out=op)
        vh = torch.tanh(a.double() @ wh0.double().t() + rx.double() @ wh1.double().t() + bh.double())
        vo = vh * z.cpu().double() + x.double() * (1 - z.cpu().double())
        tol = 2e-5 if mode == 3 else 6e-5
        assert _err(hh, vh) < tol and _err(out, vo) < tol, (mode, _err(hh, vh), _err(out, vo))
        assert _err(op.to_float(), vo) < 1e-4
    # backward epilogue: g = dh' @ Wh1; dr' = g*x*r*(1-r) -> planes (column block of the gate-gradient buffer); dx += g*r
    dh = _rand(M, H, seed=20, scale=0.1)
    rr = torch.sigmoid(_rand(M, H, seed=21))
    dx0 = _rand(M, H, seed=22, scale=0.1)
    dg = P.alloc_planes(2, M, 912, DEV, zero=True)
    dx = dev(dx0.clone())
    P.gemm_bp([(P.to_planes(dev(dh), 2), P.pack_of(dev(wh1).t()).planes, H)], M, H, mode=2, epilogue=L.BPE_DGATE_R,
              aux0=dev(x), aux1=dev(rr), out1=dx, planes_out=dg.view_cols(304, H))
    gref = dh.double() @ wh1.double()
    assert _err(dg.view_cols(304, H).to_float(), gref * x.double() * rr.double() * (1 - rr.double())) < 2e-5
    assert _err(dx, dx0.double() + gref * rr.double()) < 2e-5
    assert float(dg.t[:, :, :304].float().abs().max()) == 0.0 and float(dg.t[:, :, 608:].float().abs().max()) == 0.0
    # attention projection: tanh with a per-group row bias, column-sliced weight
    Pn, G = 100, 30
    w1cat = dev(_rand(H, 2 * H, seed=30, scale=(2 * H) ** -0.5))
    lp = dev(_rand(G, H, seed=31))
    t = torch.empty(M, H, device=DEV)
    P.gemm_bp([(ax.view_cols(0, H), P.pack_of(w1cat[:, H:]).planes, H)], M, H, mode=2, epilogue=L.BPE_TANH_ROWGROUP, C=t,
              aux0=lp, group_rows=Pn)
    ref = torch.tanh(a.double() @ w1cat[:, H:].cpu().double().t() + lp.cpu().double().repeat_interleave(Pn, 0))
    assert _err(t, ref) < 4e-5
    # dX through dropout (mask in the epilogue)
    from get_b200.dropout import keep_mask
    p, seed = 0.2, 4242
    mask = keep_mask(M * H, p, seed).view(M, H)
    o2 = torch.empty(M, H, device=DEV)
    P.gemm_bp([(ax.view_cols(0, H), P.pack_of(dev(wz0).t()).planes, H)], M, H, mode=2, C=o2, drop_out=(p, seed))
    assert _err(o2, (a.double() @ wz0.double()) * mask.double()) < 4e-5


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("Kr,Ma,Nb", [(21600, 912, 304), (4100, 300, 300), (960, 300, 1632), (77, 128, 64)])
def test_bp_weight_gradient_mn_major(Kr, Ma, Nb, mode):
    """dW = dG^T X with both operands stored (rows, cols) row-major and consumed MN-major; blocks of the result (and a
    bias row through the ones column) land in separate destinations."""
    from get_b200 import planes as P
    dg, x = _rand(Kr, Ma, seed=1, scale=0.1), _rand(Kr, Nb, seed=2, scale=0.5)
    x[:, -1] = 1.0                       # "ones column": the last column of the product is the column sum of dG
    ref = dg.double().t() @ x.double()
    mag = float((dg.double().abs().t() @ x.double().abs()).max())
    dgp, xp = P.to_planes(dg.to(DEV), 2), P.to_planes(x.to(DEV), 2)
    h = Ma // 3 // 4 * 4 if Ma >= 12 else Ma
    w_a = torch.zeros(h, Nb - 1 if Nb > 8 else Nb, device=DEV)
    w_b = torch.ones(Ma - h, Nb, device=DEV)
    b_a = torch.zeros(h, device=DEV)
    dsts = [(w_a, 0, h, 0, w_a.shape[1]), (w_b, h, Ma - h, 0, Nb), (b_a, 0, h, Nb - 1, 1)]
    P.wgrad_bp(dgp.T(), xp.T(), Ma, Nb, Kr, mode, dsts, accumulate=False)
    tol = BOUND[mode] * mag
    assert _err(w_a, ref[:h, :w_a.shape[1]]) <= tol, (_err(w_a, ref[:h, :w_a.shape[1]]), tol)
    assert _err(w_b, ref[h:]) <= tol
    assert _err(b_a, ref[:h, Nb - 1]) <= tol
    P.wgrad_bp(dgp.T(), xp.T(), Ma, Nb, Kr, mode, dsts[1:2], accumulate=True)
    assert _err(w_b, 2 * ref[h:]) <= 2 * tol


def test_bp_weight_pack_refresh():
    from get_b200 import planes as P
    M, K, N = 512, 300, 300
    a, w = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.1)
    ap, wd = P.to_planes(a.to(DEV), 3), w.to(DEV)
    out = torch.empty(M, N, device=DEV)
    P.gemm_bp([(ap, P.pack_of(wd).planes, K)], M, N, mode=3, C=out)
    with torch.no_grad():
        wd.mul_(2.0)                       # in-place edit: the version counter moves
    P.gemm_bp([(ap, P.pack_of(wd).planes, K)], M, N, mode=3, C=out)
    mag = float((a.double().abs() @ (2 * w).double().abs().t()).max())
    assert _err(out, a.double() @ (2 * w.double()).t()) <= 2e-6 * mag
    wd.data.mul_(0.5)                      # fused-optimizer style update: no version bump, epoch bump instead
    P.weights_updated()
    P.gemm_bp([(ap, P.pack_of(wd).planes, K)], M, N, mode=3, C=out)
    assert _err(out, a.double() @ w.double().t()) <= 2e-6 * mag
