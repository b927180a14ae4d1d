"""tcgen05 (3xTF32) GEMM path vs fp64, on the B200 (`-m gpu`). The error-compensated split must deliver
fp32-level accuracy: the bound used here (relative 2e-6 of the |a|.|b| magnitude) is ~50x tighter than what a
plain TF32 product would achieve, so a wrong descriptor / missing correction term fails loudly."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def _check(out, ref, mag, what, scale=1.0):
    err = float((out.cpu().double() - ref).abs().max())
    bound = 2e-6 * mag * scale
    assert err <= bound, "%s: max abs err %.3e > %.3e" % (what, err, bound)


@pytest.mark.parametrize("M,N,K,nt", [(128, 160, 32, 0), (256, 300, 300, 0), (21600, 300, 300, 0), (1000, 300, 300, 1),
                                      (333, 300, 300, 3), (700, 512, 512, 0), (130, 16, 64, 0), (5000, 96, 1628, 0),
                                      (960, 300, 600, 4)])
def test_tc_gemm_plain(M, N, K, nt):
    from get_b200 import ops
    ops.DEBUG_TC_REPORT = True
    a, b, bias = _rand(M, K, seed=1), _rand(N, K, seed=2), _rand(N, seed=3)
    ref = a.double() @ b.double().t() + bias.double()
    out = torch.empty(M, N, device=DEV)
    ops.gemm([(a.to(DEV), b.to(DEV))], out, bias0=bias.to(DEV), tc=True, tc_n_tiles=nt)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    mag = float((a.double().abs() @ b.double().abs().t()).max())
    # the tensor core accumulates in fp32 without round-to-nearest: allow the bound to grow with the contraction length
    _check(out, ref, mag, "plain", scale=max(1.0, K / 512.0))
    ops.DEBUG_TC_REPORT = False


def test_tc_gemm_segments_transposed_weights_accumulate_epilogues():
    from get_b200 import _lib as L
    from get_b200 import ops
    ops.DEBUG_TC_REPORT = True
    M, H = 2100, 300
    a, x, rx = (_rand(M, H, seed=s, scale=0.5) for s in (1, 2, 3))
    w0, w1, w2 = (_rand(H, H, seed=s, scale=H ** -0.5) for s in (4, 5, 6))
    b0, b1 = _rand(H, seed=7), _rand(H, seed=8)
    ad, xd, rxd, w0d, w1d, w2d = (t.to(DEV) for t in (a, x, rx, w0, w1, w2))
    # forward style: two segments, two biases, sigmoid + mul epilogue
    r, out1 = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
    ops.gemm([(ad, w0d), (xd, w1d)], r, epilogue=L.EPI_SIGMOID, bias0=b0.to(DEV), bias1=b1.to(DEV), aux0=xd, out1=out1,
             tc=True)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    v = a.double() @ w0.double().t() + x.double() @ w1.double().t() + b0.double() + b1.double()
    assert float((r.cpu().double() - torch.sigmoid(v)).abs().max()) < 8e-6
    assert float((out1.cpu().double() - torch.sigmoid(v) * x.double()).abs().max()) < 5e-6
    # tanh blend
    z = torch.sigmoid(_rand(M, H, seed=9))
    o, hh = torch.empty(M, H, device=DEV), torch.empty(M, H, device=DEV)
    ops.gemm([(ad, w0d), (rxd, w2d)], o, epilogue=L.EPI_TANH_BLEND, bias0=b0.to(DEV), bias1=b1.to(DEV), aux0=z.to(DEV),
             aux1=xd, out1=hh, tc=True)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    v = a.double() @ w0.double().t() + rx.double() @ w2.double().t() + b0.double() + b1.double()
    assert float((hh.cpu().double() - torch.tanh(v)).abs().max()) < 2e-5
    assert float((o.cpu().double() - (torch.tanh(v) * z.double() + x.double() * (1 - z.double()))).abs().max()) < 2e-5
    # backward style: three segments through transposed weight views, accumulate into C
    c0 = _rand(M, H, seed=10)
    c = c0.clone().to(DEV)
    ops.gemm([(ad, w0d.t()), (xd, w1d.t()), (rxd, w2d.t())], c, accumulate=True, tc=True)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    ref = c0.double() + a.double() @ w0.double() + x.double() @ w1.double() + rx.double() @ w2.double()
    mag = float((a.double().abs() @ w0.double().abs()).max()) * 3
    _check(c, ref, mag, "3-seg accumulate")
    # column-sliced weight (attention W1[:, X:]) with per-group row bias
    P, G = 100, 21
    w1cat = _rand(H, 2 * H, seed=11, scale=(2 * H) ** -0.5).to(DEV)
    lp = _rand(G, H, seed=12).to(DEV)
    t = torch.empty(M, H, device=DEV)
    ops.gemm([(ad, w1cat[:, H:])], t, epilogue=L.EPI_TANH_ROWGROUP, aux0=lp, group_rows=P, tc=True)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    ref = torch.tanh(a.double() @ w1cat[:, H:].cpu().double().t() + lp.cpu().double().repeat_interleave(P, 0))
    assert float((t.cpu().double() - ref).abs().max()) < 1e-5
    ops.DEBUG_TC_REPORT = False


def test_tc_gemm_gather_dropout_and_weight_refresh():
    from get_b200 import _lib as L
    from get_b200 import ops
    from get_b200.dropout import keep_mask
    ops.DEBUG_TC_REPORT = True
    V, M, K, N, p, seed = 500, 2160, 300, 300, 0.2, 13579
    table, w = _rand(V, K, seed=1, scale=0.2), _rand(N, K, seed=2, scale=0.1)
    ids = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(3))
    mask = keep_mask(M * K, p, seed).view(M, K)
    td, wd = table.to(DEV), w.to(DEV)
    out = torch.empty(M, N, device=DEV)
    ops.gemm([(ops.Raw(td.data_ptr(), K, 0, (M, K)), wd)], out, rowidx=ids.to(DEV), drop_p=p, drop_seed=seed,
             drop_cols=K, tc=True)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    ref = (table[ids].double() * mask.double()) @ w.double().t()
    _check(out, ref, float(((table[ids].abs().double() * 1.25) @ w.double().abs().t()).max()), "gather+dropout")
    # dX through dropout (epilogue mask) with a transposed weight
    dx = _rand(M, N, seed=4, scale=0.1).to(DEV)
    o2 = torch.empty(M, K, device=DEV)
    ops.gemm([(dx, wd.t())], o2, epilogue=L.EPI_DROPOUT_OUT, drop_out_p=p, drop_out_seed=seed, tc=True)
    assert ops.LAST_GEMM_USED_TC in (1, 2)
    ref = (dx.cpu().double() @ w.double()) * mask.double()
    assert float((o2.cpu().double() - ref).abs().max()) < 5e-6
    # in-place weight update (optimizer step) must refresh the cached split
    with torch.no_grad():
        wd.mul_(2.0)
    ops.gemm([(ops.Raw(td.data_ptr(), K, 0, (M, K)), wd)], out, rowidx=ids.to(DEV), tc=True)
    ref = table[ids].double() @ (2 * w.double()).t()
    _check(out, ref, float((table[ids].abs().double() @ (2 * w.double()).abs().t()).max()), "refreshed weight")
    # ineligible descriptors fall back to the exact SIMT kernel and stay correct
    a = _rand(100, 30, seed=5).to(DEV)          # K < 32 and K % 4 != 0
    b = _rand(20, 30, seed=6).to(DEV)
    o3 = torch.empty(100, 20, device=DEV)
    ops.gemm([(a, b)], o3, tc=True)
    assert ops.LAST_GEMM_USED_TC == 0
    assert torch.allclose(o3.cpu().double(), a.cpu().double() @ b.cpu().double().t(), atol=1e-5)
    ops.DEBUG_TC_REPORT = False


def test_tc_and_simt_paths_agree_on_model_gradients():
    """Whole model, fp32 mode: tensor-core path vs exact SIMT path (GET_B200_TC toggled at run time)."""
    from get_b200 import ops, synthetic
    from get_b200.model import Graph_basedSemantiStructure
    w = synthetic.get_workload("snopes", batch_claims=6, vocab=500, n_article_sources=16)
    torch.manual_seed(3)
    model = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(DEV).eval()
    q, d, l, kw = synthetic.batch_to_torch(synthetic.make_batch(w, seed=3), device=DEV)
    res = {}
    for mode in (True, False):
        ops.TC_ENABLED = mode
        model.zero_grad(set_to_none=True)
        logits = model(q, d, **kw)
        ops.cross_entropy(logits, l).backward()
        res[mode] = (logits.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                     model.ggnn_with_gsl.last_keep.clone())
    ops.TC_ENABLED = True
    assert torch.equal(res[True][2], res[False][2]), "kept node sets must not depend on the GEMM path"
    assert float((res[True][0] - res[False][0]).abs().max()) < 1e-5
    for n, g in res[True][1].items():
        assert float((g - res[False][1][n]).abs().max()) < 1e-5, n


@pytest.mark.parametrize("Mo,No,K,split", [(300, 300, 21600, None), (300, 1628, 960, None), (300, 300, 1000, 1), (152, 96, 520, 3),
                                           (512, 512, 4100, None)])
def test_tc2_weight_gradient_mn_major_splitk(Mo, No, K, split):
    """dW (Mo,No) = dG^T (Mo,K) @ X (K,No): both operands are (K, .) row-major activations, consumed MN-major by the
    persistent tcgen05 kernel (in-kernel hi/lo split of both operands, deterministic split-K)."""
    from get_b200 import ops
    ops.DEBUG_TC_REPORT = True
    dg, x = _rand(K, Mo, seed=21, scale=0.3), _rand(K, No, seed=22, scale=0.5)
    ref = dg.double().t() @ x.double()
    w = torch.empty(Mo, No, device=DEV)
    ops.gemm([(dg.to(DEV).t(), x.to(DEV).t())], w, tc=True, presplit=False, split_k=split)
    assert ops.LAST_GEMM_USED_TC == 2
    mag = float((dg.double().abs().t() @ x.double().abs()).max())
    _check(w, ref, mag, "wgrad", scale=1.0)
    w2 = torch.empty(Mo, No, device=DEV)
    ops.gemm([(dg.to(DEV).t(), x.to(DEV).t())], w2, tc=True, presplit=False, split_k=split)
    assert torch.equal(w, w2), "split-K reduction must be deterministic"
    ops.DEBUG_TC_REPORT = False


def test_tc2_is_the_path_for_plain_weight_gemms_and_handles_tails():
    from get_b200 import _lib as L
    from get_b200 import ops
    ops.DEBUG_TC_REPORT = True
    for (M, N, K) in [(21600, 300, 300), (129, 24, 40), (1000, 300, 52), (4097, 512, 512)]:
        a, b, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
        out = torch.empty(M, N, device=DEV)
        ops.gemm([(a.to(DEV), b.to(DEV))], out, bias0=bias.to(DEV), tc=True)
        assert ops.LAST_GEMM_USED_TC == 2, (M, N, K)
        ref = a.double() @ b.double().t() + bias.double()
        mag = float((a.double().abs() @ b.double().abs().t()).max())
        _check(out, ref, mag, "tc2 plain %s" % ((M, N, K),))
    ops.DEBUG_TC_REPORT = False
