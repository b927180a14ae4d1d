"""ops.gemm(tc=True) routing onto the tcgen05 bf16-plane GEMM, and tensor-core vs exact SIMT agreement on the whole model
(`-m gpu`). Kernel-level accuracy tests of the plane GEMM itself: tests/test_gpu_gemm_bp.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def test_gemm_routes_large_contractions_to_the_tensor_cores():
    from get_b200 import _lib as L
    from get_b200 import ops
    ops.DEBUG_TC_REPORT = True
    try:
        for (M, N, K) in [(21600, 300, 300), (129, 24, 40), (1000, 300, 52), (4097, 512, 512)]:
            a, b, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
            out = torch.empty(M, N, device=DEV)
            ops.gemm([(a.to(DEV), b.to(DEV))], out, bias0=bias.to(DEV), tc=True, exact=True)
            assert ops.LAST_GEMM_USED_TC == 2, (M, N, K)
            ref = a.double() @ b.double().t() + bias.double()
            mag = float((a.double().abs() @ b.double().abs().t()).max())
            assert float((out.cpu().double() - ref).abs().max()) <= 2e-6 * mag
        # small / odd contractions stay on the exact SIMT kernel
        a, b = _rand(100, 30, seed=5).to(DEV), _rand(20, 30, seed=6).to(DEV)
        o3 = torch.empty(100, 20, device=DEV)
        ops.gemm([(a, b)], o3, tc=True)
        assert ops.LAST_GEMM_USED_TC == 0
        assert torch.allclose(o3.cpu().double(), a.cpu().double() @ b.cpu().double().t(), atol=1e-5)
        # transposed weight view + accumulate (backward style), per-group row bias (attention), weight gradient
        M, H = 2100, 300
        x, w = _rand(M, H, seed=7, scale=0.5).to(DEV), _rand(H, H, seed=8, scale=H ** -0.5).to(DEV)
        c0 = _rand(M, H, seed=9)
        c = c0.clone().to(DEV)
        ops.gemm([(x, w.t())], c, accumulate=True, tc=True)
        assert ops.LAST_GEMM_USED_TC == 2
        mag = float((x.cpu().double().abs() @ w.cpu().double().abs()).max())
        assert float((c.cpu().double() - (c0.double() + x.cpu().double() @ w.cpu().double())).abs().max()) < 2e-5 * mag
        lp = _rand(21, H, seed=10).to(DEV)
        t = torch.empty(M, H, device=DEV)
        ops.gemm([(x, w)], t, epilogue=L.EPI_TANH_ROWGROUP, aux0=lp, group_rows=100, tc=True)
        ref = torch.tanh(x.cpu().double() @ w.cpu().double().t() + lp.cpu().double().repeat_interleave(100, 0))
        assert float((t.cpu().double() - ref).abs().max()) < 5e-5
        dg = _rand(M, H, seed=11, scale=0.1).to(DEV)
        wg = torch.empty(H, 2 * H, device=DEV)
        ops.gemm([(dg.t(), x.t())], wg[:, H:], tc=True, presplit=False)
        assert ops.LAST_GEMM_USED_TC == 2
        wg2 = torch.empty(H, H, device=DEV)
        ops.gemm([(dg.t(), x.t())], wg2, tc=True, presplit=False)
        assert torch.equal(wg[:, H:], wg2), "split-K reduction must be deterministic"
        mag = float((dg.cpu().double().abs().t() @ x.cpu().double().abs()).max())
        assert float((wg2.cpu().double() - dg.cpu().double().t() @ x.cpu().double()).abs().max()) < 2e-5 * mag
    finally:
        ops.DEBUG_TC_REPORT = False


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp32x", 1e-5)])
def test_tc_and_simt_paths_agree_on_model_gradients(precision, tol):
    """Whole model: tensor-core path vs exact SIMT path (toggled at run time). The kept node sets never depend on it."""
    from get_b200 import ops, synthetic
    from get_b200.model import Graph_basedSemantiStructure
    w = synthetic.get_workload("snopes", batch_claims=6, vocab=500, n_article_sources=16)
    torch.manual_seed(3)
    model = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(DEV).eval()
    q, d, l, kw = synthetic.batch_to_torch(synthetic.make_batch(w, seed=3), device=DEV)
    res = {}
    ops.set_precision(precision)
    try:
        for mode in (True, False):
            ops.TC_ENABLED = mode
            model.zero_grad(set_to_none=True)
            logits = model(q, d, **kw)
            ops.cross_entropy(logits, l).backward()
            res[mode] = (logits.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                         model.ggnn_with_gsl.last_keep.clone())
    finally:
        ops.TC_ENABLED = True
        ops.set_precision("fp32")
    assert torch.equal(res[True][2], res[False][2]), "kept node sets must not depend on the GEMM path"
    assert float((res[True][0] - res[False][0]).abs().max()) < tol
    for n, g in res[True][1].items():
        assert float((g - res[False][1][n]).abs().max()) < tol, n
