"""Parity of every C-ABI kernel against the oracle / fp64 restatements, on the B200 (`-m gpu`).
Tolerances: fp32 path 1e-4 absolute (BASELINE.json north_star) unless a tighter one is stated."""
import numpy as np
import pytest
import torch

from helpers import import_oracle

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from get_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(256, 300, 300), (1000, 300, 600), (37, 13, 9), (130, 66, 17), (5, 300, 3556),
                                   (2160, 600, 300)])
def test_gemm_store_nt(M, N, K):
    ops = _ops()
    a, b, bias = _rand(M, K, seed=1), _rand(N, K, seed=2), _rand(N, seed=3)
    ref = a.double() @ b.double().t() + bias.double()
    out = torch.empty(M, N, device=DEV)
    ops.gemm([(a.to(DEV), b.to(DEV))], out, bias0=bias.to(DEV))
    assert torch.allclose(out.cpu().double(), ref, atol=2e-4 * K ** 0.5 / 10, rtol=1e-5)


def test_gemm_layouts_segments_accumulate():
    ops = _ops()
    M, N, K1, K2, K3 = 333, 300, 300, 304, 36
    a1, a2, a3 = _rand(M, K1, seed=1), _rand(K2, M, seed=2), _rand(M, K3, seed=3)      # a2 stored transposed
    b1, b2, b3 = _rand(N, K1, seed=4), _rand(K2, N, seed=5), _rand(N, K3 + 4, seed=6)  # b2 transposed, b3 with ld
    c0 = _rand(M, N, seed=7)
    ref = c0.double() + 0.5 * (a1.double() @ b1.double().t() + a2.double().t() @ b2.double()
                               + a3.double() @ b3[:, :K3].double().t())
    out = c0.clone().to(DEV)
    a2d, b2d, b3d = a2.to(DEV), b2.to(DEV), b3.to(DEV)
    ops.gemm([(a1.to(DEV), b1.to(DEV)), (a2d.t(), b2d.t()), (a3.to(DEV), b3d[:, :K3])], out, alpha=0.5,
             accumulate=True)
    assert torch.allclose(out.cpu().double(), ref, atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("split", [1, 3, 16])
def test_gemm_weight_grad_shape_splitk(split):
    """dW = dY^T @ X: both operands contiguous along the output index, long contraction."""
    ops = _ops()
    M, H = 5000, 300
    dy, x = _rand(M, H, seed=1, scale=0.1), _rand(M, H, seed=2)
    ref = dy.double().t() @ x.double()
    out = torch.empty(H, H, device=DEV)
    dyd, xd = dy.to(DEV), x.to(DEV)
    ops.gemm([(dyd.t(), xd.t())], out, split_k=split)
    assert torch.allclose(out.cpu().double(), ref, atol=2e-4, rtol=1e-5)
    out2 = torch.empty(H, H, device=DEV)
    ops.gemm([(dyd.t(), xd.t())], out2, split_k=split)
    assert torch.equal(out, out2), "split-K reduction must be deterministic"


def test_gemm_epilogues():
    ops = _ops()
    from get_b200 import _lib as L
    M, N, K = 300, 300, 300
    a, b = _rand(M, K, seed=1, scale=0.2), _rand(N, K, seed=2, scale=0.2)
    b0, b1 = _rand(N, seed=3), _rand(N, seed=4)
    x, z = _rand(M, N, seed=5), torch.sigmoid(_rand(M, N, seed=6))
    v = a.double() @ b.double().t() + b0.double() + b1.double()
    ad, bd, xd, zd = a.to(DEV), b.to(DEV), x.to(DEV), z.to(DEV)
    # sigmoid (+ mul)
    out, out1 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm([(ad, bd)], out, epilogue=L.EPI_SIGMOID, bias0=b0.to(DEV), bias1=b1.to(DEV), aux0=xd, out1=out1)
    assert torch.allclose(out.cpu().double(), torch.sigmoid(v), atol=1e-5)
    assert torch.allclose(out1.cpu().double(), torch.sigmoid(v) * x.double(), atol=1e-5)
    # tanh blend
    ops.gemm([(ad, bd)], out, epilogue=L.EPI_TANH_BLEND, bias0=b0.to(DEV), bias1=b1.to(DEV), aux0=zd, aux1=xd, out1=out1)
    h = torch.tanh(v)
    assert torch.allclose(out1.cpu().double(), h, atol=1e-5)
    assert torch.allclose(out.cpu().double(), h * z.double() + x.double() * (1 - z.double()), atol=1e-5)
    # tanh with per-group row bias
    P = 100
    lp = _rand(M // P, N, seed=8)
    ops.gemm([(ad, bd)], out, epilogue=L.EPI_TANH_ROWGROUP, aux0=lp.to(DEV), group_rows=P)
    ref = torch.tanh(a.double() @ b.double().t() + lp.double().repeat_interleave(P, 0))
    assert torch.allclose(out.cpu().double(), ref, atol=1e-5)
    # GGNN backward r-gate epilogue
    r = torch.sigmoid(_rand(M, N, seed=9))
    dx0 = _rand(M, N, seed=10)
    dxd = dx0.clone().to(DEV)
    ops.gemm([(ad, bd)], out, epilogue=L.EPI_DGATE_R, aux0=xd, aux1=r.to(DEV), out1=dxd)
    g = a.double() @ b.double().t()
    assert torch.allclose(out.cpu().double(), g * x.double() * r.double() * (1 - r.double()), atol=1e-5)
    assert torch.allclose(dxd.cpu().double(), dx0.double() + g * r.double(), atol=1e-5)


def test_gemm_gather_dropout_matches_host_mirror():
    ops = _ops()
    from get_b200.dropout import keep_mask
    from get_b200 import _lib as L
    V, M, K, N, p, seed = 50, 700, 300, 300, 0.2, 987654321
    table, w = _rand(V, K, seed=1), _rand(N, K, seed=2, scale=0.1)
    ids = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(3))
    mask = keep_mask(M * K, p, seed).view(M, K)
    dev_mask = ops.dropout_mask(M * K, p, seed, DEV).view(M, K).cpu()
    assert torch.equal(mask, dev_mask), "host mirror of the dropout hash must be bit-identical"
    assert abs(float((mask == 0).float().mean()) - p) < 0.01
    ref = (table[ids].double() * mask.double()) @ w.double().t()
    out = torch.empty(M, N, device=DEV)
    td = table.to(DEV)
    ops.gemm([(ops.Raw(td.data_ptr(), K, 0, (M, K)), w.to(DEV))], out, rowidx=ids.to(DEV), drop_p=p, drop_seed=seed,
             drop_cols=K)
    assert torch.allclose(out.cpu().double(), ref, atol=1e-4)
    # transposed use (dWp^T = Xd^T @ dx) with the same gather + mask
    dx = _rand(M, N, seed=4, scale=0.1)
    wT = torch.empty(K, N, device=DEV)
    dxd = dx.to(DEV)
    ops.gemm([(ops.Raw(td.data_ptr(), K, 0, (M, K)).t(), dxd.t())], wT, rowidx=ids.to(DEV), drop_p=p, drop_seed=seed,
             drop_cols=K)
    ref = (table[ids].double() * mask.double()).t() @ dx.double()
    assert torch.allclose(wT.cpu().double(), ref, atol=1e-4)
    # epilogue dropout (dX through nn.Dropout)
    out2 = torch.empty(M, K, device=DEV)
    wd = w.to(DEV)
    ops.gemm([(dxd, wd.t())], out2, epilogue=L.EPI_DROPOUT_OUT, drop_out_p=p, drop_out_seed=seed)
    ref = (dx.double() @ w.double()) * mask.double()
    assert torch.allclose(out2.cpu().double(), ref, atol=1e-4)


def test_gemm_rejects_bad_arguments():
    ops = _ops()
    a = torch.zeros(4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        ops.gemm([(a, a)], torch.zeros(4, 4, device=DEV, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        ops.gemm([(torch.zeros(4, 4), torch.zeros(4, 4))], torch.zeros(4, 4))   # CPU tensors: no CPU path
    from get_b200 import _lib as L
    with pytest.raises(RuntimeError):
        ops.gemm([(a, a)], torch.zeros(4, 4, device=DEV), epilogue=L.EPI_TANH_BLEND)  # missing aux


# ---------------------------------------------------------------------------------------------
# graph kernels
# ---------------------------------------------------------------------------------------------
def _rand_graphs(G, N, H, seed, density=0.08, real=None):
    rng = np.random.default_rng(seed)
    adj = (rng.random((G, N, N)) < density) * rng.uniform(0.05, 1.0, (G, N, N))
    real = real if real is not None else max(1, int(0.7 * N))
    adj[:, real:, :] = 0
    adj[:, :, real:] = 0
    x = rng.standard_normal((G, N, H))
    return torch.from_numpy(adj.astype(np.float32)), torch.from_numpy(x.astype(np.float32))


@pytest.mark.parametrize("G,N,H", [(7, 100, 300), (3, 11, 13), (4, 30, 300), (2, 200, 512), (2, 250, 64)])
def test_graph_aggregate(G, N, H):
    ops = _ops()
    adj, x = _rand_graphs(G, N, H, seed=N)
    keep = (torch.rand(G, N, generator=torch.Generator().manual_seed(1)) < 0.5)
    mask = (keep.unsqueeze(2) | keep.unsqueeze(1)).double()
    ad, xd, kd = adj.to(DEV), x.to(DEV), keep.to(torch.uint8).to(DEV)
    out = ops.graph_aggregate(ad, xd)
    assert torch.allclose(out.cpu().double(), adj.double() @ x.double(), atol=1e-5)
    out = ops.graph_aggregate(ad, xd, keep=kd)
    assert torch.allclose(out.cpu().double(), (adj.double() * mask) @ x.double(), atol=1e-5)
    base = torch.randn(G, N, H)
    acc = base.clone().to(DEV)
    ops.graph_aggregate(ad, xd, keep=kd, out=acc, transpose=True, accumulate=True)
    ref = base.double() + (adj.double() * mask).transpose(1, 2) @ x.double()
    assert torch.allclose(acc.cpu().double(), ref, atol=1e-5)


def _scorer_sd(H, seed):
    from helpers import named_param_values
    shapes = {"s.proj.linear.weight": (1, H)}
    for gname in ("z0", "z1", "r0", "r1", "h0", "h1"):
        shapes["s.linear%s.linear.weight" % gname] = (1, 1)
        shapes["s.linear%s.linear.bias" % gname] = (1,)
    sd = {k: torch.from_numpy(v) for k, v in named_param_values(shapes, seed).items()}
    return sd


def _pack_gate(sd):
    names = []
    for gname in ("z0", "z1", "r0", "r1", "h0", "h1"):
        names += ["s.linear%s.linear.weight" % gname, "s.linear%s.linear.bias" % gname]
    return torch.cat([sd[n].reshape(-1) for n in names])


@pytest.mark.parametrize("G,N,H,rate", [(9, 100, 300, 0.6), (5, 12, 24, 0.3), (3, 16, 20, 0.9), (2, 200, 512, 0.6)])
def test_gsl_fused_against_oracle(G, N, H, rate):
    ops = _ops()
    O = import_oracle()
    adj, f1 = _rand_graphs(G, N, H, seed=100 + N)
    f1[:, int(0.7 * N):, :] = f1[:, -1:, :]         # pad nodes share one feature row (ties among pads)
    sd = _scorer_sd(H, seed=5)
    score_ref = O.ggnn(adj.double(), f1.double(), {k: v.double() for k, v in sd.items()}, "s")
    k = int(rate * N)
    adj_ref = O.gsl(adj.double(), score_ref, rate)
    agg_ref = adj_ref @ f1.double()
    score, keep, agg = ops.gsl_fused(adj.to(DEV), f1.to(DEV), sd["s.proj.linear.weight"].reshape(-1).to(DEV),
                                     _pack_gate(sd).to(DEV), k)
    assert torch.allclose(score.cpu().double(), score_ref.squeeze(-1), atol=1e-5)
    assert int(keep.sum()) == G * k
    near = O.near_tie_graphs(score_ref.float(), rate)
    keep_ref = torch.zeros(G, N, dtype=torch.bool).scatter_(1, O.gsl_topk(score_ref, rate), True)
    for g in range(G):
        if near[g]:
            continue
        # the refined adjacency must be identical (kept sets may differ only among zero-degree pad nodes)
        m_ref = (keep_ref[g].unsqueeze(1) | keep_ref[g].unsqueeze(0)).double() * adj[g].double()
        kk = keep[g].cpu().bool()
        m_got = (kk.unsqueeze(1) | kk.unsqueeze(0)).double() * adj[g].double()
        assert torch.equal(m_ref, m_got)
        assert torch.allclose(agg[g].cpu().double(), agg_ref[g], atol=1e-5)


def test_gsl_fused_dropout_uses_documented_masks():
    ops = _ops()
    O = import_oracle()
    from get_b200.dropout import keep_mask
    G, N, H, rate, p = 4, 100, 300, 0.6, 0.2
    adj, f1 = _rand_graphs(G, N, H, seed=77)
    sd = _scorer_sd(H, seed=6)
    ms = keep_mask(G * N * H, p, 11).view(G, N, H)
    m2 = keep_mask(G * N * H, p, 22).view(G, N, H)
    score_ref = O.ggnn(adj.double(), f1.double(), {k: v.double() for k, v in sd.items()}, "s", drop_mask=ms.double())
    agg_ref = O.gsl(adj.double(), score_ref, rate) @ (f1.double() * m2.double())
    score, keep, agg = ops.gsl_fused(adj.to(DEV), f1.to(DEV), sd["s.proj.linear.weight"].reshape(-1).to(DEV),
                                     _pack_gate(sd).to(DEV), int(rate * N), drop_p=p, seed_scorer=11, seed_layer2=22)
    assert torch.allclose(score.cpu().double(), score_ref.squeeze(-1), atol=1e-5)
    near = O.near_tie_graphs(score_ref.float(), rate)
    for g in range(G):
        if not near[g]:
            assert torch.allclose(agg[g].cpu().double(), agg_ref[g], atol=1e-5)


def test_gsl_mask_adj_matches_golden():
    from helpers import load_golden
    ops = _ops()
    gold = load_golden("modules")
    adj = torch.from_numpy(gold["ggnn/in/adj"]).to(DEV)
    score = torch.from_numpy(gold["gsl/in/score"]).to(DEV)
    N = adj.shape[-1]
    for rate in (0.3, 0.6, 0.9):
        out, _ = ops.gsl_mask_adj(adj, score, int(rate * N))
        assert np.array_equal(out.cpu().numpy(), gold["gsl/out/adj_%d" % int(rate * 10)])


# ---------------------------------------------------------------------------------------------
# attention pooling
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G,P,X,Dr,H,C", [(6, 100, 300, 300, 300, 5), (4, 30, 428, 1628, 300, 2), (5, 10, 6, 7, 13, 3),
                                          (3, 200, 512, 512, 512, 8), (3, 30, 0, 40, 24, 1)])
def test_concat_att_forward_backward(G, P, X, Dr, H, C):
    ops = _ops()
    O = import_oracle()
    left = _rand(G, X, seed=1) if X else None
    right = _rand(G, P, Dr, seed=2)
    mask = (torch.rand(G, P, generator=torch.Generator().manual_seed(3)) < 0.7)
    mask[:, 0] = True
    w1 = _rand(H, X + Dr, seed=4, scale=(X + Dr) ** -0.5)
    w2 = _rand(C, H, seed=5, scale=H ** -0.5)
    go, ga = _rand(G, Dr, C, seed=6), _rand(G, P, C, seed=7)
    # oracle (fp64 autograd)
    lv = [t.double().requires_grad_(True) if t is not None else None for t in (left, right, w1, w2)]
    if X:
        o_ref, a_ref = O.concat_not_equal_self_att(lv[0], lv[1], mask, lv[2], lv[3])
    else:
        o_ref, a_ref = O.multi_head_self_att_extend(lv[1], mask, lv[2], lv[3], return_att_weights=True)
        o_ref = o_ref.permute(0, 2, 1)
    ((o_ref * go.double()).sum() + (a_ref * ga.double()).sum()).backward()
    dv = [t.to(DEV).requires_grad_(True) if t is not None else None for t in (left, right, w1, w2)]
    o, a = ops.concat_att(dv[0], dv[1], mask.to(DEV), dv[2], dv[3])
    ((o * go.to(DEV)).sum() + (a * ga.to(DEV)).sum()).backward()
    assert torch.allclose(o.detach().cpu().double(), o_ref.detach(), atol=1e-5)
    assert torch.allclose(a.detach().cpu().double(), a_ref.detach(), atol=1e-5)
    assert torch.allclose(a.detach().sum(1).cpu(), torch.ones(G, C), atol=1e-5)     # reference runtime assert
    assert float(a.detach().cpu()[~mask].abs().max()) == 0.0                          # pad positions exactly 0
    for got, ref, name in zip(dv, lv, ("left", "right", "w1", "w2")):
        if got is not None:
            assert torch.allclose(got.grad.cpu().double(), ref.grad, atol=1e-4, rtol=1e-4), name


# ---------------------------------------------------------------------------------------------
# small kernels
# ---------------------------------------------------------------------------------------------
def test_segment_ops_masked_mean_cross_entropy_linear():
    ops = _ops()
    from get_b200.model import Graph_basedSemantiStructure as M
    cnt = torch.tensor([3, 1, 5, 2])
    B, n, W = 4, 6, 20
    b1 = int(cnt.sum())
    seg, slot, off = ops.segments(cnt.to(DEV), b1, n)
    assert seg.cpu().tolist() == [0, 0, 0, 1, 2, 2, 2, 2, 2, 3, 3]
    assert slot.cpu().tolist() == [0, 1, 2, 6, 12, 13, 14, 15, 16, 18, 19]
    assert off.cpu().tolist() == [0, 3, 4, 9, 11]
    O = import_oracle()
    src = _rand(B, W, seed=1).to(DEV).requires_grad_(True)
    out = ops.SegmentExpandFn.apply(src, seg, off)
    ref = O.pad_left(src.detach().cpu(), cnt)
    assert torch.equal(out.detach().cpu(), ref)
    g = _rand(b1, W, seed=2)
    out.backward(g.to(DEV))
    gref = torch.stack([g[off[i]:off[i + 1]].sum(0) for i in range(B)])
    assert torch.allclose(src.grad.cpu(), gref, atol=1e-6)
    rows = _rand(b1, W, seed=3).to(DEV).requires_grad_(True)
    extra = _rand(B * n, 5, seed=4).to(DEV).requires_grad_(True)
    padded = ops.SegmentPadFn.apply(rows, slot, B * n, extra)
    ref = torch.cat([O.pad_right(rows.detach().cpu(), cnt, n).view(B * n, W), extra.detach().cpu()], 1)
    assert torch.equal(padded.detach().cpu(), ref)
    gp = _rand(B * n, W + 5, seed=5)
    padded.backward(gp.to(DEV))
    assert torch.equal(rows.grad.cpu(), gp[slot.cpu().long(), :W])
    assert torch.equal(extra.grad.cpu(), gp[:, W:])
    # masked mean
    h = _rand(B, 7, W, seed=6).to(DEV).requires_grad_(True)
    ids = torch.tensor([[5, 3, 0, 0, 0, 0, 0], [1, 2, 3, 4, 5, 6, 7], [9, 0, 0, 0, 0, 0, 0], [4, 4, 4, 0, 0, 0, 0]])
    lens = torch.tensor([2, 7, 1, 3])
    mm = ops.MaskedMeanFn.apply(h, ids.to(DEV), lens.to(DEV))
    hr = h.detach().cpu().double().requires_grad_(True)
    ref = (hr * (ids > 0).unsqueeze(2).double()).sum(1) / lens.unsqueeze(1).double()
    assert torch.allclose(mm.detach().cpu().double(), ref.detach(), atol=1e-6)
    gm = _rand(B, W, seed=7)
    mm.backward(gm.to(DEV))
    ref.backward(gm.double())
    assert torch.allclose(h.grad.cpu().double(), hr.grad, atol=1e-6)
    # cross entropy
    logits = _rand(37, 2, seed=8).to(DEV).requires_grad_(True)
    labels = torch.randint(0, 2, (37,), generator=torch.Generator().manual_seed(9))
    loss = ops.cross_entropy(logits, labels.to(DEV))
    lr = logits.detach().cpu().double().requires_grad_(True)
    lref = O.cross_entropy(lr, labels)
    loss.backward()
    lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-6
    assert torch.allclose(logits.grad.cpu().double(), lr.grad, atol=1e-7)
    # linear
    x = _rand(32, 3556, seed=10).to(DEV).requires_grad_(True)
    w = _rand(300, 3556, seed=11, scale=0.02).to(DEV).requires_grad_(True)
    b = _rand(300, seed=12).to(DEV).requires_grad_(True)
    y = ops.linear(x, w, b)
    xr, wr, br = (t.detach().cpu().double().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.linear(xr, wr, br)
    gy = _rand(32, 300, seed=13)
    y.backward(gy.to(DEV))
    yr.backward(gy.double())
    assert torch.allclose(y.detach().cpu().double(), yr.detach(), atol=1e-4)
    for got, ref in ((x, xr), (w, wr), (b, br)):
        assert torch.allclose(got.grad.cpu().double(), ref.grad, atol=1e-4)
    cs = ops.colsum(gy.to(DEV))
    assert torch.allclose(cs.cpu().double(), gy.double().sum(0), atol=1e-5)


@pytest.mark.parametrize("T,N,window,vocab", [(100, 100, 3, 140), (100, 100, 5, 60), (100, 100, 9, 3000), (30, 30, 3, 12),
                                              (57, 100, 3, 40), (120, 100, 3, 150)])
def test_device_word_graphs_match_host_construction(T, N, window, vocab):
    """get_build_word_graphs against the restated convert_text + _laplacian_normalize (get_b200.synthetic.word_graph,
    itself pinned to the reference in tests/test_oracle_golden.py): node lists, counts and fp32 adjacencies identical."""
    from get_b200 import synthetic
    from get_b200.graph_build import build_word_graphs
    rng = np.random.default_rng(T * 31 + window)
    G = 23
    toks = rng.integers(2, vocab + 2, size=(G, T)).astype(np.int64)
    lens = rng.integers(1, T + 1, size=(G,)).astype(np.int32)
    lens[0], lens[1] = T, 1
    nodes, adj, nn = build_word_graphs(torch.from_numpy(toks).to(DEV), torch.from_numpy(lens).to(DEV), N, window)
    for g in range(G):
        L = int(min(lens[g], N))
        rn, ra, rc = synthetic.word_graph(toks[g, :L], N, window)
        assert int(nn[g]) == rc
        assert np.array_equal(nodes[g].cpu().numpy(), rn)
        ref32 = torch.from_numpy(ra).float()
        assert torch.equal(adj[g].cpu(), ref32), "graph %d: max diff %g" % (g, float((adj[g].cpu() - ref32).abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("V,E,R", [(512, 128, 990), (7, 128, 33), (3605, 128, 1200), (5, 12, 1)])
def test_embedding_rows_forward_backward_match_torch(V, E, R):
    """Source-embedding lookup (base_model.py:184-188; id -1 reads row 0, gbss.py:166-168) and its dense table gradient:
    duplicates summed in index order (deterministic), untouched rows exactly zero."""
    from get_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(V + R)
    table = torch.randn(V, E, device="cuda", generator=g).requires_grad_(True)
    idx = torch.randint(-1, min(V, 40), (R,), device="cuda", generator=g)          # many duplicates, some -1
    out = ops.embedding_rows(table, idx)
    ref_table = table.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.embedding(idx.clamp(min=0), ref_table)
    assert torch.equal(out, ref)
    w = torch.randn(R, E, device="cuda", generator=g)
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert (table.grad - ref_table.grad).abs().max().item() <= 1e-5 * max(1.0, ref_table.grad.abs().max().item())
    assert (table.grad[40:] == 0).all()
    g1 = table.grad.clone()
    table.grad = None
    (ops.embedding_rows(table, idx) * w).sum().backward()
    assert torch.equal(table.grad, g1)                                             # deterministic
