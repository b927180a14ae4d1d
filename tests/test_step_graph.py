"""Captured training step (get_b200/step_graph.py): batch padding is semantics-preserving (CPU, oracle), and the
CUDA-graph replay reproduces the eager step bit for bit (GPU)."""
import numpy as np
import pytest
import torch

from get_b200 import synthetic
from get_b200.keywords import KeyWordSettings as K
from get_b200.step_graph import pad_batch


def test_pad_batch_layout_and_oracle_invariance():
    from get_b200.model import Graph_basedSemantiStructure
    from oracle import get_oracle as O
    w = synthetic.get_workload("tiny")
    batch = synthetic.make_batch(w, seed=11)
    B, B1 = batch["query"].shape[0], batch["pairs"]
    for mult in (1, 4, 7, 64):
        p = pad_batch(batch, mult)
        assert p["pairs"] % mult == 0 and p["n_real_claims"] == B and p["real_pairs"] == B1
        assert int(p[K.EvidenceCountPerQuery].sum()) == p["pairs"] == p[K.DocContentNoPaddingEvidence].shape[0]
        assert p[K.Evd_Docs_Adj].shape[0] == p["pairs"] and p["document"].shape[0] == p["query"].shape[0]
        assert (p[K.EvidenceCountPerQuery] >= 1).all() and (p[K.EvidenceCountPerQuery] <= 30).all()
        # every slot listed as real in `document` has a source id, the others -1
        real = p["document"].sum(-1) >= 1
        assert ((p[K.DocSources] >= 0) == real).all()
    torch.manual_seed(3)
    model = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=False))
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = dict(gsl_rate=w.gsl_rate, use_claim_source=w.use_claim_source, use_article_source=w.use_article_source)
    q, d, l, kw = synthetic.batch_to_torch(batch)
    ref = O.model_forward(sd, cfg, q, d, kw)
    ref = ref[0] if isinstance(ref, tuple) else ref
    pq, pd_, pl, pkw = synthetic.batch_to_torch(pad_batch(batch, 7))
    out = O.model_forward(sd, cfg, pq, pd_, pkw)
    out = out[0] if isinstance(out, tuple) else out
    assert out.shape[0] > B
    assert torch.equal(out[:B], ref), "padding must not change the real claims' logits"


@pytest.mark.gpu
@pytest.mark.parametrize("train", [False, True])
def test_captured_step_matches_eager(train):
    from get_b200 import ops
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep
    dev = "cuda"
    w = synthetic.get_workload("snopes", batch_claims=5, vocab=400, n_article_sources=16)
    batch = pad_batch(synthetic.make_batch(w, seed=21), 16)
    nreal = batch["n_real_claims"]
    q, d, l, kw = synthetic.batch_to_torch(batch, device=dev)

    def fresh():
        torch.manual_seed(9)
        m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev)
        m.train(train)
        m.dropout_seeds = dict(claim=1, feat_prop1=2, word_scorer1=3, feat_prop2=4)
        return m

    # eager reference with the salt the captured step will see after its advance: v1 = v0*1664525 + 1013904223
    v0 = 12345
    v1 = (v0 * 1664525 + 1013904223) & 0xFFFFFFFF
    m1 = fresh()
    ops.dropout_salt_set(v1)
    logits = m1(q, d, **kw)
    loss1 = ops.cross_entropy(logits[:nreal], l[:nreal])
    loss1.backward()
    g1 = {n: p.grad.clone() for n, p in m1.named_parameters() if p.grad is not None}

    m2 = fresh()
    ops.dropout_salt_set(v0)
    step = CapturedTrainStep(m2)
    loss2 = step.step(q, d, l, kw, nreal).clone()
    assert ops.dropout_salt_get() == v1
    g2 = {n: p.grad.clone() for n, p in m2.named_parameters() if p.grad is not None}
    assert torch.equal(loss1, loss2)
    assert g1.keys() == g2.keys()
    for n in g1:
        assert torch.equal(g1[n], g2[n]), n
    # a second replay on the same inputs: same loss in eval mode, new dropout masks (salt advanced) in train mode
    loss3 = step.step(q, d, l, kw, nreal).clone()
    assert step.n_graphs() == 1
    if train:
        assert not torch.equal(loss3, loss2)
    else:
        assert torch.equal(loss3, loss2)
    ops.dropout_salt_set(0)


@pytest.mark.gpu
def test_captured_step_with_optimizer_trains_like_eager():
    from get_b200 import ops
    from get_b200.ddp import FlatGradAllReduce, trainable_named_parameters
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep
    dev = "cuda"
    w = synthetic.get_workload("snopes", batch_claims=4, vocab=300, n_article_sources=8)
    batches = [pad_batch(synthetic.make_batch(w, seed=s), 32) for s in (1, 2)]
    tens = [synthetic.batch_to_torch(b, device=dev) for b in batches]

    def run(captured):
        torch.manual_seed(4)
        m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev).eval()
        params = [p for _, p in trainable_named_parameters(m)]
        opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)
        red = FlatGradAllReduce(params)
        stepper = CapturedTrainStep(m, opt, red) if captured else None
        losses = []
        for it in range(6):
            b, (q, d, l, kw) = batches[it % 2], tens[it % 2]
            n = b["n_real_claims"]
            if captured:
                losses.append(float(stepper.step(q, d, l, kw, n)))
            else:
                opt.zero_grad(set_to_none=True)
                loss = ops.cross_entropy(m(q, d, **kw)[:n], l[:n])
                loss.backward()
                red.reduce()
                opt.step()
                losses.append(float(loss))
        return losses, [p.detach().clone() for p in params]

    ops.TC_ENABLED = False            # exact SIMT GEMMs, no cached weight splits: the ground truth trajectory
    ls, ps = run(False)
    ops.TC_ENABLED = True
    le, pe = run(False)
    lc, pc = run(True)
    assert np.allclose(ls, le, rtol=0, atol=2e-5), ("eager tensor-core path drifted from the SIMT path", ls, le)
    assert np.allclose(le, lc, rtol=0, atol=1e-6), ("captured step drifted from the eager step", le, lc)
    for a, b in zip(pe, pc):
        assert float((a - b).abs().max()) < 1e-6


@pytest.mark.gpu
def test_bounded_graph_cache_evicts_and_recaptures():
    """real Snopes batches span 32..960 pairs -> many padded shapes. The graph cache is bounded (LRU); a shape
    that was dropped is re-captured on demand and the trajectory is the one of an unbounded cache."""
    from get_b200.ddp import FlatAdam, FlatGradAllReduce, trainable_named_parameters
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep
    dev = "cuda"
    w = synthetic.get_workload("snopes", batch_claims=3, vocab=300, n_article_sources=8)
    batches = [pad_batch(synthetic.make_batch(w, seed=s), 16) for s in range(1, 7)]
    shapes = {b[K.DocContentNoPaddingEvidence].shape[0] for b in batches}
    assert len(shapes) >= 3, shapes
    tens = [synthetic.batch_to_torch(b, device=dev) for b in batches]
    order = [0, 1, 2, 3, 4, 5, 0, 1, 2, 0]

    def run(max_graphs):
        torch.manual_seed(4)
        m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev).eval()
        named = trainable_named_parameters(m)
        red = FlatGradAllReduce([p for _, p in named], names=[n for n, _ in named])
        opt = FlatAdam(red, lr=1e-3, weight_decay=1e-3)
        st = CapturedTrainStep(m, opt, red, max_graphs=max_graphs)
        losses = [float(st.step(*tens[i], batches[i]["n_real_claims"])) for i in order]
        red.detach()
        return losses, st.n_graphs(), st.evictions

    la, na, ea = run(32)
    lb, nb, eb = run(2)
    assert ea == 0 and na == len({(b["query"].shape[0], b[K.DocContentNoPaddingEvidence].shape[0], b["n_real_claims"]) for b in batches})
    assert nb <= 2 and eb >= 3
    assert np.allclose(la, lb, rtol=0, atol=1e-6), (la, lb)


@pytest.mark.gpu
def test_prefetched_inputs_give_the_same_step():
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep
    dev = "cuda"
    w = synthetic.get_workload("snopes", batch_claims=4, vocab=300, n_article_sources=8)
    batches = [pad_batch(synthetic.make_batch(w, seed=s), 32) for s in (5, 6, 7)]
    host = [synthetic.batch_to_torch(b, device="cpu", pin=True) for b in batches]
    torch.manual_seed(2)
    m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev).eval()
    stepper = CapturedTrainStep(m)
    direct = [float(stepper.step(*host[i], batches[i]["n_real_claims"])) for i in range(3)]
    nxt = stepper.prefetch(*host[0])
    got = []
    for i in range(3):
        loss = stepper.step_prefetched(nxt, batches[i]["n_real_claims"])
        if i + 1 < 3:
            nxt = stepper.prefetch(*host[i + 1])
        got.append(float(loss))
    assert direct == got


@pytest.mark.gpu
def test_token_batch_builds_the_same_inputs_on_the_device():
    """Compact (token-id) batch -> device graph construction reproduces the fitter-format tensors exactly."""
    from get_b200.step_graph import device_batch_from_tokens, token_batch_to_host
    w = synthetic.get_workload("snopes", batch_claims=6, vocab=300, n_article_sources=8)
    batch = pad_batch(synthetic.make_batch(w, seed=3), 16)
    q, d, l, kw = synthetic.batch_to_torch(batch, device="cuda")
    tq, td, tl, tkw = device_batch_from_tokens(token_batch_to_host(batch), "cuda")
    assert torch.equal(tq, q) and torch.equal(td, d) and torch.equal(tl, l)
    assert torch.equal(tkw[K.DocContentNoPaddingEvidence], kw[K.DocContentNoPaddingEvidence])
    assert torch.equal(tkw[K.Evd_Docs_Adj], kw[K.Evd_Docs_Adj].float())
    assert torch.equal(tkw[K.Query_Adj], kw[K.Query_Adj].float())
    assert torch.equal(tkw[K.Query_lens], kw[K.Query_lens])
    assert torch.equal(tkw[K.DocLensIndices][2], kw[K.DocLensIndices][2])
    assert torch.equal(tkw[K.EvidenceCountPerQuery], kw[K.EvidenceCountPerQuery])


@pytest.mark.gpu
def test_prefetched_token_batches_give_the_same_steps():
    """prefetch_tokens (side-stream token H2D + device graph construction) + step_prefetched == step on the dense batch."""
    from get_b200.ddp import FlatAdam, FlatGradAllReduce, trainable_named_parameters
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep, token_batch_to_host
    dev = "cuda"
    w = synthetic.get_workload("snopes", batch_claims=3, vocab=300, n_article_sources=8)
    batches = [pad_batch(synthetic.make_batch(w, seed=s), 16) for s in (1, 2, 3)]

    def run(tokens):
        torch.manual_seed(4)
        m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev).eval()
        named = trainable_named_parameters(m)
        red = FlatGradAllReduce([p for _, p in named], names=[n for n, _ in named])
        st = CapturedTrainStep(m, FlatAdam(red, lr=1e-3, weight_decay=1e-3), red)
        losses = []
        if tokens:
            tbs = [token_batch_to_host(b) for b in batches]
            nxt = st.prefetch_tokens(tbs[0])
            for i in range(6):
                loss = st.step_prefetched(nxt, batches[i % 3]["n_real_claims"])
                if i + 1 < 6:
                    nxt = st.prefetch_tokens(tbs[(i + 1) % 3])
                losses.append(float(loss))
        else:
            for i in range(6):
                b = batches[i % 3]
                q, d, l, kw = synthetic.batch_to_torch(b, device=dev)
                kw[K.Evd_Docs_Adj] = kw[K.Evd_Docs_Adj].float()
                kw[K.Query_Adj] = kw[K.Query_Adj].float()
                losses.append(float(st.step(q, d, l, kw, b["n_real_claims"])))
        red.detach()
        return losses

    a, b = run(False), run(True)
    assert np.allclose(a, b, rtol=0, atol=2e-6), (a, b)
