"""The oracle (oracle/get_oracle.py) against outputs recorded from the unmodified reference
(tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from helpers import load_golden, split_prefixed, to_torch_sd, named_param_values, keep_sets, import_oracle
from get_b200 import synthetic
from get_b200.keywords import KeyWordSettings as K

O = import_oracle()
TOL = 2e-6   # same ATen ops in the same order; slack only for threading-dependent summation order


def _kwargs(gin):
    batch = {k: gin[k] for k in gin}
    batch["pairs"] = int(gin[K.EvidenceCountPerQuery].sum())
    batch[K.FIXED_NUM_EVIDENCES] = gin["document"].shape[1]
    return synthetic.batch_to_torch(batch)


CASES = {
    "tiny_snopes": dict(gsl_rate=0.6, use_claim_source=False, use_article_source=True),
    "tiny_politifact": dict(gsl_rate=0.3, use_claim_source=True, use_article_source=True),
    "tiny_rate09": dict(gsl_rate=0.9, use_claim_source=False, use_article_source=False),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_model_forward_backward_matches_reference(case):
    g = load_golden(case)
    cfg = CASES[case]
    sd = to_torch_sd(split_prefixed(g, "sd/"))
    query, document, labels, kw = _kwargs(split_prefixed(g, "in/"))
    logits, parts = O.model_forward(sd, cfg, query, document, kw, return_parts=True)
    out = split_prefixed(g, "out/")
    assert np.abs(logits.numpy() - out["logits"]).max() < TOL
    assert np.abs(parts["f1"].numpy() - out["f1"]).max() < TOL
    assert np.abs(parts["score"].numpy() - out["score"]).max() < TOL
    assert keep_sets(np.sort(parts["keep_idx"].numpy(), 1)) == keep_sets(out["keep_idx"])
    assert np.abs(parts["adj_refined"].numpy() - out["adj_refined"]).max() < TOL
    assert np.abs(parts["doc_out"].numpy() - out["doc_out"]).max() < TOL
    assert np.abs(parts["word_att"].numpy() - out["word_att"]).max() < TOL
    assert np.abs(parts["evd_att"].numpy() - out["evd_att"]).max() < TOL
    # reference runtime assert: every head sums to 1 (char_man_fitter_query_repr1.py:433-434,448)
    assert np.abs(parts["word_att"].sum(1).numpy() - 1).max() < 1e-5
    loss, _, grads = O.loss_and_grads(sd, cfg, query, document, labels, kw)
    assert abs(float(loss) - float(out["loss"])) < TOL
    ggold = split_prefixed(g, "grad/")
    assert set(grads) == set(ggold)
    for n, gr in grads.items():
        assert np.abs(gr.numpy() - ggold[n]).max() < TOL * max(1.0, np.abs(ggold[n]).max()), n
    # parameters the reference never trains
    for n in out["no_grad_params"]:
        assert str(n).startswith(O.INERT_PREFIXES), n


def test_snopes_dims_matches_reference():
    g = load_golden("snopes_dims")
    out = split_prefixed(g, "out/")
    w = synthetic.get_workload("snopes", name="snopes_dims", batch_claims=3, vocab=400, n_article_sources=16,
                               evd_mean=2.5)
    seed = int(g["cfg/seed"])
    shapes = {str(n): tuple(int(s) for s in str(sh).split(",")) for n, sh in zip(out["param_names"], out["param_shapes"])}
    emb, art, clm = synthetic.make_embeddings(w, seed)
    vals = named_param_values({k: s for k, s in shapes.items() if not k.endswith("embs.weight") and k != "embedding.weight"}, seed)
    vals["embedding.weight"], vals["article_source_embs.weight"] = emb, art
    sd = to_torch_sd(vals)
    query, document, labels, kw = _kwargs(split_prefixed(g, "in/"))
    cfg = dict(gsl_rate=0.6, use_claim_source=False, use_article_source=True)
    logits, parts = O.model_forward(sd, cfg, query, document, kw, return_parts=True)
    assert np.abs(logits.numpy() - out["logits"]).max() < 1e-5
    assert keep_sets(np.sort(parts["keep_idx"].numpy(), 1)) == keep_sets(out["keep_idx"])
    assert np.abs(parts["doc_out"][:, ::17, :].numpy() - out["doc_out_rows"]).max() < 1e-5
    assert np.abs(parts["word_att"].numpy() - out["word_att"]).max() < 1e-5
    loss, _, grads = O.loss_and_grads(sd, cfg, query, document, labels, kw)
    for n, gr in grads.items():
        idx = g["gradidx/" + n]
        ref = g["gradval/" + n]
        got = gr.reshape(-1)[torch.from_numpy(idx)].numpy()
        assert np.abs(got - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), n


def test_module_surfaces_match_reference():
    g = load_golden("modules")
    adj = torch.from_numpy(g["ggnn/in/adj"])
    x = torch.from_numpy(g["ggnn/in/x"]).requires_grad_(True)
    sd = {k: v.requires_grad_(True) for k, v in to_torch_sd(split_prefixed(g, "ggnn/sd/")).items()}
    y = O.ggnn(adj, x, {"L." + k: v for k, v in sd.items()}, "L")
    assert np.abs(y.detach().numpy() - g["ggnn/out/y"]).max() < TOL
    y.backward(torch.from_numpy(g["ggnn/in/gy"]))
    assert np.abs(x.grad.numpy() - g["ggnn/out/gx"]).max() < 1e-5
    for k, v in sd.items():
        assert np.abs(v.grad.numpy() - g["ggnn/grad/" + k]).max() < 1e-5, k
    score = torch.from_numpy(g["gsl/in/score"])
    for rate in (0.3, 0.6, 0.9):
        assert np.array_equal(O.gsl(adj, score, rate).numpy(), g["gsl/out/adj_%d" % int(rate * 10)])
    bsd = {"B." + k: v for k, v in to_torch_sd(split_prefixed(g, "block/sd/")).items()}
    yb = O.ggnn_with_gsl(adj, torch.from_numpy(g["ggnn/in/x"]), bsd, "B", 0.6)
    assert np.abs(yb.numpy() - g["block/out/y"]).max() < TOL
    asd = to_torch_sd(split_prefixed(g, "att/sd/"))
    o, a = O.concat_not_equal_self_att(torch.from_numpy(g["att/in/left"]), torch.from_numpy(g["att/in/right"]),
                                       torch.from_numpy(g["att/in/mask"]), asd["linear1.weight"], asd["linear2.weight"])
    assert np.abs(o.numpy() - g["att/out/attended"]).max() < TOL
    assert np.abs(a.numpy() - g["att/out/att"]).max() < TOL
    esd = to_torch_sd(split_prefixed(g, "ext/sd/"))
    eo, ea = O.multi_head_self_att_extend(torch.from_numpy(g["att/in/right"]), torch.from_numpy(g["att/in/mask"]),
                                          esd["linear1.weight"], esd["linear2.weight"], return_att_weights=True)
    assert np.abs(eo.numpy() - g["ext/out/attended"]).max() < TOL
    assert np.abs(ea.numpy() - g["ext/out/att"]).max() < TOL


def test_word_graph_matches_reference_convert_text():
    g = load_golden("graphs")
    i = 0
    while "g%d/cfg" % i in g:
        R, w, n_nodes = (int(v) for v in g["g%d/cfg" % i])
        nodes, adj, nn = synthetic.word_graph(g["g%d/tokens" % i], R, w)
        assert nn == n_nodes
        assert np.array_equal(nodes, g["g%d/nodes" % i])
        assert np.array_equal(adj, g["g%d/adj" % i])
        i += 1
    assert i == 5


def test_oracle_matches_reference_on_a_real_snopes_batch():
    """Real TSV rows (formatted_data/declare/Snopes, 6 claims / 53 evidences, one claim with 28 evidences, hub nodes with up to
    38 neighbours): the graph restatement equals the reference's convert_text, and the oracle equals the reference model."""
    from helpers import named_param_values, real_batch_from_golden
    from get_b200 import synthetic
    from get_b200.keywords import KeyWordSettings as K
    from get_b200.model import Graph_basedSemantiStructure
    O = import_oracle()
    gold = load_golden("real_snopes")
    w, batch = real_batch_from_golden(gold)
    assert int(batch[K.EvidenceCountPerQuery].max()) >= 20 and int((batch[K.Evd_Docs_Adj] != 0).sum(-1).max()) >= 30
    seed = int(gold["cfg/seed"])
    model = Graph_basedSemantiStructure(synthetic.match_params(w, seed=seed, cuda=False))
    sd = model.state_dict()
    vals = named_param_values({k: tuple(v.shape) for k, v in sd.items() if not k.endswith("embs.weight") and k != "embedding.weight"}, seed)
    sd.update({k: torch.from_numpy(v) for k, v in vals.items()})
    q, d, l, kw = synthetic.batch_to_torch(batch)
    cfg = dict(gsl_rate=w.gsl_rate, use_claim_source=w.use_claim_source, use_article_source=w.use_article_source)
    loss, logits, grads = O.loss_and_grads(sd, cfg, q, d, l, kw)
    _, parts = O.model_forward(sd, cfg, q, d, kw, return_parts=True)
    assert np.allclose(logits.numpy(), gold["out/logits"], atol=2e-6)
    assert abs(float(loss) - float(gold["out/loss"])) < 2e-6
    assert np.allclose(parts["word_att"].numpy(), gold["out/word_att"], atol=2e-6)
    assert np.allclose(parts["evd_att"].numpy(), gold["out/evd_att"], atol=2e-6)
    assert np.array_equal(np.sort(parts["keep_idx"].numpy(), axis=1), gold["out/keep_idx"])
    for n, g in grads.items():
        flat = g.reshape(-1)
        assert np.allclose(flat[torch.from_numpy(gold["gradidx/" + n])].numpy(), gold["gradval/" + n], atol=2e-6), n


def test_neighbor_list_restatement_equals_dense_products_on_reference_graphs():
    """The CSR record format of the list kernels, restated in numpy, against the reference's own graphs (convert_text +
    _laplacian_normalize outputs in graphs.npz): adj @ x, adj^T @ x and the GSL-masked product are reproduced exactly."""
    oracle = import_oracle()
    d = load_golden("graphs")
    rng = np.random.default_rng(0)
    names = sorted({k.split("/")[0] for k in d.keys()})
    assert names
    for g in names:
        adj = d[g + "/adj"].astype(np.float64)
        n = adj.shape[0]
        x = rng.standard_normal((n, 12))
        keep = rng.random(n) < 0.6
        for tr in (False, True):
            rowptr, idx, w, used = oracle.neighbor_lists(adj, transpose=tr)
            a = adj.T if tr else adj
            assert rowptr[-1] == (a != 0).sum() and used == (np.nonzero(a)[1].max() + 1 if (a != 0).any() else 0)
            assert all((np.diff(idx[rowptr[i]:rowptr[i + 1]]) > 0).all() for i in range(n))
            assert np.allclose(oracle.aggregate_lists(rowptr, idx, w, x), a @ x, rtol=0, atol=1e-12)
            masked = a * (keep[:, None] | keep[None, :])
            assert np.allclose(oracle.aggregate_lists(rowptr, idx, w, x, keep), masked @ x, rtol=0, atol=1e-12)
        assert np.allclose(adj, adj.T)          # GET graphs are symmetric: the builder writes both records in one pass
