"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own for this path, so the
oracle is pinned against these recorded outputs of the reference modules themselves.

npz layout: 'sd/<param>' reference state_dict, 'in/<name>' inputs, 'out/<name>' outputs,
'grad/<param>' autograd gradients (eval mode => dropout off), 'cfg/<key>' scalars.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _ref_import import import_reference  # noqa: E402
from helpers import named_param_values  # noqa: E402
from get_b200 import synthetic  # noqa: E402
from get_b200.keywords import KeyWordSettings as K  # noqa: E402

gbss, wrapper, tba, sa = import_reference()
torch.set_num_threads(4)


def build_reference_model(w, seed):
    params = synthetic.match_params(w, seed=seed, cuda=False)
    model = gbss.Graph_basedSemantiStructure(params)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    vals = named_param_values({k: s for k, s in shapes.items() if not k.endswith("embs.weight") and k != "embedding.weight"},
                              seed)
    sd = model.state_dict()
    for k, v in vals.items():
        sd[k] = torch.from_numpy(v)
    model.load_state_dict(sd)
    model.eval()
    return model, params


def run_model_case(name, w, seed, store_sd=True, sample_only=False):
    model, params = build_reference_model(w, seed)
    batch = synthetic.make_batch(w, seed=seed)
    query, document, labels, kw = synthetic.batch_to_torch(batch)
    kw[K.OutputRankingKey] = True
    cap = {}
    blk = model.ggnn_with_gsl
    hooks = [
        blk.feat_prop1.register_forward_hook(lambda m, i, o: cap.__setitem__("f1", o.detach())),
        blk.word_scorer1.register_forward_hook(lambda m, i, o: cap.__setitem__("score", o.detach())),
        blk.gsl1.register_forward_hook(lambda m, i, o: cap.__setitem__("adj_refined", o.detach())),
        blk.feat_prop2.register_forward_hook(lambda m, i, o: cap.__setitem__("doc_out", o.detach())),
        model.ggnn4claim_1.register_forward_hook(lambda m, i, o: cap.__setitem__("claim_hidden", o.detach())),
    ]
    logits, (word_att, evd_att) = model(query, document, **kw)
    loss = torch.nn.CrossEntropyLoss()(logits, labels.long())          # losses.py:29-32
    loss.backward()
    for h in hooks:
        h.remove()
    k = int(w.gsl_rate * w.len_right)
    keep_idx = cap["score"].topk(k, 1)[1].squeeze(-1)
    out = {}
    for key in ("query", "document", "labels", "e_lens"):
        out["in/" + key] = batch[key]
    for key in (K.Query_lens, K.Doc_lens, K.Query_Adj, K.Evd_Docs_Adj, K.DocContentNoPaddingEvidence,
                K.EvidenceCountPerQuery, K.QuerySources, K.DocSources):
        out["in/" + key] = batch[key]
    out["cfg/seed"] = np.int64(seed)
    out["cfg/workload"] = np.array(w.name)
    out["out/logits"] = logits.detach().numpy()
    out["out/loss"] = loss.detach().numpy()
    out["out/word_att"] = word_att.detach().numpy()
    out["out/evd_att"] = evd_att.detach().numpy()
    out["out/score"] = cap["score"].numpy()
    out["out/keep_idx"] = np.sort(keep_idx.numpy(), axis=1)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    no_grad = sorted(n for n, p in model.named_parameters() if p.requires_grad and p.grad is None)
    out["out/no_grad_params"] = np.array(no_grad)
    out["out/param_names"] = np.array(list(model.state_dict().keys()))
    out["out/param_shapes"] = np.array([",".join(str(int(s)) for s in v.shape) for v in model.state_dict().values()])
    if sample_only:
        # big dims: parameters are regenerated from (name, seed) by tests/helpers.named_param_values and
        # the embedding tables by synthetic.match_params; store only sampled / reduced outputs
        out["out/f1_rows"] = cap["f1"][:, ::17, :].numpy()
        out["out/doc_out_rows"] = cap["doc_out"][:, ::17, :].numpy()
        out["out/claim_hidden_rows"] = cap["claim_hidden"][:, ::7, :].numpy()
        for n, g in grads.items():
            g = g.detach()
            out["gradsum/" + n] = np.array([g.double().sum().item(), g.double().abs().sum().item()])
            flat = g.reshape(-1)
            idx = np.random.default_rng([seed, zlib.crc32(n.encode())]).integers(0, flat.numel(), size=min(64, flat.numel()))
            out["gradidx/" + n] = idx
            out["gradval/" + n] = flat[torch.from_numpy(idx)].numpy()
    else:
        out["out/f1"] = cap["f1"].numpy()
        out["out/doc_out"] = cap["doc_out"].numpy()
        out["out/adj_refined"] = cap["adj_refined"].numpy()
        out["out/claim_hidden"] = cap["claim_hidden"].numpy()
        for n, g in grads.items():
            out["grad/" + n] = g.detach().numpy()
    if store_sd:
        for n, v in model.state_dict().items():
            out["sd/" + n] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote %s: pairs=%d loss=%.6f logits[0]=%s no_grad=%d params" % (
        name, batch["pairs"], float(loss), logits[0].detach().numpy(), len(no_grad)))


def run_module_cases():
    """Op-level surfaces on inputs the model never produces: non-symmetric dense
    adjacency, all heads, masks with holes, the `left`-less variant of self_attention.py."""
    rng = np.random.default_rng(20240521)
    out = {}
    G, N, Din, H = 5, 11, 9, 13
    adj = (rng.random((G, N, N)) < 0.35) * rng.uniform(0.1, 1.0, (G, N, N))
    adj[:, 8:, :] = 0
    adj[:, :, 8:] = 0
    adj = adj.astype(np.float32)
    x = rng.standard_normal((G, N, Din)).astype(np.float32)
    layer = wrapper.GGNN(Din, H, dropout=0.2).eval()
    sd = named_param_values({k: v.shape for k, v in layer.state_dict().items()}, 7)
    layer.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    xt = torch.from_numpy(x).requires_grad_(True)
    y = layer(torch.from_numpy(adj), xt)
    gy = torch.from_numpy(rng.standard_normal(y.shape).astype(np.float32))
    y.backward(gy)
    out.update({"ggnn/in/adj": adj, "ggnn/in/x": x, "ggnn/in/gy": gy.numpy(), "ggnn/out/y": y.detach().numpy(),
                "ggnn/out/gx": xt.grad.numpy()})
    for k, v in sd.items():
        out["ggnn/sd/" + k] = v
    for n, p in layer.named_parameters():
        out["ggnn/grad/" + n] = p.grad.numpy()
    # GSL alone, incl. rates where int(rate*N) truncates
    score = rng.standard_normal((G, N, 1)).astype(np.float32)
    for rate in (0.3, 0.6, 0.9):
        out["gsl/out/adj_%d" % int(rate * 10)] = wrapper.GSL(rate)(torch.from_numpy(adj), torch.from_numpy(score)).numpy()
    out["gsl/in/score"] = score
    # whole block
    blk = wrapper.GGNN_with_GSL(Din, H, H, rate=0.6, dropout=0.2).eval()
    bsd = named_param_values({k: v.shape for k, v in blk.state_dict().items()}, 8)
    blk.load_state_dict({k: torch.from_numpy(v) for k, v in bsd.items()})
    xb = torch.from_numpy(x).requires_grad_(True)
    yb = blk(torch.from_numpy(adj), xb)
    yb.backward(gy)
    out.update({"block/out/y": yb.detach().numpy(), "block/out/gx": xb.grad.numpy()})
    for k, v in bsd.items():
        out["block/sd/" + k] = v
    for n, p in blk.named_parameters():
        if p.grad is not None:
            out["block/grad/" + n] = p.grad.numpy()
    # attention, both flavours
    P, X, Dr, C = 10, 6, 7, 3
    att = tba.ConcatNotEqualSelfAtt(X + Dr, H, C)
    asd = named_param_values({k: v.shape for k, v in att.state_dict().items()}, 9)
    att.load_state_dict({k: torch.from_numpy(v) for k, v in asd.items()})
    left = torch.from_numpy(rng.standard_normal((G, X)).astype(np.float32)).requires_grad_(True)
    right = torch.from_numpy(rng.standard_normal((G, P, Dr)).astype(np.float32)).requires_grad_(True)
    mask = (rng.random((G, P)) < 0.7).astype(np.int64)
    mask[:, 0] = 1
    o, a = att(left, right, torch.from_numpy(mask))
    go = torch.from_numpy(rng.standard_normal(o.shape).astype(np.float32))
    ga = torch.from_numpy(rng.standard_normal(a.shape).astype(np.float32))
    (o * go).sum().backward(retain_graph=True)
    out.update({"att/in/left": left.detach().numpy(), "att/in/right": right.detach().numpy(), "att/in/mask": mask,
                "att/in/go": go.numpy(), "att/out/attended": o.detach().numpy(), "att/out/att": a.detach().numpy(),
                "att/out/gleft": left.grad.numpy(), "att/out/gright": right.grad.numpy()})
    for k, v in asd.items():
        out["att/sd/" + k] = v
    for n, p in att.named_parameters():
        out["att/grad/" + n] = p.grad.numpy()
    ext = sa.MultiHeadSelfAttentionICLR2017Extend(Dr, H, C)
    esd = named_param_values({k: v.shape for k, v in ext.state_dict().items()}, 10)
    ext.load_state_dict({k: torch.from_numpy(v) for k, v in esd.items()})
    eo, ea = ext(right.detach(), torch.from_numpy(mask), return_att_weights=True)
    out.update({"ext/out/attended": eo.detach().numpy(), "ext/out/att": ea.detach().numpy()})
    for k, v in esd.items():
        out["ext/sd/" + k] = v
    np.savez_compressed(os.path.join(HERE, "modules.npz"), **out)
    print("wrote modules")


def run_graph_cases():
    """Adjacency construction: reference `convert_text` (interactions.py:334-351) vs tokens. scipy removed
    `.A` (interactions.py:18), so `_laplacian_normalize` is re-bound to the same expression with `.toarray()`."""
    import scipy.sparse as sp
    import interactions as I

    def lap(adj):
        adj = sp.coo_matrix(adj)
        rowsum = np.array(adj.sum(1))
        with np.errstate(divide="ignore"):
            d_inv_sqrt = np.power(rowsum, -0.5).flatten()
        d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.
        d = sp.diags(d_inv_sqrt)
        return (adj.dot(d).transpose().dot(d)).toarray()

    I._laplacian_normalize = lap
    rng = np.random.default_rng(5)
    out = {}
    for i, (R, w) in enumerate([(30, 3), (100, 3), (100, 5), (100, 9), (16, 2)]):
        length = int(rng.integers(R // 2, R + 1))
        toks = [int(t) for t in rng.integers(2, 40, size=R)]
        nodes, adj, n_nodes = I.ClassificationInteractions.convert_text(None, toks, R, length, w)
        out["g%d/tokens" % i] = np.array(toks[:length])
        out["g%d/cfg" % i] = np.array([R, w, n_nodes])
        out["g%d/nodes" % i] = np.array(nodes, dtype=np.int64)
        out["g%d/adj" % i] = adj
    np.savez_compressed(os.path.join(HERE, "graphs.npz"), **out)
    print("wrote graphs")


def run_real_case(name="real_snopes", n_claims=6, D=100, H=100, seed=15):
    """One mini-batch built from REAL rows of formatted_data/declare/Snopes (5fold/train_0.tsv) the way the reference's data
    layer does it: whitespace tokens of the mapped text, the last 30 / 100 tokens (FixedLength truncate_mode='pre' keeps
    the tail, matchzoo/preprocessors/units/fixed_length.py:24-30,63-64), word graphs by the reference's own
    `ClassificationInteractions.convert_text` (interactions.py:334-351), article sources as ids, padding conventions of
    handlers/mz_sampler.py:127-160. Real text brings what the synthetic generator does not: hub rows with dozens of
    neighbours, the data's real pad fractions, a claim with >20 evidences. Stored: raw token ids (the test rebuilds the
    batch and checks the graph restatement against the stored reference adjacency triplets) and the outputs / sampled
    gradients of the unmodified reference model at reduced feature sizes (D = H = 100, heads 5/2)."""
    import csv
    import scipy.sparse as sp
    import interactions as I

    def lap(adj):
        adj = sp.coo_matrix(adj)
        rowsum = np.array(adj.sum(1))
        with np.errstate(divide="ignore"):
            d_inv_sqrt = np.power(rowsum, -0.5).flatten()
        d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.
        d = sp.diags(d_inv_sqrt)
        return (adj.dot(d).transpose().dot(d)).toarray()
    I._laplacian_normalize = lap
    L, R, n = 30, 100, 30
    path = os.path.join(_ref_root(), "formatted_data", "declare", "Snopes", "mapped_data", "5fold", "train_0.tsv")
    claims = {}
    with open(path, newline="") as fh:
        for row in csv.DictReader(fh, delimiter="\t", quoting=csv.QUOTE_NONE):
            c = claims.setdefault(row["id_left"], {"text": row["claim_text"], "label": row["cred_label"], "evd": []})
            c["evd"].append((row["evidence"], row["evidence_source"]))
    # the claim with the most evidences, one with a single evidence, and the next few in file order
    by_cnt = sorted(claims.items(), key=lambda kv: -len(kv[1]["evd"]))
    chosen = [by_cnt[0][0], by_cnt[-1][0]] + [k for k in list(claims)[:40] if k not in (by_cnt[0][0], by_cnt[-1][0])][:n_claims - 2]
    vocab, sources = {"<PAD>": 0, "<OOV>": 1}, {}

    def ids(text, keep):
        toks = text.split()[-keep:]
        return [vocab.setdefault(t, len(vocab)) for t in toks]
    B = len(chosen)
    cnt = np.array([min(n, len(claims[c]["evd"])) for c in chosen], np.int64)
    B1 = int(cnt.sum())
    batch = {"query": np.zeros((B, L), np.int64), K.Query_Adj: np.zeros((B, L, L)), K.Query_lens: np.zeros((B,), np.int64),
             "document": np.zeros((B, n, R), np.int64), K.Doc_lens: np.zeros((B, n), np.int64),
             K.DocSources: np.full((B, n), -1, np.int64), K.DocContentNoPaddingEvidence: np.zeros((B1, R), np.int64),
             K.Evd_Docs_Adj: np.zeros((B1, R, R)), "e_lens": np.zeros((B1,), np.int64),
             "raw_query_tokens": np.zeros((B, L), np.int64), "raw_query_lens": np.zeros((B,), np.int32),
             "raw_doc_tokens": np.zeros((B1, R), np.int64), "raw_doc_lens": np.zeros((B1,), np.int32),
             K.EvidenceCountPerQuery: cnt, K.FIXED_NUM_EVIDENCES: n, "window": 3,
             K.QuerySources: np.zeros((B, 1), np.int64), "labels": np.zeros((B,), np.int64)}
    g = 0
    for b, cid in enumerate(chosen):
        c = claims[cid]
        toks = ids(c["text"], L)
        nodes, adj, nn = I.ClassificationInteractions.convert_text(None, toks + [0] * (L - len(toks)), L, len(toks), 3)
        batch["query"][b], batch[K.Query_Adj][b], batch[K.Query_lens][b] = nodes, adj, nn
        batch["raw_query_tokens"][b, :len(toks)], batch["raw_query_lens"][b] = toks, len(toks)
        batch["labels"][b] = 1 if c["label"].strip().lower() == "true" else 0
        for j, (text, src) in enumerate(c["evd"][:n]):
            toks = ids(text, R)
            nodes, adj, nn = I.ClassificationInteractions.convert_text(None, toks + [0] * (R - len(toks)), R, len(toks), 3)
            batch["document"][b, j], batch[K.Doc_lens][b, j] = nodes, nn
            batch[K.DocSources][b, j] = sources.setdefault(src, len(sources))
            batch[K.DocContentNoPaddingEvidence][g], batch[K.Evd_Docs_Adj][g], batch["e_lens"][g] = nodes, adj, nn
            batch["raw_doc_tokens"][g, :len(toks)], batch["raw_doc_lens"][g] = toks, len(toks)
            g += 1
    batch["pairs"] = B1
    w = synthetic.get_workload("snopes", name=name, batch_claims=B, vocab=len(vocab), emb_dim=D, hidden=H,
                               n_article_sources=max(8, len(sources)))
    model, params = build_reference_model(w, seed)
    query, document, labels, kw = synthetic.batch_to_torch(batch)
    kw[K.OutputRankingKey] = True
    cap = {}
    h = model.ggnn_with_gsl.word_scorer1.register_forward_hook(lambda m, i, o: cap.__setitem__("score", o.detach()))
    logits, (word_att, evd_att) = model(query, document, **kw)
    h.remove()
    loss = torch.nn.CrossEntropyLoss()(logits, labels.long())
    loss.backward()
    out = {"cfg/seed": np.int64(seed), "cfg/dims": np.array([B, L, R, n, D, H, len(vocab), w.n_article_sources]),
           "in/raw_query_tokens": batch["raw_query_tokens"], "in/raw_query_lens": batch["raw_query_lens"],
           "in/raw_doc_tokens": batch["raw_doc_tokens"], "in/raw_doc_lens": batch["raw_doc_lens"],
           "in/" + K.EvidenceCountPerQuery: cnt, "in/" + K.DocSources: batch[K.DocSources], "in/labels": batch["labels"],
           "out/logits": logits.detach().numpy(), "out/loss": loss.detach().numpy(), "out/word_att": word_att.detach().numpy(),
           "out/evd_att": evd_att.detach().numpy(), "out/score": cap["score"].numpy(),
           "out/keep_idx": np.sort(cap["score"].topk(int(w.gsl_rate * R), 1)[1].squeeze(-1).numpy(), axis=1)}
    # the reference adjacencies as sparse triplets (graph, row, col, value)
    gi, ri, ci = np.nonzero(batch[K.Evd_Docs_Adj])
    out["in/adj_idx"] = np.stack([gi, ri, ci]).astype(np.int16)
    out["in/adj_val"] = batch[K.Evd_Docs_Adj][gi, ri, ci]
    deg = (batch[K.Evd_Docs_Adj] != 0).sum(-1)
    for nme, gr in ((n_, p.grad) for n_, p in model.named_parameters() if p.grad is not None):
        gr = gr.detach()
        out["gradsum/" + nme] = np.array([gr.double().sum().item(), gr.double().abs().sum().item()])
        flat = gr.reshape(-1)
        idx = np.random.default_rng([seed, zlib.crc32(nme.encode())]).integers(0, flat.numel(), size=min(64, flat.numel()))
        out["gradidx/" + nme] = idx
        out["gradval/" + nme] = flat[torch.from_numpy(idx)].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote %s: claims=%d pairs=%d evidences/claim=%s max neighbours per node=%d real nodes/graph=%.1f loss=%.6f" % (
        name, B, B1, cnt.tolist(), int(deg.max()), float(batch["e_lens"].mean()), float(loss)))


def _ref_root():
    from _ref_import import REF_ROOT
    return "/root/reference" if os.path.isdir("/root/reference/formatted_data") else REF_ROOT


if __name__ == "__main__":
    if "--real-only" in sys.argv:
        run_real_case()
        sys.exit(0)
    run_graph_cases()
    run_module_cases()
    run_model_case("tiny_snopes", synthetic.get_workload("tiny"), seed=11)
    run_model_case("tiny_politifact", synthetic.get_workload("tiny", name="tiny_politifact", use_claim_source=True,
                                                             heads_words=2, heads_evds=1, window=5, gsl_rate=0.3,
                                                             batch_claims=4), seed=12)
    run_model_case("tiny_rate09", synthetic.get_workload("tiny", name="tiny_rate09", len_right=16, window=9,
                                                         gsl_rate=0.9, use_article_source=False), seed=13)
    run_model_case("snopes_dims", synthetic.get_workload("snopes", name="snopes_dims", batch_claims=3, vocab=400,
                                                         n_article_sources=16, evd_mean=2.5),
                   seed=14, store_sd=False, sample_only=True)
    run_real_case()
