"""Seeded synthetic mini-batches in the reference fitter's calling convention.

Host-side (numpy) restatement of the input construction the hot path consumes:

* word graph of one text  -> `ClassificationInteractions.convert_text` (reference `interactions.py:334-351`)
  + `_laplacian_normalize` (`interactions.py:11-18`): nodes = de-duplicated tokens in first-occurrence
  order, edge (u,v) iff the words co-occur within `window-1` positions (self loops included),
  then D^-1/2 A D^-1/2, dense float64, zero rows/cols for pad nodes;
* padding conventions     -> `handlers/mz_sampler.py:127-160` (token id 0, source id -1, length 0);
* B -> B1 flattening      -> `Fitting/FittingFC/char_man_fitter_query_repr1.py:207-250`.

Shapes follow BASELINE.json `configs` / SURVEY.md section 8(d).
"""
from dataclasses import dataclass, replace
from typing import Dict, Optional

import numpy as np

from .keywords import KeyWordSettings as K


@dataclass(frozen=True)
class Workload:
    name: str
    batch_claims: int = 32          # B
    len_left: int = 30              # L (claim node slots)
    len_right: int = 100            # R (evidence node slots)
    emb_dim: int = 300              # D
    hidden: int = 300               # H
    heads_words: int = 5
    heads_evds: int = 2
    fixed_num_evidences: int = 30   # n, asserted by the reference (gbss.py:93)
    vocab: int = 5000
    window: int = 3
    gsl_rate: float = 0.6
    use_claim_source: bool = False
    use_article_source: bool = True
    n_claim_sources: int = 8
    n_article_sources: int = 512
    src_dim: int = 128
    evd_mean: float = 6.74          # mean real evidences per claim; <=0 means always `fixed_num_evidences`
    evd_pool: int = 140             # words in the per-document pool (=> ~71 unique nodes of 100)
    num_classes: int = 2
    true_rate: float = 0.27


WORKLOADS: Dict[str, Workload] = {
    # BASELINE.json configs[0]/[1]: run_snopes.sh hyper-parameters
    "snopes": Workload(name="snopes"),
    # configs[2]: run_politifact.sh (claim source on, heads 3/1, 8.17 evidences per claim)
    "politifact": Workload(name="politifact", heads_words=3, heads_evds=1, use_claim_source=True,
                           n_claim_sources=544, n_article_sources=3605, evd_mean=8.17, true_rate=0.5),
    # configs[3]: B=512 claims x 30 evidences, R=200, D=H=512, 8 word heads
    "synthetic512": Workload(name="synthetic512", batch_claims=512, len_right=200, emb_dim=512, hidden=512,
                             heads_words=8, heads_evds=2, evd_mean=0.0, evd_pool=280),
    # configs[4]: stream of claims x 30 evidences at Snopes dims (batch_claims = claims per step)
    "stream": Workload(name="stream", batch_claims=512, evd_mean=0.0),
    # tiny shape for CPU tests / smoke
    "tiny": Workload(name="tiny", batch_claims=3, len_left=6, len_right=12, emb_dim=20, hidden=24,
                     heads_words=3, heads_evds=2, vocab=60, n_article_sources=7, n_claim_sources=4,
                     src_dim=8, evd_mean=2.0, evd_pool=14),
}


def get_workload(base: str, **overrides) -> Workload:
    return replace(WORKLOADS[base], **overrides)


def word_graph(tokens: np.ndarray, fixed_length: int, window: int):
    """One text -> (node token ids (fixed_length,), normalised adjacency (fixed_length, fixed_length) f64, n_nodes).

    Restates `convert_text` (interactions.py:334-351): `tokens` are the first `length` real tokens."""
    tokens = np.asarray(tokens, dtype=np.int64)
    length = int(tokens.shape[0])
    assert 0 < length <= fixed_length
    uniq, first_pos, inverse = np.unique(tokens, return_index=True, return_inverse=True)
    order = np.argsort(first_pos, kind="stable")          # unique words in first-occurrence order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    node_of_pos = rank[inverse]                           # node id of every position
    n_nodes = int(uniq.shape[0])
    adj = np.zeros((fixed_length, fixed_length), dtype=np.float64)
    for d in range(-(window - 1), window):               # j in [i-w+1, i+w-1], clipped to the text
        lo, hi = max(0, -d), min(length, length - d)
        if hi > lo:
            adj[node_of_pos[lo:hi], node_of_pos[lo + d:hi + d]] = 1.0
    # symmetric normalisation D^-1/2 A D^-1/2 with 0 for isolated (pad) nodes (interactions.py:11-18)
    deg = adj.sum(axis=1)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5)
    dis[np.isinf(dis)] = 0.0
    adj = (adj * dis[None, :]).T * dis[None, :]
    nodes = np.zeros((fixed_length,), dtype=np.int64)
    nodes[:n_nodes] = uniq[order]
    return nodes, adj, n_nodes


def make_embeddings(w: Workload, seed: int = 123756):
    """Frozen word table U(-0.2,0.2) (pad/OOV rows are non-zero in the reference too)
    and trainable source tables."""
    rng = np.random.default_rng(seed + 17)
    emb = rng.uniform(-0.2, 0.2, size=(w.vocab, w.emb_dim)).astype(np.float32)
    art = rng.uniform(-0.2, 0.2, size=(w.n_article_sources, w.src_dim)).astype(np.float32)
    clm = rng.uniform(-0.2, 0.2, size=(w.n_claim_sources, w.src_dim)).astype(np.float32)
    return emb, art, clm


def match_params(w: Workload, seed: int = 123756, cuda: bool = False, dropout_gnn: float = 0.2) -> dict:
    """The constructor dict of `Graph_basedSemantiStructure` (master_get.py:118-144)."""
    emb, art, clm = make_embeddings(w, seed)
    return {
        "embedding": emb, "embedding_freeze": True, "num_classes": w.num_classes,
        "fixed_length_left": w.len_left, "fixed_length_right": w.len_right,
        "use_claim_source": bool(w.use_claim_source), "claim_source_embeddings": clm,
        "use_article_source": bool(w.use_article_source), "article_source_embeddings": art,
        "cuda": bool(cuda), "num_att_heads_for_words": w.heads_words, "num_att_heads_for_evds": w.heads_evds,
        "dropout_gnn": dropout_gnn, "dropout_left": 0.2, "dropout_right": 0.2,
        "hidden_size": w.hidden, "gsl_rate": w.gsl_rate, "output_size": w.num_classes,
    }


def make_batch(w: Workload, seed: int = 123756, n_claims: Optional[int] = None, adj_dtype=np.float64) -> dict:
    """One mini-batch as numpy arrays, already flattened B -> B1 the way the fitter does it.

    Keys: query (B,L) i64, document (B,n,R) i64, labels (B,) i64 and the kwargs of the boundary
     under their `KeyWordSettings` names, plus `e_lens` (B1,) and `pairs` = B1."""
    rng = np.random.default_rng(seed)
    B = int(n_claims if n_claims is not None else w.batch_claims)
    n, L, R = w.fixed_num_evidences, w.len_left, w.len_right
    if w.evd_mean > 0:
        cnt = np.clip(rng.geometric(1.0 / w.evd_mean, size=B), 1, n).astype(np.int64)
    else:
        cnt = np.full((B,), n, dtype=np.int64)
    B1 = int(cnt.sum())
    query = np.zeros((B, L), np.int64)
    query_adj = np.zeros((B, L, L), adj_dtype)
    query_lens = np.zeros((B,), np.int64)
    document = np.zeros((B, n, R), np.int64)
    docs_lens = np.zeros((B, n), np.int64)
    doc_sources = np.full((B, n), -1, np.int64)
    flat_doc = np.zeros((B1, R), np.int64)
    flat_adj = np.zeros((B1, R, R), adj_dtype)
    e_lens = np.zeros((B1,), np.int64)
    # the raw (not de-duplicated) token sequences the graphs are built from: input of the device-side graph construction
    raw_q = np.zeros((B, L), np.int64)
    raw_q_len = np.zeros((B,), np.int32)
    raw_d = np.zeros((B1, R), np.int64)
    raw_d_len = np.zeros((B1,), np.int32)
    g = 0
    for c in range(B):
        qlen = int(np.clip(rng.poisson(9.6), 3, L))
        toks = rng.integers(2, w.vocab, size=qlen)
        nodes, adj, nn = word_graph(toks, L, w.window)
        query[c], query_adj[c], query_lens[c] = nodes, adj, nn
        raw_q[c, :qlen], raw_q_len[c] = toks, qlen
        for j in range(int(cnt[c])):
            pool = rng.integers(2, w.vocab, size=min(w.evd_pool, w.vocab - 2))
            dlen = R if rng.random() < 0.9 else int(rng.integers(max(2, R // 4), R + 1))
            toks = pool[rng.integers(0, pool.shape[0], size=dlen)]
            nodes, adj, nn = word_graph(toks, R, w.window)
            document[c, j], docs_lens[c, j] = nodes, nn
            doc_sources[c, j] = rng.integers(0, w.n_article_sources)
            flat_doc[g], flat_adj[g], e_lens[g] = nodes, adj, nn
            raw_d[g, :dlen], raw_d_len[g] = toks, dlen
            g += 1
    labels = (rng.random(B) < w.true_rate).astype(np.int64)
    query_sources = rng.integers(0, w.n_claim_sources, size=(B, 1)).astype(np.int64)
    return {
        "query": query, "document": document, "labels": labels, "e_lens": e_lens, "pairs": B1,
        "raw_query_tokens": raw_q, "raw_query_lens": raw_q_len, "raw_doc_tokens": raw_d, "raw_doc_lens": raw_d_len,
        "window": w.window,
        K.Query_lens: query_lens, K.Doc_lens: docs_lens, K.Query_Adj: query_adj,
        K.Evd_Docs_Adj: flat_adj, K.DocContentNoPaddingEvidence: flat_doc,
        K.EvidenceCountPerQuery: cnt, K.FIXED_NUM_EVIDENCES: n,
        K.QuerySources: query_sources, K.DocSources: doc_sources,
    }


def batch_to_torch(batch: dict, device="cpu", adj_dtype=None, pin: bool = False):
    """numpy batch -> (query, document, labels, kwargs) torch tensors with the fitter's kwargs layout
    (char_man_fitter_query_repr1.py:234-250)."""
    import torch

    def t(x, dtype=None):
        y = torch.from_numpy(np.ascontiguousarray(x))
        if dtype is not None:
            y = y.to(dtype)
        if pin and device == "cpu" and torch.cuda.is_available():
            y = y.pin_memory()
        return y.to(device) if device != "cpu" else y

    e_lens = t(batch["e_lens"])
    kw = {
        K.Query_lens: t(batch[K.Query_lens]),
        K.Doc_lens: batch[K.Doc_lens],
        K.DocLensIndices: (None, None, e_lens),
        K.QueryLensIndices: (None, None, t(batch[K.Query_lens])),
        K.QuerySources: t(batch[K.QuerySources]),
        K.DocSources: t(batch[K.DocSources]),
        K.DocContentNoPaddingEvidence: t(batch[K.DocContentNoPaddingEvidence]),
        K.EvidenceCountPerQuery: t(batch[K.EvidenceCountPerQuery]),
        K.FIXED_NUM_EVIDENCES: int(batch[K.FIXED_NUM_EVIDENCES]),
        K.Query_Adj: t(batch[K.Query_Adj], adj_dtype),
        K.Evd_Docs_Adj: t(batch[K.Evd_Docs_Adj], adj_dtype),
    }
    return t(batch["query"]), t(batch["document"]), t(batch["labels"]), kw
