"""Shared test helpers: deterministic parameters, golden loading, oracle import (tests only)."""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def named_param_values(shapes: dict, seed: int, scale: float = 1.0) -> dict:
    """Portable deterministic parameter values: one numpy PCG64 stream per parameter NAME, so values do not
    depend on dict order, torch version or device. Weights ~ U(-a,a) with a = scale*sqrt(3/fan_in)
    (variance 1/fan_in keeps activations O(1) like the reference's Kaiming/Xavier inits); biases U(-.1,.1)."""
    out = {}
    for name, shape in shapes.items():
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        shape = tuple(int(s) for s in shape)
        if len(shape) >= 2:
            a = scale * np.sqrt(3.0 / shape[-1])
            v = rng.uniform(-a, a, size=shape)
        else:
            v = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = v.astype(np.float32)
    return out


def load_golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def split_prefixed(d: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def to_torch_sd(npd: dict, dtype=torch.float32, device="cpu") -> dict:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype if v.dtype.kind == "f" else None).to(device)
            for k, v in npd.items()}


def import_oracle():
    from oracle import get_oracle
    return get_oracle


def keep_sets(idx) -> list:
    return [frozenset(int(i) for i in row) for row in np.asarray(idx)]


def real_batch_from_golden(gold: dict):
    """Rebuild the real-data mini-batch of tests/golden/real_snopes.npz from its raw token ids with the repo's restatement
    of the reference's graph construction (get_b200.synthetic.word_graph), and check it against the stored triplets of the
    adjacencies the REFERENCE's convert_text produced. Returns (workload, batch dict in the fitter's flattened layout)."""
    from get_b200 import synthetic
    from get_b200.keywords import KeyWordSettings as K
    B, L, R, n, D, H, V, n_src = (int(x) for x in gold["cfg/dims"])
    cnt = gold["in/" + K.EvidenceCountPerQuery].astype(np.int64)
    B1 = int(cnt.sum())
    w = synthetic.get_workload("snopes", name="real_snopes", batch_claims=B, vocab=V, emb_dim=D, hidden=H, n_article_sources=n_src)
    batch = {"query": np.zeros((B, L), np.int64), K.Query_Adj: np.zeros((B, L, L)), K.Query_lens: np.zeros((B,), np.int64),
             "document": np.zeros((B, n, R), np.int64), K.Doc_lens: np.zeros((B, n), np.int64),
             K.DocSources: gold["in/" + K.DocSources].astype(np.int64), K.DocContentNoPaddingEvidence: np.zeros((B1, R), np.int64),
             K.Evd_Docs_Adj: np.zeros((B1, R, R)), "e_lens": np.zeros((B1,), np.int64),
             "raw_query_tokens": gold["in/raw_query_tokens"], "raw_query_lens": gold["in/raw_query_lens"],
             "raw_doc_tokens": gold["in/raw_doc_tokens"], "raw_doc_lens": gold["in/raw_doc_lens"],
             K.EvidenceCountPerQuery: cnt, K.FIXED_NUM_EVIDENCES: n, "window": 3,
             K.QuerySources: np.zeros((B, 1), np.int64), "labels": gold["in/labels"].astype(np.int64), "pairs": B1}
    g = 0
    for b in range(B):
        ql = int(gold["in/raw_query_lens"][b])
        nodes, adj, nn = synthetic.word_graph(gold["in/raw_query_tokens"][b, :ql], L, 3)
        batch["query"][b], batch[K.Query_Adj][b], batch[K.Query_lens][b] = nodes, adj, nn
        for j in range(int(cnt[b])):
            dl = int(gold["in/raw_doc_lens"][g])
            nodes, adj, nn = synthetic.word_graph(gold["in/raw_doc_tokens"][g, :dl], R, 3)
            batch["document"][b, j], batch[K.Doc_lens][b, j] = nodes, nn
            batch[K.DocContentNoPaddingEvidence][g], batch[K.Evd_Docs_Adj][g], batch["e_lens"][g] = nodes, adj, nn
            g += 1
    ref_adj = np.zeros_like(batch[K.Evd_Docs_Adj])
    gi, ri, ci = (gold["in/adj_idx"][i].astype(np.int64) for i in range(3))
    ref_adj[gi, ri, ci] = gold["in/adj_val"]
    assert np.array_equal(ref_adj, batch[K.Evd_Docs_Adj]), "graph restatement differs from the reference's convert_text on real text"
    return w, batch
