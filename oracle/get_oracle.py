"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

CPU restatement (plain PyTorch fp32/fp64 tensor ops, no custom kernels) of the reference algorithm for
GET's hot path, written as pure functions over the reference's `state_dict` names. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this module,
and only as the checker or the timed CPU baseline; `get_b200` never imports it.

Parity pinning: the reference has no golden vectors or known-answer tests for this path (SURVEY.md
section 4 / 8c). This oracle is therefore pinned against outputs of the reference itself, produced in the
build container by `tests/golden/make_golden.py` (imports the unmodified reference from /root/reference)
and committed under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks the oracle against them.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# kwargs names (reference setting_keywords.py:17-47); literal strings so the oracle has no package imports
KW_QUERY_ADJ = "query_adj"
KW_DOCS_ADJ = "docs_adj"
KW_QUERY_LENS = "query_lens"
KW_DOC_CONTENT = "doc_content_without_padding_evidences"
KW_EVD_CNT = "evd_cnt_each_query"
KW_FIXED_N = "fixed_num_evidences"
KW_QUERY_SOURCES = "query_sources"
KW_DOC_SOURCES = "doc_sources"
KW_OUTPUT_RANKING = "output_ranking"


def _lin(x: Tensor, sd: SD, prefix: str, bias: bool = True) -> Tensor:
    """`Linear` wrapper = nn.Linear (Models/BiDAF/wrapper.py:330-347)."""
    return F.linear(x, sd[prefix + ".linear.weight"], sd[prefix + ".linear.bias"] if bias else None)


def ggnn(adj: Tensor, x: Tensor, sd: SD, prefix: str, drop_mask: Optional[Tensor] = None,
         return_parts: bool = False):
    """`GGNN.forward` (Models/BiDAF/wrapper.py:188-208).

    `drop_mask` (same shape as x, already scaled by 1/(1-p)) stands in for `nn.Dropout` (:189-190) so that
    train-mode parity can be checked with externally supplied masks; None = eval mode."""
    if drop_mask is not None:
        x = x * drop_mask
    x = _lin(x, sd, prefix + ".proj", bias=False)                                   # :191
    a = adj.matmul(x)                                                               # :192
    z = torch.sigmoid(_lin(a, sd, prefix + ".linearz0") + _lin(x, sd, prefix + ".linearz1"))   # :194-196
    r = torch.sigmoid(_lin(a, sd, prefix + ".linearr0") + _lin(x, sd, prefix + ".linearr1"))   # :198-200
    h = torch.tanh(_lin(a, sd, prefix + ".linearh0") + _lin(r * x, sd, prefix + ".linearh1"))  # :202-204
    out = h * z + x * (1 - z)                                                       # :206
    if return_parts:
        return out, dict(x=x, a=a, z=z, r=r, h=h)
    return out


def gsl_topk(score: Tensor, rate: float) -> Tensor:
    """Indices of the preserved nodes: `score.topk(int(rate*N), 1)` (wrapper.py:215-219). score (G,N,1)."""
    n = score.shape[1]
    k = int(rate * n)
    _, idx = score.topk(k, 1)
    return idx.squeeze(-1)


def gsl(adj: Tensor, score: Tensor, rate: float) -> Tensor:
    """`GSL.forward` (wrapper.py:215-227): keep edge (i,j) iff i or j is among the top-k scored nodes."""
    idx = gsl_topk(score, rate)
    g, n = adj.shape[0], adj.shape[-1]
    keep = torch.zeros(g, n, dtype=torch.bool, device=adj.device)
    keep.scatter_(1, idx, True)
    mask = keep.unsqueeze(2) | keep.unsqueeze(1)          # rows filled, then columns filled (:222-224)
    return adj * mask.to(adj.dtype)                       # :225


def ggnn_with_gsl(adj: Tensor, feat: Tensor, sd: SD, prefix: str, rate: float,
                  drop_masks: Optional[Tuple[Tensor, Tensor, Tensor]] = None, return_parts: bool = False,
                  keep_override: Optional[Tuple[Tensor, Tensor]] = None):
    """`GGNN_with_GSL.forward` (wrapper.py:165-172). drop_masks = (feat_prop1, word_scorer1, feat_prop2).

    keep_override = (graphs (G,) bool, keep (G,N) bool): near-tie policy of the parity harness -- on
    the listed graphs (those whose k-th score gap is below the fp32 re-association noise) the oracle follows the GIVEN
    kept-node set instead of its own top-k, so that everything downstream can still be compared."""
    m1, ms, m2 = drop_masks if drop_masks is not None else (None, None, None)
    f1 = ggnn(adj, feat, sd, prefix + ".feat_prop1", m1)          # :166
    score = ggnn(adj, f1, sd, prefix + ".word_scorer1", ms)       # :167
    adj_refined = gsl(adj, score, rate)                           # :168
    if keep_override is not None:
        graphs, keep = keep_override
        graphs, keep = graphs.to(adj.device), keep.to(adj.device)
        forced = adj * (keep.unsqueeze(2) | keep.unsqueeze(1)).to(adj.dtype)
        adj_refined = torch.where(graphs.view(-1, 1, 1), forced, adj_refined)
    f2 = ggnn(adj_refined, f1, sd, prefix + ".feat_prop2", m2)    # :169
    if return_parts:
        return f2, dict(f1=f1, score=score, keep_idx=gsl_topk(score, rate), adj_refined=adj_refined)
    return f2


def concat_not_equal_self_att(left: Tensor, right: Tensor, mask: Tensor, w1: Tensor, w2: Tensor):
    """`ConcatNotEqualSelfAtt.forward` (thirdparty/two_branches_attention.py:121-148)."""
    g, p, _ = right.shape
    assert left.shape[0] == g and left.dim() == 2 and right.dim() == 3          # :133-134
    assert w1.shape[1] == left.shape[-1] + right.shape[-1]                      # :135
    tsr = torch.cat([left.unsqueeze(1).expand(g, p, -1), right], dim=-1)        # :137-138
    e = F.linear(torch.tanh(F.linear(tsr, w1)), w2)                             # :140-141
    e = e.masked_fill((mask == 0).unsqueeze(-1).expand(g, p, w2.shape[0]), float("-inf"))   # :142-144
    att = F.softmax(e, dim=1)                                                   # :146
    attended = torch.bmm(right.permute(0, 2, 1), att)                           # :147
    return attended, att


def multi_head_self_att_extend(tsr: Tensor, mask: Tensor, w1: Tensor, w2: Tensor, return_att_weights=False):
    """`MultiHeadSelfAttentionICLR2017Extend.forward` (thirdparty/self_attention.py:75-100)."""
    g, p, _ = tsr.shape
    e = F.linear(torch.tanh(F.linear(tsr, w1)), w2)                             # :89-90
    e = e.masked_fill((mask == 0).unsqueeze(-1).expand(g, p, w2.shape[0]), float("-inf"))   # :91-93
    att = F.softmax(e, dim=1)                                                   # :95
    attended = torch.bmm(tsr.permute(0, 2, 1), att).permute(0, 2, 1)            # :96-99
    return (attended, att) if return_att_weights else attended


def pad_left(left: Tensor, evd_cnt: Tensor) -> Tensor:
    """`_pad_left_tensor` (Models/FCWithEvidences/basic_fc_model.py:80-92): repeat row c n_c times."""
    assert evd_cnt.shape[0] == left.shape[0]
    return torch.repeat_interleave(left, evd_cnt.to(torch.long).to(left.device), dim=0)


def pad_right(tsr: Tensor, evd_cnt: Tensor, max_num_evd: int) -> Tensor:
    """`_pad_right_tensor` (basic_fc_model.py:94-121): (B1,Y) -> zero padded (B,n,Y)."""
    cnt = [int(c) for c in evd_cnt.tolist()]
    last = 0
    rows = []
    for b, c in enumerate(cnt):
        rows.append(F.pad(tsr[last:last + c], (0, 0, 0, max_num_evd - c)))
        last += c
    out = torch.stack(rows, dim=0)
    assert out.shape == (len(cnt), max_num_evd, tsr.shape[1])
    return out


def model_forward(sd: SD, cfg: dict, query: Tensor, document: Tensor, kargs: dict,
                  drop_masks: Optional[dict] = None, dtype=torch.float32, return_parts: bool = False,
                  keep_override: Optional[Tuple[Tensor, Tensor]] = None):
    """`Graph_basedSemantiStructure.forward` (Models/FCWithEvidences/graph_based_semantic_structure.py:76-125).

    cfg keys: gsl_rate, use_claim_source, use_article_source. drop_masks (train-mode parity) may hold
    'claim', 'feat_prop1', 'word_scorer1', 'feat_prop2' pre-scaled masks."""
    dm = drop_masks or {}
    assert query.shape[0] == document.shape[0]                                   # :91
    n = document.shape[1]
    doc = kargs[KW_DOC_CONTENT]                                                  # :97
    doc_mask = doc >= 1                                                          # :98
    doc_adj = kargs[KW_DOCS_ADJ].to(dtype)                                       # :99
    emb = sd["embedding.weight"]
    embed_doc = F.embedding(doc.long(), emb)                                     # :100
    evd_cnt = kargs[KW_EVD_CNT]
    # claim representation (:144-155)
    q_mask = (query > 0).unsqueeze(2)
    q_lens = kargs[KW_QUERY_LENS].unsqueeze(-1)
    q_adj = kargs[KW_QUERY_ADJ].to(dtype)
    q_hid = ggnn(q_adj, F.embedding(query.long(), emb), sd, "ggnn4claim_1", dm.get("claim"))
    q_claim = torch.sum(q_hid * q_mask.to(dtype), dim=1) / q_lens.to(dtype)      # :153  (B,H)
    query_repr = pad_left(q_claim, evd_cnt)                                      # :154  (B1,H)
    # evidence graphs (:107)
    gm = (dm.get("feat_prop1"), dm.get("word_scorer1"), dm.get("feat_prop2"))
    doc_out, parts = ggnn_with_gsl(doc_adj, embed_doc, sd, "ggnn_with_gsl", cfg["gsl_rate"],
                                   gm if any(m is not None for m in gm) else None, return_parts=True,
                                   keep_override=keep_override)
    # word-level attention (:110, :173-193)
    avg, word_att = concat_not_equal_self_att(query_repr, doc_out, doc_mask,
                                              sd["self_att_word.linear1.weight"], sd["self_att_word.linear2.weight"])
    avg = torch.flatten(avg, start_dim=1)                                        # :191 index = d*heads+head
    if cfg["use_claim_source"]:                                                  # :113-118
        c_emb = F.embedding(kargs[KW_QUERY_SOURCES].long(), sd["claim_source_embs.weight"]).squeeze(1)
        query_repr = torch.cat([pad_left(c_emb, evd_cnt), query_repr], dim=-1)
    # evidence-level attention (:195-221)
    new_left = pad_right(query_repr, evd_cnt, n)[:, 0, :]                        # :210-211
    padded = pad_right(avg, evd_cnt, n)                                          # :213
    evd_mask = (torch.sum(document, dim=-1) >= 1).to(dtype)                      # :215
    if cfg["use_article_source"]:                                                # :157-171
        src = kargs[KW_DOC_SOURCES]
        src = src.masked_fill(src == -1, 0)
        padded = torch.cat([padded, F.embedding(src.long(), sd["article_source_embs.weight"])], dim=-1)
    att_avg, evd_att = concat_not_equal_self_att(new_left, padded, evd_mask,
                                                 sd["self_att_evd.linear1.weight"], sd["self_att_evd.linear2.weight"])
    final = torch.cat([new_left, torch.flatten(att_avg, start_dim=1)], dim=-1)   # :220, :264-267
    hid = F.linear(final, sd["out.0.weight"], sd["out.0.bias"])                  # :121 (no activation between)
    logits = F.linear(hid, sd["out.1.weight"], sd["out.1.bias"])
    if return_parts:
        parts = dict(parts, q_claim=q_claim, doc_out=doc_out, word_pooled=avg, evd_pooled=att_avg,
                     word_att=word_att, evd_att=evd_att, final=final)
        return logits, parts
    if kargs.get(KW_OUTPUT_RANKING, False):                                      # :123-124
        return logits, (word_att, evd_att)
    return logits


def cross_entropy(logits: Tensor, labels: Tensor) -> Tensor:
    """`losses.cross_entroy` (losses.py:29-32): mean CE over the claims of the batch."""
    assert logits.shape[0] == labels.shape[0]
    return F.cross_entropy(logits, labels.long())


# parameters that never receive a gradient in the reference
INERT_PREFIXES = ("bilstm.", "query_bilstm.", "trans.", "ggnn_with_gsl.word_scorer1.", "embedding.")


def loss_and_grads(sd: SD, cfg: dict, query, document, labels, kargs, drop_masks=None, dtype=torch.float32,
                   keep_override=None):
    """Forward + autograd backward; returns (loss, logits, {name: grad}) for grad-receiving parameters."""
    leaves = {}
    for k, v in sd.items():
        v = v.detach().to(dtype) if v.is_floating_point() else v
        if v.is_floating_point() and not k.startswith(INERT_PREFIXES):
            v = v.clone().requires_grad_(True)
        leaves[k] = v
    logits = model_forward(leaves, cfg, query, document, kargs, drop_masks, dtype, keep_override=keep_override)
    loss = cross_entropy(logits, labels)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.is_floating_point() and v.requires_grad and v.grad is not None}
    return loss.detach(), logits.detach(), grads


def near_tie_graphs(score: Tensor, rate: float, tol: float = 1e-5) -> Tensor:
    """SURVEY.md App. A.2 near-tie policy: graphs whose k-th / (k+1)-th score gap is below `tol` may
    legitimately select a different node under fp32 re-association. Returns a bool mask (G,)."""
    s = score.squeeze(-1)
    k = int(rate * s.shape[1])
    top, _ = s.topk(min(k + 1, s.shape[1]), 1)
    if k >= s.shape[1]:
        return torch.zeros(s.shape[0], dtype=torch.bool)
    gap = (top[:, k - 1] - top[:, k]).abs()
    # gap == 0 is an exact tie, which only happens among pad nodes (identical features, zero adjacency rows
    # and columns): whichever of them is kept, the refined adjacency is the same
    return (gap < tol) & (gap > 0)


def neighbor_lists(adj: np.ndarray, transpose: bool = False):
    """CSR restatement of a dense normalised adjacency (the operand of `adj.matmul(x)`, wrapper.py:192): (rowptr (N+1,),
    index (nnz,), weight (nnz,), used) with rows in order and neighbours in increasing index; used = 1 + the highest
    neighbour index (0 for an empty graph). This is the record format of get_build_neighbor_lists (include/get_b200.h);
    sum_e weight[e] * x[index[e]] over a row's entries IS the row of adj @ x."""
    a = np.asarray(adj)
    if transpose:
        a = a.T
    rows, cols = np.nonzero(a)                              # row-major order
    counts = np.bincount(rows, minlength=a.shape[0])
    rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    used = int(cols.max()) + 1 if cols.size else 0
    return rowptr, cols.astype(np.int64), a[rows, cols], used


def aggregate_lists(rowptr, index, weight, x: np.ndarray, keep: Optional[np.ndarray] = None) -> np.ndarray:
    """out[i] = sum over the entries e of row i of weight[e] * x[index[e]]; with `keep`, edges whose two endpoints are both
    dropped are removed (the GSL mask of wrapper.py:221-225 applied per edge)."""
    out = np.zeros((len(rowptr) - 1, x.shape[1]), dtype=np.float64)
    for i in range(len(rowptr) - 1):
        for e in range(rowptr[i], rowptr[i + 1]):
            j = index[e]
            if keep is not None and not keep[i] and not keep[j]:
                continue
            out[i] += float(weight[e]) * x[j].astype(np.float64)
    return out
