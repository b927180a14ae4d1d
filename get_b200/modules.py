"""Drop-in `nn.Module`s with the reference's names, constructor arguments, forward signatures and
`state_dict` keys, computing through libget_b200.so (no torch math on the hot path, no CPU path).

| class here                            | reference                                                     |
|---------------------------------------|---------------------------------------------------------------|
| Linear                                | Models/BiDAF/wrapper.py:330-347                               |
| GGNN                                  | Models/BiDAF/wrapper.py:174-208                               |
| GSL                                   | Models/BiDAF/wrapper.py:210-227                               |
| GGNN_with_GSL                         | Models/BiDAF/wrapper.py:153-172                               |
| ConcatNotEqualSelfAtt                 | thirdparty/two_branches_attention.py:112-148                  |
| MultiHeadSelfAttentionICLR2017Extend  | thirdparty/self_attention.py:51-100                           |
| LSTM (inert parameters only)          | Models/BiDAF/wrapper.py:229-276 (never called by GET)         |
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops


class Linear(nn.Module):
    """`Linear` wrapper (wrapper.py:330-347): Kaiming-normal weight; the bias keeps nn.Linear's default init
    because `hasattr(self, 'linear.bias')` is always False in the reference (:341)."""

    def __init__(self, in_features, out_features, bias=True, dropout=0.0):
        super().__init__()
        self.linear = nn.Linear(in_features=in_features, out_features=out_features, bias=bias)
        self.p_drop = float(dropout)
        self.reset_params()

    def reset_params(self):
        nn.init.kaiming_normal_(self.linear.weight)

    def forward(self, x):
        if self.p_drop > 0 and self.training:
            raise NotImplementedError("Linear(dropout>0) is never used by GET (wrapper.py:333-334)")
        return ops.linear(x, self.linear.weight, self.linear.bias)


class GGNN(nn.Module):
    """Gated graph layer (wrapper.py:174-208). forward(adj (G,N,N), x (G,N,in)) -> (G,N,out)."""

    def __init__(self, in_features, out_features, dropout=0.2):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.proj = Linear(in_features, out_features, bias=False)
        self.linearz0 = Linear(out_features, out_features)
        self.linearz1 = Linear(out_features, out_features)
        self.linearr0 = Linear(out_features, out_features)
        self.linearr1 = Linear(out_features, out_features)
        self.linearh0 = Linear(out_features, out_features)
        self.linearh1 = Linear(out_features, out_features)
        self.p_drop = float(dropout) if dropout and dropout > 0 else 0.0
        self.last_out_planes = None   # bf16 planes of the last output when requested (package-internal)
        self.last_rowdot = None       # partial row dots of the last output when requested (package-internal)

    def _params(self):
        return (self.proj.linear.weight,
                self.linearz0.linear.weight, self.linearz0.linear.bias,
                self.linearz1.linear.weight, self.linearz1.linear.bias,
                self.linearr0.linear.weight, self.linearr0.linear.bias,
                self.linearr1.linear.weight, self.linearr1.linear.bias,
                self.linearh0.linear.weight, self.linearh0.linear.bias,
                self.linearh1.linear.weight, self.linearh1.linear.bias)

    def forward(self, adj, x=None, *, table=None, ids=None, keep=None, pre_agg=None, seed=None, exact_fwd=False,
                out_planes=0, rowdot_of=None):
        """Reference call: forward(adj, x). Extensions used inside this package: (table, ids) = frozen embedding
        table + token ids instead of x (gather fused into the projection), keep / pre_agg from the GSL kernel,
        explicit dropout seed (tests), out_planes > 0: also keep the bf16 planes of the output (the operand of the next
        tensor-core contraction) in `self.last_out_planes`."""
        p = self.p_drop if self.training else 0.0
        if p > 0 and seed is None:
            seed = ops.new_seed()
        adj = ops.as_lists(adj if isinstance(adj, ops.NeighborLists) else adj.float())
        out, op, rd = ops.ggnn_layer(adj, x, table, ids, keep, pre_agg, p, seed or 0, self._params(), exact_fwd=exact_fwd,
                                     out_planes=out_planes, rowdot_of=rowdot_of)
        self.last_out_planes, self.last_rowdot = op, rd
        return out


class GSL(nn.Module):
    """Graph-structure refinement (wrapper.py:210-227): keep edge (i,j) iff i or j is in the top-int(rate*N)
    scored nodes. Stand-alone op-level surface; GGNN_with_GSL uses the fused kernel instead."""

    def __init__(self, rate):
        super().__init__()
        self.rate = rate

    def forward(self, adj, score):
        n = adj.shape[-1]
        out, _ = ops.gsl_mask_adj(adj.float(), score.float(), int(self.rate * n))
        return out


class GGNN_with_GSL(nn.Module):
    """feat_prop1 -> word_scorer1 -> gsl1 -> feat_prop2 (wrapper.py:153-172), with scorer + top-k + refined
    aggregation fused in one kernel. `word_scorer1` keeps the reference's parameter names/shapes
    (GGNN(hidden, 1)); it never receives a gradient in the reference either (top-k indices only)."""

    def __init__(self, input_dim, hidden_dim, output_dim, rate=0.8, dropout=0.2):
        super().__init__()
        self.feat_prop1 = GGNN(input_dim, hidden_dim, dropout)
        self.word_scorer1 = GGNN(hidden_dim, 1, dropout)
        self.gsl1 = GSL(rate)
        self.feat_prop2 = GGNN(hidden_dim, output_dim, dropout)
        self.last_out_planes = None
        self.last_keep = None     # (G,N) uint8 keep set of the last forward (introspection / tests)
        self.last_score = None

    def _scorer_params(self):
        s = self.word_scorer1
        wp = s.proj.linear.weight.detach().reshape(-1).contiguous()
        gate = torch.cat([t.detach().reshape(-1) for t in (
            s.linearz0.linear.weight, s.linearz0.linear.bias, s.linearz1.linear.weight, s.linearz1.linear.bias,
            s.linearr0.linear.weight, s.linearr0.linear.bias, s.linearr1.linear.weight, s.linearr1.linear.bias,
            s.linearh0.linear.weight, s.linearh0.linear.bias, s.linearh1.linear.weight, s.linearh1.linear.bias)])
        return wp, gate

    def forward(self, adj, feat=None, *, table=None, ids=None, seeds=None, want_score=True, out_planes=0):
        # ONE pass over the dense adjacency: neighbour lists shared by both layers, the GSL kernel and the backward pass
        adj = ops.as_lists(adj if isinstance(adj, ops.NeighborLists) else adj.float())
        n = adj.N
        k = int(self.gsl1.rate * n)
        p = self.feat_prop2.p_drop if self.training else 0.0
        if seeds is None:
            seeds = tuple(ops.new_seed() for _ in range(3)) if self.training else (0, 0, 0)
        wp, gate = self._scorer_params()
        ps = self.word_scorer1.p_drop if self.training else 0.0
        assert ps == p or ps == 0 or p == 0, "scorer / layer-2 dropout rates are the same value in the reference"
        # the scorer's projection dropout_s(F1) . w_p leaves feat_prop1's last GEMM as a by-product of its epilogue
        f1 = self.feat_prop1(adj, feat, table=table, ids=ids, seed=seeds[0], exact_fwd=True,    # decides the kept node set
                             rowdot_of=(wp, max(p, ps), seeds[1]))
        sp_parts = self.feat_prop1.last_rowdot
        f1 = ops.grad_marker(f1, "feat_prop2")      # backward: feat_prop2's gradients are complete when dF1 arrives here
        # the refined aggregation leaves the fused kernel as bf16 planes: it is only ever the operand of feat_prop2's
        # projection (wrapper.py:191-192, by linearity of the bias-free proj)
        G, N, H1 = f1.shape
        tc = ops.layer_uses_tc(G * N, self.feat_prop2.out_features, H1)
        score, keep, agg = ops.gsl_fused(adj, f1.detach(), wp, gate, k, drop_p=max(p, ps), seed_scorer=seeds[1],
                                         seed_layer2=seeds[2], want_score=want_score,
                                         planes_n=ops.gemm_mode(False) if tc else 0, sp_parts=sp_parts)
        self.last_keep, self.last_score = keep, score
        out = self.feat_prop2(adj, f1, keep=keep, pre_agg=agg.t if tc else agg, seed=seeds[2], out_planes=out_planes)
        self.last_out_planes = self.feat_prop2.last_out_planes
        return out


class ConcatNotEqualSelfAtt(nn.Module):
    """Multi-head additive attention pooling of `right` conditioned on `left`
    (two_branches_attention.py:112-148). forward(left (G,X), right (G,P,D), mask (G,P)) ->
    (attended (G,D,heads), attention (G,P,heads))."""

    def __init__(self, inp_dim: int, out_dim: int, num_heads: int = 1):
        super().__init__()
        self.inp_dim, self.out_dim, self.num_heads = inp_dim, out_dim, num_heads
        self.linear1 = nn.Linear(inp_dim, out_dim, bias=False)
        self.linear2 = nn.Linear(out_dim, num_heads, bias=False)

    def forward(self, left: torch.Tensor, right: torch.Tensor, mask: torch.Tensor,
                right_planes: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference call: forward(left, right, mask). right_planes: bf16 planes of `right` when the producing kernel
        already wrote them (package-internal)."""
        assert left.size(0) == right.size(0), "Must same dimensions"
        assert len(left.size()) == 2 and len(right.size()) == 3
        assert self.inp_dim == (left.size(-1) + right.size(-1))  # due to concat
        return ops.concat_att(left, right, mask, self.linear1.weight, self.linear2.weight, right_planes)


class MultiHeadSelfAttentionICLR2017Extend(nn.Module):
    """self_attention.py:51-100: the same pooling without `left`; returns (G,heads,D)."""

    def __init__(self, inp_dim: int, out_dim: int, num_heads: int):
        super().__init__()
        self.inp_dim, self.out_dim, self.num_heads = inp_dim, out_dim, num_heads
        self.linear1 = nn.Linear(inp_dim, out_dim, bias=False)
        self.linear2 = nn.Linear(out_dim, num_heads, bias=False)

    def forward(self, tsr: torch.Tensor, mask: torch.Tensor, return_att_weights=False):
        assert len(tsr.size()) == 3
        assert tsr.size(-1) == self.inp_dim
        attended, att = ops.concat_att(None, tsr, mask, self.linear1.weight, self.linear2.weight)
        attended = attended.permute(0, 2, 1)
        if return_att_weights:
            return attended, att
        return attended


class LSTM(nn.Module):
    """Parameter container only: the reference builds `bilstm` / `query_bilstm` (basic_fc_model.py:49-52) but GET
    never calls them; they exist here so `state_dict()` keys and shapes match the reference checkpoint."""

    def __init__(self, input_size, hidden_size, batch_first=False, num_layers=1, bidirectional=False, dropout=0.2):
        super().__init__()
        self.rnn = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                           bidirectional=bidirectional, batch_first=batch_first)
        self.reset_params()

    def reset_params(self):
        for i in range(self.rnn.num_layers):
            for sfx in ("", "_reverse") if self.rnn.bidirectional else ("",):
                nn.init.orthogonal_(getattr(self.rnn, "weight_hh_l%s%s" % (i, sfx)))
                nn.init.kaiming_normal_(getattr(self.rnn, "weight_ih_l%s%s" % (i, sfx)))
                nn.init.constant_(getattr(self.rnn, "bias_hh_l%s%s" % (i, sfx)), val=0)
                nn.init.constant_(getattr(self.rnn, "bias_ih_l%s%s" % (i, sfx)), val=0)

    def forward(self, *a, **k):
        raise NotImplementedError("LSTM is dead code in GET (never called by Graph_basedSemantiStructure)")
