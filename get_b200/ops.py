"""Host-side wrappers over the C-ABI (raw device pointers in, nothing allocated natively) and the
`torch.autograd.Function`s the drop-in modules are built from.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every arithmetic step of the
hot path is a kernel of libget_b200.so. Nothing in this file falls back to torch math when the library or a
GPU is missing -- the calls raise.
"""
import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, planes
from ._lib import (BPE_DGATE_R, BPE_STORE, BPE_TANH, BPE_TANH_BLEND, BPE_TANH_ROWGROUP, BPE_ZR, EPI_DGATE_R,
                   EPI_DROPOUT_OUT, EPI_SIGMOID, EPI_STORE, EPI_TANH, EPI_TANH_BLEND, EPI_TANH_ROWGROUP, GemmDesc)
from .planes import Planes, alloc_planes, gemm_bp, get_pack, pack_of, round_up, to_planes, wgrad_bp

_SM_COUNT = 148
# bench.py sets this to a list to time the fused GSL kernel with CUDA events on its launch stream:
# entries are (start_event, stop_event, n_graphs)
PROFILE_GSL_EVENTS = None
# bench.py sets this to a list to record the exact arguments of every fused GSL launch of a step (tensors kept alive) so
# that the very same launches can be replayed back to back under CUDA events (gsl_fused_replay)
PROFILE_GSL_ARGS = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError("get_b200: %s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("get_b200: %s must be float32, got %s" % (name, t.dtype))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Raw(object):
    """Explicit GEMM operand (used for the embedding table addressed through a row gather, whose logical
    (M, K) extent is larger than the table itself)."""
    __slots__ = ("ptr", "ld", "trans", "shape")

    def __init__(self, ptr, ld, trans, shape):
        self.ptr, self.ld, self.trans, self.shape = ptr, ld, trans, tuple(shape)

    def t(self):
        return Raw(self.ptr, self.ld, 1 - self.trans, self.shape[::-1])


def _operand(t, name: str):
    """2-D logical (i, k) tensor view -> (ptr, ld, trans)."""
    if isinstance(t, Raw):
        return t.ptr, t.ld, t.trans
    _chk_f32(t, name)
    assert t.dim() == 2, name
    s0, s1 = t.stride()
    if t.shape[1] == 1 or s1 == 1:
        return t.data_ptr(), (s0 if t.shape[0] > 1 else max(s0, t.shape[1])), 0
    if t.shape[0] == 1 or s0 == 1:
        return t.data_ptr(), s1, 1
    raise RuntimeError("get_b200.gemm: operand %s has no unit stride %s" % (name, (t.stride(),)))


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1)
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# Precision class of the contractions that do NOT feed the GSL top-k (the forward of feat_prop1 decides the kept node set
# and runs in the fp32-exact class in every mode):
#   "fp32"  (default): 16-bit operands (two bf16 planes, three tensor-core products; relative error ~1e-5) -- the 1e-4
#           parity class of BASELINE.json configs[1];
#   "fp32x": the fp32-exact class everywhere (three planes, six products);
#   "bf16":  plain bf16 operands, fp32 accumulation -- the 1e-2 parity class of BASELINE.json configs[2] ("fast" = alias).
PRECISION = os.environ.get("GET_B200_PRECISION", "fp32")


def set_precision(mode: str):
    global PRECISION
    if mode == "fast":
        mode = "bf16"
    assert mode in ("fp32", "fp32x", "bf16"), mode
    PRECISION = mode


def gemm_mode(exact: bool = False) -> int:
    """Plane-product mode of get_gemm_bp for the current precision class (3 = fp32-exact, 2 = 16-bit operands, 1 = bf16)."""
    if exact or PRECISION == "fp32x":
        return 3
    return 2 if PRECISION == "fp32" else 1


# Tensor-core (tcgen05, bf16 planes) path for the large contractions; GET_B200_TC=0 forces the exact SIMT path everywhere
# (debugging / A-B comparisons only).
TC_ENABLED = os.environ.get("GET_B200_TC", "1") != "0"
DEBUG_TC_REPORT = False      # tests: record in LAST_GEMM_USED_TC whether the last gemm() ran on the tcgen05 path
LAST_GEMM_USED_TC = None

# Weight gradients written straight into caller-owned buffers (the flat gradient bucket of get_b200.ddp): maps
# parameter data_ptr -> (fp32 view of the same shape, weakref to the parameter). When a parameter is registered here the backward passes ACCUMULATE
# into the view (the owner zeroes the bucket at the start of a step) and return no gradient for it.
GRAD_SINK = {}


def sink_of(param: torch.Tensor):
    """Registered gradient sink of a parameter or None. Entries are (view, weakref to the parameter): an entry whose
    parameter has died (its address may have been recycled) is dropped."""
    ent = GRAD_SINK.get(param.data_ptr())
    if ent is None:
        return None
    view, ref = ent
    owner = ref()
    if owner is None or owner.data_ptr() != param.data_ptr() or tuple(view.shape) != tuple(param.shape):
        GRAD_SINK.pop(param.data_ptr(), None)
        return None
    if owner.grad is None or owner.grad.data_ptr() != view.data_ptr():
        return None        # the caller re-pointed / dropped p.grad (e.g. zero_grad(set_to_none=True)): plain autograd flow
    return view


# Called with a tag from the backward pass when every gradient of a group of layers has been written (see grad_marker):
# the data-parallel reducer starts the all-reduce of that chunk of the bucket on its communication stream.
GRAD_READY_HOOK = None


class _GradMarkerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        hook = GRAD_READY_HOOK
        if hook is not None:
            hook(ctx.tag)
        return g, None


def grad_marker(x: torch.Tensor, tag: str) -> torch.Tensor:
    """Identity. In the backward pass, when the gradient of `x` arrives here every layer applied AFTER this point in the
    forward pass has finished its backward: GRAD_READY_HOOK(tag) is called at that moment."""
    if GRAD_READY_HOOK is None or not torch.is_grad_enabled() or not x.requires_grad:
        return x
    return _GradMarkerFn.apply(x, tag)


def weights_updated(*_args, **_kwargs):
    """Invalidate every packed weight. Registered as a global optimizer post-step hook: fused / foreach optimizer
    kernels update parameters without moving their version counters, so the counter alone cannot be trusted."""
    planes.weights_updated()


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(weights_updated)
except Exception:      # very old torch: callers must invoke ops.weights_updated() after optimizer.step()
    pass


def begin_step_capture():
    """Call at the head of a training step that is being captured into a CUDA graph: the weights differ at every
    replay, so every weight must be re-packed INSIDE the captured step (first use), whatever its version counter says."""
    planes.begin_step_capture()


def prepare_split_table():
    """(Re)build the device job table covering every packed weight (host -> device copy: outside stream captures only)."""
    planes.prepare_pack_table()


def refresh_weight_splits() -> bool:
    """Bring every packed weight up to date NOW on the current stream (one launch); True when all packs are current."""
    return planes.refresh_packs()


def dropout_salt_set(value: int):
    _lib.check(_lib.load().get_dropout_salt_set(int(value) & 0xFFFFFFFF, _stream()), "get_dropout_salt_set")


def dropout_salt_advance():
    _lib.check(_lib.load().get_dropout_salt_advance(_stream()), "get_dropout_salt_advance")


def dropout_salt_get() -> int:
    v = C.c_uint32(0)
    _lib.check(_lib.load().get_dropout_salt_get(C.byref(v)), "get_dropout_salt_get")
    return int(v.value)


def rows_gather_dropout(src: torch.Tensor, idx: Optional[torch.Tensor], rows: int, p: float, seed: int) -> torch.Tensor:
    """out (rows, W) = dropout(src[idx]) (idx int64 or None = identity); see get_rows_gather_dropout_f32."""
    lib = _lib.load()
    _chk_f32(src, "src")
    assert src.dim() == 2 and src.stride(1) == 1
    W = src.shape[1]
    out = torch.empty((rows, W), dtype=torch.float32, device=src.device)
    if idx is not None:
        assert idx.dtype == torch.int64 and idx.is_cuda and idx.is_contiguous() and idx.numel() == rows
    _lib.check(lib.get_rows_gather_dropout_f32(src.data_ptr(), src.stride(0), _ptr(idx), rows, W, float(p), seed & 0xFFFFFFFF,
                                               out.data_ptr(), W, _stream()), "get_rows_gather_dropout_f32")
    return out


def rows_gather_dropout_planes(src: torch.Tensor, idx: Optional[torch.Tensor], rows: int, p: float, seed: int,
                               nplanes: int) -> Planes:
    """Planes of dropout(src[idx]) (idx int64 or None = identity): the A operand of a GGNN input projection."""
    lib = _lib.load()
    _chk_f32(src, "src")
    assert src.dim() == 2 and src.stride(1) == 1
    W = src.shape[1]
    out = alloc_planes(nplanes, rows, W, src.device)
    if idx is not None:
        assert idx.dtype == torch.int64 and idx.is_cuda and idx.is_contiguous() and idx.numel() == rows
    _lib.check(lib.get_rows_gather_dropout_bp(src.data_ptr(), src.stride(0), src.shape[0] if idx is not None else 0, _ptr(idx),
                                              rows, W, float(p), seed & 0xFFFFFFFF, out.ptr, out.ld, out.plane_stride, nplanes,
                                              _stream()), "get_rows_gather_dropout_bp")
    return out


def segments(evd_cnt: torch.Tensor, b1: int, n: int):
    """seg_of_row (B1,), slot_of_row (B1,) = claim*n + j, offsets (B+1,) (all int32) from the per-claim evidence counts in one
    launch: the index arithmetic that replaces the per-claim loops of basic_fc_model.py:80-121. No host sync (B1 is the
    shape of the flattened evidence tensor)."""
    lib = _lib.load()
    if not evd_cnt.is_cuda:
        raise RuntimeError("get_b200: evd_cnt must be a CUDA tensor (there is no CPU path)")
    if evd_cnt.dtype not in (torch.int64, torch.int32):
        evd_cnt = evd_cnt.to(torch.int64)
    evd_cnt = evd_cnt.contiguous()
    B = evd_cnt.shape[0]
    dev = evd_cnt.device
    seg = torch.empty((b1,), dtype=torch.int32, device=dev)
    slot = torch.empty((b1,), dtype=torch.int32, device=dev)
    off = torch.empty((B + 1,), dtype=torch.int32, device=dev)
    _lib.check(lib.get_segments_i32(evd_cnt.data_ptr(), int(evd_cnt.dtype == torch.int64), B, int(b1), int(n), seg.data_ptr(),
                                    slot.data_ptr(), off.data_ptr(), _stream()), "get_segments_i32")
    return seg, slot, off


def ids_mask(ids: torch.Tensor, reduce_last: bool = False) -> torch.Tensor:
    """uint8 attention mask from token ids: `ids >= 1` element-wise (gbss.py:98), or with reduce_last
    `sum(ids, -1) >= 1` (gbss.py:215)."""
    lib = _lib.load()
    if not ids.is_cuda:
        raise RuntimeError("get_b200: ids must be a CUDA tensor (there is no CPU path)")
    if ids.dtype not in (torch.int64, torch.int32):
        ids = ids.to(torch.int64)
    ids = ids.contiguous()
    W = ids.shape[-1] if reduce_last else 1
    shape = ids.shape[:-1] if reduce_last else ids.shape
    out = torch.empty(shape, dtype=torch.uint8, device=ids.device)
    _lib.check(lib.get_ids_mask_u8(ids.data_ptr(), int(ids.dtype == torch.int64), out.numel(), W, out.data_ptr(), _stream()),
               "get_ids_mask_u8")
    return out


class EmbeddingRowsFn(torch.autograd.Function):
    """Trainable source embeddings (base_model.py:184-188; the -1 padding id reads row 0, gbss.py:166-168) written into a
    column block of `out` when given; dense deterministic table gradient."""

    @staticmethod
    def forward(ctx, table, idx):
        lib = _lib.load()
        _chk_f32(table, "table")
        table = table.contiguous()
        idx = idx.reshape(-1).to(torch.int64).contiguous()
        R, (V, E) = idx.shape[0], table.shape
        out = torch.empty((R, E), dtype=torch.float32, device=table.device)
        _lib.check(lib.get_embedding_rows_fwd_f32(table.data_ptr(), V, E, idx.data_ptr(), R, out.data_ptr(), E, _stream()),
                   "get_embedding_rows_fwd_f32")
        ctx.save_for_backward(idx, table)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        idx, table = ctx.saved_tensors
        V, E = table.shape
        g = g if (g.dim() == 2 and g.stride(1) == 1) else g.contiguous()
        dst, acc, ret = _grad_target(table)
        _lib.check(lib.get_embedding_rows_bwd_f32(g.data_ptr(), g.stride(0), idx.data_ptr(), idx.shape[0], V, E, dst.data_ptr(),
                                                  int(acc), _stream()), "get_embedding_rows_bwd_f32")
        return ret, None


def embedding_rows(table, idx):
    return EmbeddingRowsFn.apply(table, idx)


def new_seed() -> int:
    """32-bit dropout seed drawn from torch's CPU generator (follows torch.manual_seed)."""
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


def _as_rowmajor_2d(t: torch.Tensor):
    """(tensor with unit column stride, transposed?) for a 2-D fp32 view with one unit stride."""
    assert t.dim() == 2
    if t.stride(1) == 1 or t.shape[1] == 1:
        return t, False
    if t.stride(0) == 1 or t.shape[0] == 1:
        return t.t(), True
    raise RuntimeError("get_b200.gemm: operand has no unit stride %s" % (t.stride(),))


_BPE_OF = {EPI_STORE: BPE_STORE, EPI_TANH: BPE_TANH, EPI_TANH_ROWGROUP: BPE_TANH_ROWGROUP, EPI_DROPOUT_OUT: BPE_STORE}


def _gemm_tc(segments, out, epilogue, bias0, bias1, aux0, group_rows, accumulate, drop_out_p, drop_out_seed, presplit,
             exact, planes_out=None):
    """The tensor-core route of gemm(): fp32 A operands are converted to planes on the fly (callers on the hot path hand
    over Planes written by the producing kernel instead), weights come from the pack cache."""
    M, N = out.shape
    mode = gemm_mode(exact)
    if presplit:
        segs = []
        bias = None
        for s_i, (a, b) in enumerate(segments):
            ap = a if isinstance(a, Planes) else to_planes(_rows_view(a), mode)
            pk = pack_of(b, bias0, bias1) if (s_i == 0 and bias0 is not None) else pack_of(b)
            if s_i == 0 and bias0 is not None:
                bias = pk.bias
            segs.append((ap, pk.planes, b.shape[1]))
        gemm_bp(segs, M, N, mode=mode, epilogue=_BPE_OF[epilogue], C=out, bias=bias, aux0=aux0, group_rows=group_rows,
                accumulate=accumulate, drop_out=(drop_out_p, drop_out_seed) if epilogue == EPI_DROPOUT_OUT else None,
                planes_out=planes_out)
        return
    # activation x activation (weight gradient): both operands are transposed views of (K, .) row-major matrices
    assert len(segments) == 1 and epilogue == EPI_STORE and bias0 is None
    a, b = segments[0]
    ar, at = (a, True) if isinstance(a, Planes) else _as_rowmajor_2d(a)
    br, bt = (b, True) if isinstance(b, Planes) else _as_rowmajor_2d(b)
    assert at and bt, "weight-gradient contraction expects (K, .) row-major operands passed as .t() views"
    ap = ar if isinstance(ar, Planes) else to_planes(ar, mode)
    bp_ = br if isinstance(br, Planes) else to_planes(br, mode)
    K = ap.rows
    wgrad_bp(ap.T() if ap.trans == 0 else ap, bp_.T() if bp_.trans == 0 else bp_, M, N, K, mode,
             [(out, 0, M, 0, N)], accumulate=accumulate)


def _rows_view(a: torch.Tensor) -> torch.Tensor:
    r, t = _as_rowmajor_2d(a)
    return a.contiguous() if t else r


def gemm(segments: Sequence[Tuple[torch.Tensor, torch.Tensor]], out: torch.Tensor, *, epilogue: int = EPI_STORE,
         bias0=None, bias1=None, aux0=None, aux1=None, out1=None, group_rows: int = 0, alpha: float = 1.0,
         accumulate: bool = False, rowidx: Optional[torch.Tensor] = None, drop_p: float = 0.0, drop_seed: int = 0,
         drop_cols: int = 0, drop_out_p: float = 0.0, drop_out_seed: int = 0, split_k: Optional[int] = None,
         tc: bool = False, presplit: bool = True, exact: bool = False, planes_out: Optional[Planes] = None):
    """out[m,n] = epilogue(sum_s A_s[m,:] . B_s[n,:]); A_s logical (M,K_s), B_s logical (N,K_s).
    tc=True: run on the tcgen05 plane GEMM when the shape allows it (M >= 128, N % 4 == 0, a store / tanh epilogue);
    presplit=True: every B_s is a weight (packed planes come from the cache); presplit=False: both operands are
    activations (a weight gradient). Otherwise (small / odd contractions: output MLP, per-claim projections) the exact
    fp32 SIMT kernel runs. exact=True pins the contraction to the fp32-exact class (the GSL top-k chain)."""
    lib = _lib.load()
    M, N = out.shape
    has_planes = any(isinstance(a, Planes) or isinstance(b, Planes) for a, b in segments)
    want_tc = (tc and TC_ENABLED and (M >= 128 or has_planes) and N % 4 == 0 and N >= 8 and alpha == 1.0 and rowidx is None
               and drop_p == 0.0 and epilogue in _BPE_OF and aux1 is None and out1 is None and out.stride(1) == 1
               and out.stride(0) % 4 == 0 and (not presplit or all(b.shape[1] >= 8 for _, b in segments)))
    if has_planes and not want_tc:
        raise RuntimeError("get_b200.gemm: plane operands need the tensor-core route (N %% 4 == 0, N >= 8, store/tanh epilogue)")
    if DEBUG_TC_REPORT:
        global LAST_GEMM_USED_TC
        LAST_GEMM_USED_TC = 2 if want_tc else 0
    if want_tc:
        _gemm_tc(segments, out, epilogue, bias0, bias1, aux0, group_rows, accumulate, drop_out_p, drop_out_seed, presplit,
                 exact, planes_out)
        return out
    assert planes_out is None
    d = GemmDesc()
    d.nseg = len(segments)
    ktiles = 0
    for s, (a, b) in enumerate(segments):
        ks = a.shape[1]
        assert a.shape[0] == M and b.shape[0] == N and b.shape[1] == ks, (a.shape, b.shape, out.shape)
        d.A[s].ptr, d.A[s].ld, d.A[s].trans = _operand(a, "A%d" % s)
        d.B[s].ptr, d.B[s].ld, d.B[s].trans = _operand(b, "B%d" % s)
        d.K[s] = ks
        ktiles += (ks + 15) // 16
    if rowidx is not None:
        assert rowidx.dtype == torch.int64 and rowidx.is_cuda and rowidx.is_contiguous()
        d.A[0].rowidx = rowidx.data_ptr()
    _chk_f32(out, "C")
    d.M, d.N, d.C, d.ldc = M, N, out.data_ptr(), _ld(out)
    d.alpha, d.accumulate, d.epilogue = alpha, int(accumulate), epilogue
    for name, t in (("bias0", bias0), ("bias1", bias1)):
        if t is not None:
            _chk_f32(t, name)
            assert t.numel() == N and t.is_contiguous()
            setattr(d, name, t.data_ptr())
    for name, t in (("aux0", aux0), ("aux1", aux1), ("out1", out1)):
        if t is not None:
            _chk_f32(t, name)
            setattr(d, name, t.data_ptr())
            setattr(d, "ld_" + name, _ld(t))
    d.group_rows = group_rows
    d.drop_p, d.drop_seed, d.drop_cols = drop_p, drop_seed & 0xFFFFFFFF, drop_cols
    d.drop_out_p, d.drop_out_seed = drop_out_p, drop_out_seed & 0xFFFFFFFF
    if split_k is None:
        tiles = ((M + 127) // 128) * ((N + 63) // 64)
        split_k = 1
        if tiles < _SM_COUNT and ktiles >= 16:
            split_k = max(1, min(ktiles // 8, (2 * _SM_COUNT + tiles - 1) // tiles))
    ws = None
    if split_k > 1:
        ws = torch.empty((split_k * M * N,), dtype=torch.float32, device=out.device)
        d.workspace = ws.data_ptr()
    d.split_k = split_k
    _lib.check(lib.get_gemm_f32(C.byref(d), _stream()), "get_gemm_f32")
    return out


GL_MAX_N = 232          # graph_lists.cu


class NeighborLists(object):
    """Packed neighbour lists of a batch of dense adjacencies (G,N,N), built ONCE per step by one kernel and walked by every
    aggregation of the step (layer-1 and GSL-refined aggregation, the scorer's adj @ s_p, both adj^T products of the
    backward pass) instead of re-scanning the 4-13 % dense matrix. `adj` stays available for
    the dense fallback kernels (N > 256, feature widths that are not multiples of 4)."""

    def __init__(self, adj: torch.Tensor):
        _chk_f32(adj, "adj")
        assert adj.dim() == 3 and adj.shape[1] == adj.shape[2]
        self.adj = adj.contiguous()
        self.G, self.N = int(adj.shape[0]), int(adj.shape[1])
        self.shape, self.device = self.adj.shape, adj.device
        self.ent = self.rowptr = self.used = None
        if self.N <= GL_MAX_N and os.environ.get("GET_B200_GRAPH_LISTS", "1") != "0":
            lib = _lib.load()
            G, N = self.G, self.N
            self.ecap = int(lib.get_neighbor_lists_entry_capacity(N))
            self.pitch = int(lib.get_neighbor_lists_rowptr_pitch(N))
            # CSR records of adj ([0]) and adj^T ([1]): entries {index, weight}, row pointers, rows referred to
            self.ent = torch.empty((2, G, self.ecap, 2), dtype=torch.float32, device=adj.device)
            self.rowptr = torch.empty((2, G, self.pitch), dtype=torch.int32, device=adj.device)
            self.used = torch.empty((2, G), dtype=torch.int32, device=adj.device)
            self.rebuild()

    def rebuild(self):
        """Re-pack after the dense adjacency changed in place."""
        if self.ent is not None:
            _lib.check(_lib.load().get_build_neighbor_lists(self.adj.data_ptr(), self.G, self.N, self.ent.data_ptr(), self.rowptr.data_ptr(),
                                                            self.used.data_ptr(), _stream()), "get_build_neighbor_lists")
        return self

    def usable(self, H: int) -> bool:
        return self.ent is not None and H % 4 == 0 and H >= 4

    def pointers(self, transpose: bool = False):
        o = 1 if transpose else 0
        return self.ent[o].data_ptr(), self.rowptr[o].data_ptr(), self.used[o].data_ptr()

    def nnz(self) -> torch.Tensor:
        """(G,) edges per graph (device tensor)."""
        return self.rowptr[0][:, self.N]


def as_lists(adj) -> NeighborLists:
    return adj if isinstance(adj, NeighborLists) else NeighborLists(adj)


def dense_adj(adj) -> torch.Tensor:
    return adj.adj if isinstance(adj, NeighborLists) else adj


def graph_aggregate(adj, x, keep=None, out=None, transpose=False, accumulate=False, planes_out: Optional[Planes] = None,
                    pad_one: bool = False, want_f32: bool = True):
    """out[g] (+)= op(adj'[g]) @ x[g]; optionally also (or only, want_f32=False) as bf16 planes for the next contraction.
    adj: dense (G,N,N) tensor, or NeighborLists (then the list kernel runs)."""
    lib = _lib.load()
    _chk_f32(x, "x")
    G, N, H = x.shape
    assert tuple(adj.shape) == (G, N, N) and x.is_contiguous()
    if out is None and want_f32:
        assert not accumulate
        out = torch.empty_like(x)
    if keep is not None:
        assert keep.dtype == torch.uint8 and keep.shape == (G, N) and keep.is_contiguous()
    if planes_out is not None:
        assert planes_out.rows == G * N and planes_out.cols == H
    if isinstance(adj, NeighborLists) and adj.usable(H):
        pl = planes_out
        _lib.check(lib.get_graph_gather(*adj.pointers(transpose), x.data_ptr(), _ptr(keep), _ptr(out), pl.ptr if pl else None,
                                        pl.ld if pl else 0, pl.plane_stride if pl else 0, pl.nplanes if pl else 0, int(pad_one),
                                        G, N, H, int(accumulate), _stream()), "get_graph_gather")
        return out
    adj = dense_adj(adj)
    _chk_f32(adj, "adj")
    assert adj.is_contiguous()
    if planes_out is None:
        _lib.check(lib.get_graph_aggregate_f32(adj.data_ptr(), x.data_ptr(), _ptr(keep), out.data_ptr(), G, N, H,
                                               int(transpose), int(accumulate), _stream()), "get_graph_aggregate_f32")
    else:
        _lib.check(lib.get_graph_aggregate_bp(adj.data_ptr(), x.data_ptr(), _ptr(keep), _ptr(out), planes_out.ptr, planes_out.ld,
                                              planes_out.plane_stride, planes_out.nplanes, int(pad_one), G, N, H,
                                              int(transpose), int(accumulate), _stream()), "get_graph_aggregate_bp")
    return out


def rowdot(feat2d: torch.Tensor, w: torch.Tensor, drop_p: float = 0.0, seed: int = 0) -> torch.Tensor:
    """(M,) = dropout(feat2d) . w -- the projection of the GSL scorer GGNN(H -> 1) (wrapper.py:158,167,191)."""
    _chk_f32(feat2d, "feat"); _chk_f32(w, "w")
    M, H = feat2d.shape
    out = torch.empty((1, M), dtype=torch.float32, device=feat2d.device)
    _lib.check(_lib.load().get_rowdot_f32(feat2d.data_ptr(), w.data_ptr(), M, H, float(drop_p), seed & 0xFFFFFFFF, out.data_ptr(),
                                          _stream()), "get_rowdot_f32")
    return out


def gsl_fused(adj, feat, wp, gate, k, drop_p=0.0, seed_scorer=0, seed_layer2=0, want_score=True, planes_n: int = 0,
              sp_parts: Optional[torch.Tensor] = None):
    """Fused scorer -> top-k -> refined aggregation. Returns (score (G,N) | None, keep (G,N) uint8, out): out is fp32
    (G,N,H), or with planes_n > 0 the refined aggregation as bf16 Planes (G*N, H) for the layer-2 projection.
    adj: NeighborLists (the list kernel: default) or a dense tensor (lists are built here).
    sp_parts (n, G*N): the scorer projection dropout_s(feat) . wp precomputed as partial sums (by-product of the GEMM that
    wrote feat); computed here by a row-dot kernel when absent. Shapes the list kernel does not cover (N > 256, H % 4)
    take the dense one-CTA-per-graph kernel."""
    lib = _lib.load()
    _chk_f32(feat, "feat"); _chk_f32(wp, "wp"); _chk_f32(gate, "gate")
    G, N, H = feat.shape
    assert tuple(adj.shape) == (G, N, N) and feat.is_contiguous()
    assert wp.numel() == H and wp.is_contiguous() and gate.numel() == 12 and gate.is_contiguous()
    score = torch.empty((G, N), dtype=torch.float32, device=feat.device) if want_score else None
    keep = torch.empty((G, N), dtype=torch.uint8, device=feat.device)
    out = alloc_planes(planes_n, G * N, H, feat.device) if planes_n else torch.empty_like(feat)
    adj = as_lists(adj)
    mode = "lists" if adj.usable(H) else "dense"
    if mode != "dense" and sp_parts is None:
        sp_parts = rowdot(feat.view(G * N, H), wp, drop_p, seed_scorer)
    if mode == "dense":
        sp_parts = None
    rec = (adj, feat, wp, gate, int(k), float(drop_p), seed_scorer & 0xFFFFFFFF, seed_layer2 & 0xFFFFFFFF, score, keep, out, sp_parts,
           mode)
    if PROFILE_GSL_ARGS is not None:
        PROFILE_GSL_ARGS.append(rec)
    prof = PROFILE_GSL_EVENTS
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _gsl_launch(*rec)
    if prof is not None:
        e1.record()
        prof.append((e0, e1, G))
    return score, keep, out


def _gsl_launch(adj, feat, wp, gate, k, drop_p, s1, s2, score, keep, out, sp_parts=None, mode="dense"):
    lib = _lib.load()
    G, N, H = feat.shape
    pl = out if isinstance(out, Planes) else None
    plane_args = (None if pl else out.data_ptr(), pl.ptr if pl else None, pl.ld if pl else 0, pl.plane_stride if pl else 0,
                  pl.nplanes if pl else 0)
    if mode == "lists":
        assert sp_parts.is_contiguous() and sp_parts.shape[1] == G * N
        _lib.check(lib.get_gsl_gather(*adj.pointers(False), feat.data_ptr(), sp_parts.data_ptr(), sp_parts.shape[0],
                                      gate.data_ptr(), G, N, H, k, drop_p, s2, _ptr(score), keep.data_ptr(), *plane_args, _stream()),
                   "get_gsl_gather")
        return
    adj = dense_adj(adj)
    if pl is not None:
        _lib.check(lib.get_gsl_fused_bp(adj.data_ptr(), feat.data_ptr(), wp.data_ptr(), gate.data_ptr(), G, N, H, k, drop_p, s1, s2,
                                        _ptr(score), keep.data_ptr(), None, pl.ptr, pl.ld, pl.plane_stride, pl.nplanes,
                                        _stream()), "get_gsl_fused_bp")
    else:
        _lib.check(lib.get_gsl_fused_f32(adj.data_ptr(), feat.data_ptr(), wp.data_ptr(), gate.data_ptr(), G, N, H, k, drop_p, s1,
                                         s2, _ptr(score), keep.data_ptr(), out.data_ptr(), _stream()), "get_gsl_fused_f32")


def gsl_fused_replay(rec):
    """Re-issue one recorded fused GSL launch (PROFILE_GSL_ARGS entry) with the same inputs and output buffers."""
    _gsl_launch(*rec)
    return rec[1].shape[0]


def gsl_mask_adj(adj, score, k):
    lib = _lib.load()
    _chk_f32(adj, "adj"); _chk_f32(score, "score")
    G, N, _ = adj.shape
    adj = adj.contiguous()
    score = score.reshape(G, N).contiguous()
    out = torch.empty_like(adj)
    keep = torch.empty((G, N), dtype=torch.uint8, device=adj.device)
    _lib.check(lib.get_gsl_mask_adj_f32(adj.data_ptr(), score.data_ptr(), G, N, int(k), out.data_ptr(),
                                        keep.data_ptr(), _stream()), "get_gsl_mask_adj_f32")
    return out, keep


def dropout_mask(numel: int, p: float, seed: int, device) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty((numel,), dtype=torch.float32, device=device)
    _lib.check(lib.get_dropout_mask_f32(out.data_ptr(), numel, float(p), seed & 0xFFFFFFFF, _stream()),
               "get_dropout_mask_f32")
    return out


def colsum(a2d: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _chk_f32(a2d, "a")
    M, N = a2d.shape
    out = torch.empty((N,), dtype=torch.float32, device=a2d.device)
    ws = torch.empty((int(lib.get_colsum_workspace_floats(M, N)),), dtype=torch.float32, device=a2d.device)
    _lib.check(lib.get_colsum_f32(a2d.data_ptr(), _ld(a2d), M, N, out.data_ptr(), ws.data_ptr(), _stream()),
               "get_colsum_f32")
    return out


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """(G,P,D) -> (G*P, D) view with a uniform row stride (copies only if the layout forbids it)."""
    G, P, D = t.shape
    if t.stride(2) == 1 and t.stride(0) == P * t.stride(1):
        return t.as_strided((G * P, D), (t.stride(1), 1), t.storage_offset())
    return t.contiguous().view(G * P, D)


# =================================================================================================
# GGNN layer (reference Models/BiDAF/wrapper.py:174-208; backward per SURVEY.md Appendix A.1)
# =================================================================================================
def _grad_target(param: torch.Tensor, shape=None):
    """(destination tensor, accumulate?, returned gradient) for a weight / bias gradient: the registered sink view
    (accumulated in place, nothing returned to autograd) or a fresh tensor."""
    sink = sink_of(param)
    if sink is not None:
        return sink, True, None
    t = torch.empty(tuple(param.shape) if shape is None else shape, dtype=torch.float32, device=param.device)
    return t, False, t


def _ggnn_packs(H, Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1, gs):
    """Packed B operands of one GGNN layer. Activation buffers keep [x | a | r*x] side by side (column blocks of pitch Hp =
    round_up(H + 1, 8): room for the ones column of the `a` block), the gate-gradient buffer [dz' | dr' | dh']; the packs
    use the same column blocks, and every contraction runs over exact-width (K = H) segments of them:
      zr : rows [z: 0..H) [r: gs..gs+H), col blocks [x: Wz1|Wr1][a: Wz0|Wr0]        A segments x, a
      h  : rows 0..H, col blocks [a: Wh0][r*x: Wh1]                                  A segments a, r*x
      da : rows 0..H, col blocks [dz': Wz0^T][dr': Wr0^T][dh': Wh0^T]                A segments dz', dr', dh'
      dx : rows 0..H, col blocks [dz': Wz1^T][dr': Wr1^T]                            A segments dz', dr'
      drx: Wh1^T; p: Wp; pT: Wp^T."""
    Hp = round_up(H + 1, 8)
    key = ("ggnn", gs) + tuple(planes._view_key(w) for w in (Wp, Wz0, Wz1, Wr0, Wr1, Wh0, Wh1))

    def b_zr():
        return 2 * gs, 2 * Hp, [(Wz1, 0, 0), (Wz0, 0, Hp), (Wr1, gs, 0), (Wr0, gs, Hp)], 2 * gs, [(bz0, bz1, 0), (br0, br1, gs)]

    def b_h():
        return H, 2 * Hp, [(Wh0, 0, 0), (Wh1, 0, Hp)], Hp, [(bh0, bh1, 0)]

    def b_da():
        return H, 3 * Hp, [(Wz0.t(), 0, 0), (Wr0.t(), 0, Hp), (Wh0.t(), 0, 2 * Hp)], 0, []

    def b_dx():
        return H, 2 * Hp, [(Wz1.t(), 0, 0), (Wr1.t(), 0, Hp)], 0, []
    return {"zr": lambda: get_pack(key + ("zr",), b_zr), "h": lambda: get_pack(key + ("h",), b_h),
            "da": lambda: get_pack(key + ("da",), b_da), "dx": lambda: get_pack(key + ("dx",), b_dx),
            "drx": lambda: pack_of(Wh1.t()), "p": lambda: pack_of(Wp), "pT": lambda: pack_of(Wp.t())}


class GGNNLayerFn(torch.autograd.Function):
    """out = GGNN(adj', x_in).  x_in is either `feat` (G,N,Din) or rows `ids` (G,N) of the frozen `table`.
    keep (G,N) uint8 restricts the adjacency to edges with a kept endpoint (GSL); pre_agg (bf16 planes tensor of
    adj' @ drop(x_in), from the fused GSL kernel) replaces the aggregation of the projected features by linearity.
    Returns (out (G,N,H) fp32, out planes (P, G*N, Hp) bf16 or an empty tensor): the planes feed the next contraction."""

    @staticmethod
    def forward(ctx, adj, feat, table, ids, keep, pre_agg, p_drop, seed,
                Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1, exact_fwd=False, out_planes=0,
                rd_w=None, rd_p=0.0, rd_seed=0):
        adj = as_lists(adj)                 # built here when the caller passed a dense adjacency
        G, N = adj.G, adj.N
        M = G * N
        H, Din = Wp.shape
        Hp = round_up(H + 1, 8)
        dev = adj.device
        f32 = dict(dtype=torch.float32, device=dev)
        mode = gemm_mode(exact_fwd)
        bn = planes.tile_n(M, H, mode)
        gs = round_up(Hp, bn)
        pk = _ggnn_packs(H, Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1, gs)
        # projection input (embedding gather and / or dropout) written ONCE as planes: read by the forward projection and
        # by the weight gradient dWp
        if feat is not None:
            xd = rows_gather_dropout_planes(_rows2d(feat), None, M, p_drop, seed, mode)
        else:
            _chk_f32(table, "table")
            xd = rows_gather_dropout_planes(table, ids.reshape(-1).to(torch.int64).contiguous(), M, p_drop, seed, mode)
        xar = alloc_planes(mode, M, 3 * Hp, dev)          # [x | a | r*x]
        xP, aP, rxP = xar.view_cols(0, H), xar.view_cols(Hp, H), xar.view_cols(2 * Hp, H)
        x = torch.empty((M, H), **f32)
        gemm_bp([(xd, pk["p"]().planes, Din)], M, H, mode=mode, C=x, planes_out=xP)
        if pre_agg is not None:
            gemm_bp([(Planes(pre_agg, Din), pk["p"]().planes, Din)], M, H, mode=mode, planes_out=aP, pad_one=True)
        else:
            graph_aggregate(adj, x.view(G, N, H), keep, planes_out=aP, pad_one=True, want_f32=False)
        z = torch.empty((M, H), **f32)
        r = torch.empty((M, H), **f32)
        h = torch.empty((M, H), **f32)
        out = torch.empty((M, H), **f32)
        pzr = pk["zr"]()
        gemm_bp([(xP, pzr.planes.view_cols(0, H), H), (aP, pzr.planes.view_cols(Hp, H), H)], M, 2 * gs, mode=mode,
                epilogue=BPE_ZR, C=z, out1=r, bias=pzr.bias, aux0=x, planes_out=rxP, zr=(gs, H), tn=bn)
        ph = pk["h"]()
        op = alloc_planes(out_planes, M, H, dev) if out_planes else None
        # row-dot by-product: the scorer projection dropout_s(out) . w_p in partial sums per N tile half (the fused GSL kernel
        # adds them up), so that kernel never needs whole feature rows
        rd = None
        if rd_w is not None:
            n_cols = round_up(H, 8) if op is not None else H       # columns the launch tiles (planes add the padding)
            rd = torch.empty((2 * ((n_cols + bn - 1) // bn), M), dtype=torch.float32, device=dev)
        gemm_bp([(aP, ph.planes.view_cols(0, H), H), (rxP, ph.planes.view_cols(Hp, H), H)], M, H, mode=mode,
                epilogue=BPE_TANH_BLEND, C=out, out1=h, bias=ph.bias, aux0=z, aux1=x, planes_out=op, tn=bn,
                rowdot=(rd_w, rd, rd_p, rd_seed) if rd is not None else None)
        ctx.save_for_backward(keep, xd.t, xar.t, x, z, r, h, Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1)
        ctx.adj = adj
        ctx.p_drop, ctx.seed, ctx.dims, ctx.has_feat, ctx.gs = p_drop, seed, (G, N, H, Din), feat is not None, gs
        op_t = op.t if op is not None else torch.empty(0, dtype=torch.bfloat16, device=dev)
        rd_t = rd if rd is not None else torch.empty(0, dtype=torch.float32, device=dev)
        ctx.mark_non_differentiable(op_t, rd_t)
        return out.view(G, N, H), op_t, rd_t

    @staticmethod
    def backward(ctx, dout, _dplanes, _drd):
        (keep, xd_t, xar_t, x, z, r, h, Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1) = ctx.saved_tensors
        adj = ctx.adj
        G, N, H, Din = ctx.dims
        M = G * N
        Hp = round_up(H + 1, 8)
        dev = dout.device
        lib = _lib.load()
        mode = gemm_mode(False)
        pk = _ggnn_packs(H, Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1, ctx.gs)
        xd, xar = Planes(xd_t, Din), Planes(xar_t, 3 * Hp)
        dout = dout.contiguous().view(M, H)
        # gate gradients as planes, side by side [dz' | dr' | dh']: the contractions below and the weight gradients that
        # share an activation operand are ONE tensor-core launch each
        dg = alloc_planes(mode, M, 3 * Hp, dev)
        dx = torch.empty((M, H), dtype=torch.float32, device=dev)
        _lib.check(lib.get_ggnn_gate_bwd_bp(dout.data_ptr(), z.data_ptr(), h.data_ptr(), x.data_ptr(), M, H, dg.ptr, dg.ld,
                                            dg.plane_stride, mode, 0, 2 * Hp, dx.data_ptr(), _stream()), "get_ggnn_gate_bwd_bp")
        # d(rx) = dh' @ Wh1 ; dr' = d(rx)*x*r*(1-r) -> planes ; dx += d(rx)*r
        gemm_bp([(dg.view_cols(2 * Hp, H), pk["drx"]().planes, H)], M, H, mode=mode, epilogue=BPE_DGATE_R, aux0=x, aux1=r,
                out1=dx, planes_out=dg.view_cols(Hp, H))
        # da = dz'@Wz0 + dr'@Wr0 + dh'@Wh0
        da = torch.empty((M, H), dtype=torch.float32, device=dev)
        dzP, drP, dhP = dg.view_cols(0, H), dg.view_cols(Hp, H), dg.view_cols(2 * Hp, H)
        pda, pdx = pk["da"]().planes, pk["dx"]().planes
        gemm_bp([(dzP, pda.view_cols(0, H), H), (drP, pda.view_cols(Hp, H), H), (dhP, pda.view_cols(2 * Hp, H), H)], M, H,
                mode=mode, C=da)
        # dx += dz'@Wz1 + dr'@Wr1 + adj'^T @ da ; the aggregation kernel also writes the planes of the final dx
        gemm_bp([(dzP, pdx.view_cols(0, H), H), (drP, pdx.view_cols(Hp, H), H)], M, H, mode=mode, C=dx, accumulate=True)
        dxP = alloc_planes(mode, M, H, dev)
        graph_aggregate(adj, da.view(G, N, H), keep, out=dx.view(G, N, H), transpose=True, accumulate=True, planes_out=dxP)
        need = ctx.needs_input_grad
        grads = [None] * 26
        # order of inputs: ... 8:Wp 9:Wz0 10:bz0 11:Wz1 12:bz1 13:Wr0 14:br0 15:Wr1 16:br1 17:Wh0 18:bh0 19:Wh1 20:bh1
        xP, aP1, rxP = xar.view_cols(0, H), xar.view_cols(Hp, Hp), xar.view_cols(2 * Hp, H)
        if need[9] or need[13] or need[17] or need[10] or need[14] or need[18]:
            # [dWz0; dWr0; dWh0 | db] = [dz'|dr'|dh']^T [a | 1]: the ones column of the `a` block yields the bias gradients
            dsts = []
            for i_w, i_b0, i_b1, r0, W, b0, b1 in ((9, 10, 12, 0, Wz0, bz0, bz1), (13, 14, 16, Hp, Wr0, br0, br1),
                                                   (17, 18, 20, 2 * Hp, Wh0, bh0, bh1)):
                t, acc, g = _grad_target(W)
                assert acc == (sink_of(Wz0) is not None), "grad sinks must cover all parameters of a GGNN layer or none"
                dsts.append((t, r0, H, 0, H)); grads[i_w] = g
                for i_b, bb in ((i_b0, b0), (i_b1, b1)):
                    t, _, g = _grad_target(bb)
                    dsts.append((t, r0, H, H, 1)); grads[i_b] = g
            wgrad_bp(dg.view_cols(0, 3 * Hp).T(), aP1.T(), 3 * Hp, Hp, M, mode, dsts, accumulate=sink_of(Wz0) is not None)
        if need[11] or need[15]:
            dsts = []
            for i_w, r0, W in ((11, 0, Wz1), (15, Hp, Wr1)):
                t, acc, g = _grad_target(W)
                dsts.append((t, r0, H, 0, H)); grads[i_w] = g
            wgrad_bp(dg.view_cols(0, 2 * Hp).T(), xP.T(), 2 * Hp, H, M, mode, dsts, accumulate=sink_of(Wz1) is not None)
        if need[19]:
            t, acc, g = _grad_target(Wh1)
            grads[19] = g
            wgrad_bp(dg.view_cols(2 * Hp, H).T(), rxP.T(), H, H, M, mode, [(t, 0, H, 0, H)], accumulate=acc)
        if need[8]:
            t, acc, g = _grad_target(Wp)
            grads[8] = g
            wgrad_bp(dxP.T(), xd.T(), H, Din, M, mode, [(t, 0, H, 0, Din)], accumulate=acc)
        if ctx.has_feat and need[1]:
            dfeat = torch.empty((M, Din), dtype=torch.float32, device=dev)
            gemm_bp([(dxP, pk["pT"]().planes, H)], M, Din, mode=mode, C=dfeat,
                    drop_out=(ctx.p_drop, ctx.seed) if ctx.p_drop > 0 else None)
            grads[1] = dfeat.view(G, N, Din)
        return tuple(grads)


class GGNNLayerSimtFn(torch.autograd.Function):
    """Exact-fp32 SIMT implementation for small / odd shapes (M < 128 rows, feature sizes not multiples of 4, e.g. the
    stand-alone GGNN(H -> 1) surface): the same computation as GGNNLayerFn without tensor cores.
    out = GGNN(adj', x_in).  x_in is either `feat` (G,N,Din) or rows `ids` (G,N) of the frozen `table`.
    keep (G,N) uint8 restricts the adjacency to edges with a kept endpoint (GSL); pre_agg = adj' @ drop(x_in)
    (from the fused GSL kernel) replaces the aggregation of the projected features by linearity."""

    @staticmethod
    def forward(ctx, adj, feat, table, ids, keep, pre_agg, p_drop, seed,
                Wp, Wz0, bz0, Wz1, bz1, Wr0, br0, Wr1, br1, Wh0, bh0, Wh1, bh1, exact_fwd=False):
        adj = as_lists(adj)
        G, N = adj.G, adj.N
        M = G * N
        H, Din = Wp.shape
        dev = adj.device
        f32 = dict(dtype=torch.float32, device=dev)
        x = torch.empty((M, H), **f32)
        # the projection input (embedding gather and / or dropout) is materialised once: the forward projection and the
        # weight gradient dWp then both read plain (M, Din) tiles through TMA
        if feat is not None:
            feat2d = _rows2d(feat)
            xd = rows_gather_dropout(feat2d, None, M, p_drop, seed) if p_drop > 0 else feat2d
        else:
            _chk_f32(table, "table")
            rowidx = ids.reshape(-1).to(torch.int64).contiguous()
            xd = rows_gather_dropout(table, rowidx, M, p_drop, seed)
        gemm([(xd, Wp)], x)
        a = torch.empty((M, H), **f32)
        if pre_agg is not None:
            gemm([(_rows2d(pre_agg), Wp)], a)
        else:
            graph_aggregate(adj, x.view(G, N, H), keep, out=a.view(G, N, H))
        z = torch.empty((M, H), **f32)
        r = torch.empty((M, H), **f32)
        rx = torch.empty((M, H), **f32)
        h = torch.empty((M, H), **f32)
        out = torch.empty((M, H), **f32)
        gemm([(a, Wz0), (x, Wz1)], z, epilogue=EPI_SIGMOID, bias0=bz0, bias1=bz1)
        gemm([(a, Wr0), (x, Wr1)], r, epilogue=EPI_SIGMOID, bias0=br0, bias1=br1, aux0=x, out1=rx)
        gemm([(a, Wh0), (rx, Wh1)], out, epilogue=EPI_TANH_BLEND, bias0=bh0, bias1=bh1, aux0=z, aux1=x, out1=h)
        ctx.save_for_backward(xd, keep, x, a, z, r, rx, h, Wp, Wz0, Wz1, Wr0, Wr1, Wh0, Wh1)
        ctx.adj = adj
        ctx.p_drop, ctx.seed, ctx.dims, ctx.has_feat = p_drop, seed, (G, N, H, Din), feat is not None
        return out.view(G, N, H)

    @staticmethod
    def backward(ctx, dout):
        xd, keep, x, a, z, r, rx, h, Wp, Wz0, Wz1, Wr0, Wr1, Wh0, Wh1 = ctx.saved_tensors
        adj = ctx.adj
        G, N, H, Din = ctx.dims
        M = G * N
        dev = dout.device
        f32 = dict(dtype=torch.float32, device=dev)
        lib = _lib.load()
        dout = dout.contiguous().view(M, H)
        # gate gradients live side by side in one (M, 3H) buffer [dz' | dr' | dh'] so that the weight gradients that
        # share an activation operand are ONE contraction each: [dz'|dr'|dh']^T a and [dz'|dr']^T x
        dg = torch.empty((M, 3 * H), **f32)
        dzp, drp, dhp = dg[:, :H], dg[:, H:2 * H], dg[:, 2 * H:]
        dx = torch.empty((M, H), **f32)
        da = torch.empty((M, H), **f32)
        _lib.check(lib.get_ggnn_gate_bwd_f32(dout.data_ptr(), z.data_ptr(), h.data_ptr(), x.data_ptr(), M, H, 3 * H,
                                             dhp.data_ptr(), dzp.data_ptr(), dx.data_ptr(), _stream()),
                   "get_ggnn_gate_bwd_f32")
        # d(rx) = dhp @ Wh1 ; drp = d(rx)*x*r*(1-r) ; dx += d(rx)*r
        gemm([(dhp, Wh1.t())], drp, epilogue=EPI_DGATE_R, aux0=x, aux1=r, out1=dx)
        # da = dhp@Wh0 + dzp@Wz0 + drp@Wr0
        gemm([(dhp, Wh0.t()), (dzp, Wz0.t()), (drp, Wr0.t())], da)
        # dx += dzp@Wz1 + drp@Wr1 + adj'^T @ da
        gemm([(dzp, Wz1.t()), (drp, Wr1.t())], dx, accumulate=True)
        graph_aggregate(adj, da.view(G, N, H), keep, out=dx.view(G, N, H), transpose=True, accumulate=True)
        need = ctx.needs_input_grad
        grads = [None] * 22
        # order of inputs: ... 8:Wp 9:Wz0 10:bz0 11:Wz1 12:bz1 13:Wr0 14:br0 15:Wr1 16:br1 17:Wh0 18:bh0 19:Wh1 20:bh1
        if need[9] or need[13] or need[17]:
            wa = torch.empty((3 * H, H), **f32)                      # [dWz0; dWr0; dWh0]
            gemm([(dg.t(), a.t())], wa)
            grads[9], grads[13], grads[17] = wa[:H], wa[H:2 * H], wa[2 * H:]
        if need[11] or need[15]:
            wx = torch.empty((2 * H, H), **f32)                      # [dWz1; dWr1]
            gemm([(dg[:, :2 * H].t(), x.t())], wx)
            grads[11], grads[15] = wx[:H], wx[H:]
        if need[19]:
            w = torch.empty((H, H), **f32)
            gemm([(dhp.t(), rx.t())], w)
            grads[19] = w
        if any(need[i] for i in (10, 12, 14, 16, 18, 20)):
            gb = colsum(dg)                                          # [dbz | dbr | dbh]
            grads[10] = grads[12] = gb[:H]
            grads[14] = grads[16] = gb[H:2 * H]
            grads[18] = grads[20] = gb[2 * H:]
        if need[8]:
            # dWp^T (Din,H) = Xd^T @ dx on the materialised projection input
            wT = torch.empty((Din, H), **f32)
            gemm([(xd.t(), dx.t())], wT)
            grads[8] = wT.t()
        if ctx.has_feat and need[1]:
            dfeat = torch.empty((M, Din), **f32)
            if ctx.p_drop > 0:
                gemm([(dx, Wp.t())], dfeat, epilogue=EPI_DROPOUT_OUT, drop_out_p=ctx.p_drop, drop_out_seed=ctx.seed)
            else:
                gemm([(dx, Wp.t())], dfeat)
            grads[1] = dfeat.view(G, N, Din)
        return tuple(grads)


def layer_uses_tc(M: int, H: int, Din: int) -> bool:
    """Does a GGNN layer of these sizes run on the tensor-core plane path (else: exact SIMT path)?"""
    return bool(TC_ENABLED and M >= 128 and H % 4 == 0 and Din % 4 == 0 and H >= 8 and Din >= 8)


def ggnn_layer(adj, feat, table, ids, keep, pre_agg, p_drop, seed, params: Sequence[torch.Tensor], exact_fwd: bool = False,
               out_planes: int = 0, rowdot_of=None):
    """exact_fwd=True: the forward contractions stay fp32-exact in every precision mode (feat_prop1: GSL top-k chain).
    rowdot_of = (w (H,), p, seed): also return the partial row dots dropout(out; p, seed) . w (the GSL scorer projection).
    Returns (out, planes tensor of out | None, row-dot partial sums | None). Large regular shapes run on the tensor cores;
    the rest on the SIMT path."""
    H, Din = params[0].shape
    M = adj.shape[0] * adj.shape[1]
    if layer_uses_tc(M, H, Din):
        rd_w, rd_p, rd_seed = rowdot_of if rowdot_of is not None else (None, 0.0, 0)
        out, op, rd = GGNNLayerFn.apply(adj, feat, table, ids, keep, pre_agg, float(p_drop), int(seed), *params, bool(exact_fwd),
                                        int(out_planes), rd_w, float(rd_p), int(rd_seed))
        return out, (op if out_planes else None), (rd if rowdot_of is not None else None)
    if isinstance(pre_agg, torch.Tensor) and pre_agg.dtype == torch.bfloat16:
        pre_agg = Planes(pre_agg, Din).to_float().view(adj.shape[0], adj.shape[1], Din)
    return GGNNLayerSimtFn.apply(adj, feat, table, ids, keep, pre_agg, float(p_drop), int(seed), *params, bool(exact_fwd)), None, None


# =================================================================================================
# ConcatNotEqualSelfAtt / MultiHeadSelfAttentionICLR2017Extend
# (reference thirdparty/two_branches_attention.py:121-148, thirdparty/self_attention.py:75-100)
# =================================================================================================
class ConcatAttFn(torch.autograd.Function):
    """right_planes: optional bf16 plane tensor of `right` (written by the producing GEMM's epilogue)."""

    @staticmethod
    def forward(ctx, left, right, mask_u8, W1, W2, right_planes=None):
        lib = _lib.load()
        G, P, Dr = right.shape
        H = W1.shape[0]
        X = W1.shape[1] - Dr
        Cn = W2.shape[0]
        dev = right.device
        f32 = dict(dtype=torch.float32, device=dev)
        right2d = _rows2d(right)
        W1 = W1.contiguous()
        W2 = W2.contiguous()
        rP = Planes(right_planes, Dr) if right_planes is not None else None
        if rP is None and TC_ENABLED and G * P >= 128 and Dr % 4 == 0 and H % 4 == 0:
            rP = to_planes(right2d, gemm_mode(False))          # converted once: forward projection + weight gradient
        t = torch.empty((G * P, H), **f32)
        if left is not None:
            lp = torch.empty((G, H), **f32)
            gemm([(left.contiguous(), W1[:, :X])], lp, tc=True)
            gemm([(rP if rP is not None else right2d, W1[:, X:])], t, epilogue=EPI_TANH_ROWGROUP, aux0=lp, group_rows=P, tc=True)
        else:
            gemm([(rP if rP is not None else right2d, W1)], t, epilogue=EPI_TANH, tc=True)
        att = torch.empty((G, P, Cn), **f32)
        pooled = torch.empty((G, Dr, Cn), **f32)
        mask_u8 = mask_u8.contiguous()
        _lib.check(lib.get_att_pool_fwd_f32(t.data_ptr(), right2d.data_ptr(), _ld(right2d), W2.data_ptr(),
                                            mask_u8.data_ptr(), G, P, H, Dr, Cn, att.data_ptr(), pooled.data_ptr(),
                                            Dr * Cn, _stream()), "get_att_pool_fwd_f32")
        ctx.save_for_backward(left, right2d, W1, W2, t, att, rP.t if rP is not None else None)
        ctx.dims = (G, P, H, Dr, Cn, X)
        return pooled, att

    @staticmethod
    def backward(ctx, d_pooled, d_att):
        lib = _lib.load()
        left, right2d, W1, W2, t, att, rP_t = ctx.saved_tensors
        G, P, H, Dr, Cn, X = ctx.dims
        dev = right2d.device
        f32 = dict(dtype=torch.float32, device=dev)
        if d_pooled is None:
            d_pooled = torch.zeros((G, Dr, Cn), **f32)
        d_pooled = d_pooled.contiguous()
        if d_att is not None:
            d_att = d_att.contiguous()
        de = torch.empty((G * P, Cn), **f32)
        du_sum = torch.empty((G, H), **f32)
        dright = torch.empty((G * P, Dr), **f32)
        use_tc = rP_t is not None
        du = duP = None
        if use_tc:
            # du leaves the kernel as bf16 planes: it is only ever the operand of the two contractions below
            duP = alloc_planes(gemm_mode(False), G * P, H, dev)
            _lib.check(lib.get_att_pool_bwd_bp(t.data_ptr(), right2d.data_ptr(), _ld(right2d), W2.data_ptr(), att.data_ptr(),
                                               d_pooled.data_ptr(), Dr * Cn, _ptr(d_att), G, P, H, Dr, Cn, de.data_ptr(),
                                               duP.ptr, duP.ld, duP.plane_stride, duP.nplanes, du_sum.data_ptr(),
                                               dright.data_ptr(), Dr, 0, _stream()), "get_att_pool_bwd_bp")
        else:
            du = torch.empty((G * P, H), **f32)
            _lib.check(lib.get_att_pool_bwd_f32(t.data_ptr(), right2d.data_ptr(), _ld(right2d), W2.data_ptr(),
                                                att.data_ptr(), d_pooled.data_ptr(), Dr * Cn, _ptr(d_att),
                                                G, P, H, Dr, Cn, de.data_ptr(), du.data_ptr(), du_sum.data_ptr(),
                                                dright.data_ptr(), Dr, 0, _stream()), "get_att_pool_bwd_f32")
        need = ctx.needs_input_grad
        dleft = dW1 = dW2 = None
        W1R = W1[:, X:]
        if need[1]:
            gemm([(duP if duP is not None else du, W1R.t())], dright, accumulate=True, tc=True)
        if need[4]:
            dW2 = torch.empty((Cn, H), **f32)
            gemm([(de.t(), t.t())], dW2)
        if need[3]:
            sink = sink_of(W1)
            dW1 = sink if sink is not None else torch.empty((H, X + Dr), **f32)
            if use_tc:
                gemm([(duP, Planes(rP_t, Dr))], dW1[:, X:], tc=True, presplit=False, accumulate=sink is not None)
            else:
                gemm([(du.t(), right2d.t())], dW1[:, X:], accumulate=sink is not None)
            if left is not None:
                gemm([(du_sum.t(), left.contiguous().t())], dW1[:, :X], accumulate=sink is not None)
            if sink is not None:
                dW1 = None
        if left is not None and need[0]:
            dleft = torch.empty((G, X), **f32)
            gemm([(du_sum, W1[:, :X].t())], dleft, tc=True)
        return dleft, (dright.view(G, P, Dr) if need[1] else None), None, dW1, dW2, None


def concat_att(left, right, mask, W1, W2, right_planes=None):
    mask_u8 = mask if mask.dtype == torch.uint8 else (mask != 0).to(torch.uint8)
    return ConcatAttFn.apply(left, right, mask_u8, W1, W2, right_planes)


# =================================================================================================
# glue ops with custom kernels: linear, segment expand / pad, masked mean, cross entropy
# =================================================================================================
class LinearFn(torch.autograd.Function):
    """y = x @ W^T + b (reference nn.Linear of the output MLP, graph_based_semantic_structure.py:69-74,121)."""

    @staticmethod
    def forward(ctx, x, W, b):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y = torch.empty((x2.shape[0], W.shape[0]), dtype=torch.float32, device=x.device)
        gemm([(x2, W.contiguous())], y, bias0=b)
        ctx.save_for_backward(x2, W)
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            gemm([(dy2, W.t())], dx)
            dx = dx.view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dst, acc, dW = _grad_target(W)
            gemm([(dy2.t(), x2.t())], dst, accumulate=acc)
        if ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dW, db


def linear(x, W, b=None):
    return LinearFn.apply(x, W, b)


class SegmentExpandFn(torch.autograd.Function):
    """`_pad_left_tensor` (reference basic_fc_model.py:80-92): out[r] = src[seg_of_row[r]]."""

    @staticmethod
    def forward(ctx, src, seg_of_row, offsets):
        lib = _lib.load()
        src = src.contiguous()
        R, W = seg_of_row.shape[0], src.shape[1]
        out = torch.empty((R, W), dtype=torch.float32, device=src.device)
        _lib.check(lib.get_rows_gather_f32(src.data_ptr(), W, seg_of_row.data_ptr(), R, W, out.data_ptr(), W,
                                           _stream()), "get_rows_gather_f32")
        ctx.save_for_backward(offsets)
        ctx.S = src.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (offsets,) = ctx.saved_tensors
        dout = dout.contiguous()
        W = dout.shape[1]
        dsrc = torch.empty((ctx.S, W), dtype=torch.float32, device=dout.device)
        _lib.check(lib.get_segment_sum_f32(dout.data_ptr(), W, offsets.data_ptr(), ctx.S, W, dsrc.data_ptr(), W,
                                           _stream()), "get_segment_sum_f32")
        return dsrc, None, None


class SegmentPadFn(torch.autograd.Function):
    """`_pad_right_tensor` (reference basic_fc_model.py:94-121) fused with the article-source concat
    (graph_based_semantic_structure.py:157-171): out (S_slots, W+E) zero buffer, out[slot_of_row[r], :W] = src[r],
    out[:, W:] = extra."""

    @staticmethod
    def forward(ctx, src, slot_of_row, n_slots, extra):
        lib = _lib.load()
        src = src.contiguous()
        R, W = src.shape
        E = 0 if extra is None else extra.shape[1]
        out = torch.zeros((n_slots, W + E), dtype=torch.float32, device=src.device)
        _lib.check(lib.get_rows_scatter_f32(src.data_ptr(), W, slot_of_row.data_ptr(), R, W, out.data_ptr(),
                                            W + E, _stream()), "get_rows_scatter_f32")
        if extra is not None:
            out[:, W:].copy_(extra)
        ctx.save_for_backward(slot_of_row)
        ctx.W, ctx.E = W, E
        return out

    @staticmethod
    def backward(ctx, dbuf):
        lib = _lib.load()
        (slot_of_row,) = ctx.saved_tensors
        R, W = slot_of_row.shape[0], ctx.W
        dbuf = dbuf.contiguous()
        dsrc = None
        if ctx.needs_input_grad[0]:
            dsrc = torch.empty((R, W), dtype=torch.float32, device=dbuf.device)
            _lib.check(lib.get_rows_gather_f32(dbuf.data_ptr(), W + ctx.E, slot_of_row.data_ptr(), R, W,
                                               dsrc.data_ptr(), W, _stream()), "get_rows_gather_f32")
        dextra = dbuf[:, W:] if (ctx.E and ctx.needs_input_grad[3]) else None
        return dsrc, None, None, dextra


class MaskedMeanFn(torch.autograd.Function):
    """Claim read-out (reference graph_based_semantic_structure.py:145-153)."""

    @staticmethod
    def forward(ctx, h, ids, lens):
        lib = _lib.load()
        G, N, H = h.shape
        h = h.contiguous()
        ids = ids.to(torch.int64).contiguous()
        lens = lens.to(torch.int64).contiguous()
        out = torch.empty((G, H), dtype=torch.float32, device=h.device)
        _lib.check(lib.get_masked_mean_fwd_f32(h.data_ptr(), ids.data_ptr(), lens.data_ptr(), G, N, H,
                                               out.data_ptr(), _stream()), "get_masked_mean_fwd_f32")
        ctx.save_for_backward(ids, lens)
        ctx.dims = (G, N, H)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        ids, lens = ctx.saved_tensors
        G, N, H = ctx.dims
        dout = dout.contiguous()
        dh = torch.empty((G, N, H), dtype=torch.float32, device=dout.device)
        _lib.check(lib.get_masked_mean_bwd_f32(dout.data_ptr(), ids.data_ptr(), lens.data_ptr(), G, N, H,
                                               dh.data_ptr(), _stream()), "get_masked_mean_bwd_f32")
        return dh, None, None


class CrossEntropyFn(torch.autograd.Function):
    """`losses.cross_entroy` (reference losses.py:29-32): mean CE over the claims."""

    @staticmethod
    def forward(ctx, logits, labels):
        lib = _lib.load()
        logits = logits.contiguous()
        labels = labels.to(torch.int64).contiguous()
        B, Cn = logits.shape
        loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        _lib.check(lib.get_cross_entropy_f32(logits.data_ptr(), labels.data_ptr(), B, Cn, loss.data_ptr(),
                                             dlogits.data_ptr(), _stream()), "get_cross_entropy_f32")
        ctx.save_for_backward(dlogits)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None


def cross_entropy(logits, labels):
    return CrossEntropyFn.apply(logits, labels)
