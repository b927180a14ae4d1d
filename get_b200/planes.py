"""bf16 plane tensors, packed weights and the host wrapper of the tcgen05 plane GEMM (include/get_b200.h: get_gemm_bp).

A fp32 value v travels between the kernels of the hot path as up to three bf16 planes (p0 = bf16(v), p1 = bf16(v - p0),
p2 = bf16(v - p0 - p1)); the tensor-core contraction multiplies planes pairwise (mode 1 / 2 / 3 = bf16 / 16-bit / fp32-exact
class). Activations get their planes from the kernel that produces them; weights are packed once per optimizer step.
PyTorch is used for device memory only. Nothing here falls back to torch math.
"""
import ctypes as ct
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import BPE_STORE

_SM_COUNT = 148


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def round_up(v: int, q: int) -> int:
    return (v + q - 1) // q * q


class Planes(object):
    """View of `cols` logical columns starting at column `col0` of a bf16 plane tensor t (P, rows, ld)."""
    __slots__ = ("t", "cols", "col0", "trans")

    def __init__(self, t: torch.Tensor, cols: int, col0: int = 0, trans: int = 0):
        assert t.dtype == torch.bfloat16 and t.dim() == 3 and t.stride(2) == 1 and t.is_cuda
        assert t.stride(1) % 8 == 0 and (t.shape[0] == 1 or t.stride(0) % 8 == 0) and col0 % 8 == 0
        self.t, self.cols, self.col0, self.trans = t, int(cols), int(col0), int(trans)

    @property
    def rows(self) -> int:
        return self.t.shape[1]

    @property
    def nplanes(self) -> int:
        return self.t.shape[0]

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + 2 * self.col0

    @property
    def ld(self) -> int:
        return self.t.stride(1)

    @property
    def plane_stride(self) -> int:
        return self.t.stride(0) if self.t.shape[0] > 1 else self.t.stride(1) * self.t.shape[1]

    def view_cols(self, col0: int, cols: int) -> "Planes":
        return Planes(self.t, cols, self.col0 + col0, self.trans)

    def view_rows(self, r0: int, r1: int) -> "Planes":
        return Planes(self.t[:, r0:r1], self.cols, self.col0, self.trans)

    def T(self) -> "Planes":
        """The same storage read MN-major: logical (cols, rows) operand whose contraction index is the stored row."""
        return Planes(self.t, self.cols, self.col0, 1 - self.trans)

    def to_float(self, nplanes: Optional[int] = None) -> torch.Tensor:
        """Sum of the planes as fp32 (rows, cols) -- tests / debugging only."""
        n = nplanes or self.nplanes
        return self.t[:n, :, self.col0:self.col0 + self.cols].float().sum(0)


def alloc_planes(nplanes: int, rows: int, cols: int, device, ld: Optional[int] = None, zero: bool = False) -> Planes:
    ld = ld or round_up(cols, 8)
    f = torch.zeros if zero else torch.empty
    return Planes(f((nplanes, rows, ld), dtype=torch.bfloat16, device=device), cols)


def to_planes(x: torch.Tensor, nplanes: int, pad_one: bool = False, out: Optional[Planes] = None) -> Planes:
    """fp32 (rows, cols) matrix (unit column stride) -> planes; pad columns are written."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and (x.stride(1) == 1 or x.shape[1] == 1)
    rows, cols = x.shape
    if out is None:
        out = alloc_planes(nplanes, rows, cols, x.device, ld=round_up(cols + (1 if pad_one else 0), 8))
    assert out.rows == rows and out.cols == cols and out.nplanes >= nplanes
    _lib.check(_lib.load().get_to_planes_bf16(x.data_ptr(), x.stride(0) if rows > 1 else max(x.stride(0), cols), rows, cols,
                                              out.ptr, out.ld, out.plane_stride, nplanes, int(pad_one), _stream()),
               "get_to_planes_bf16")
    return out


# =====================================================================================================================
# Packed weights: every B operand of a forward / backward contraction is a bf16 plane tensor (3 planes) assembled from
# blocks of parameters (possibly transposed views, stacked, side by side) plus optional fused bias vectors. All packs of
# the process are refreshed by ONE kernel launch when the weights change (optimizer step / captured training step).
# =====================================================================================================================
class WeightPack(object):
    __slots__ = ("key", "planes", "bias", "blocks", "bias_blocks", "state", "pins")

    def __init__(self, key, rows, cols, blocks, bias_len, bias_blocks, device):
        self.key = key
        self.planes = alloc_planes(3, rows, cols, device, zero=True)      # padding rows / columns stay zero forever
        self.bias = torch.zeros((bias_len,), dtype=torch.float32, device=device) if bias_len else None
        self.blocks = blocks              # [(param view 2-D logical (r, c), row_off, col_off)]
        self.bias_blocks = bias_blocks    # [(b0, b1 | None, off)]
        self.state = None                 # (versions, epoch) the planes were built from
        # pin the source storages: while cached the allocator cannot hand the same address to another tensor
        self.pins = [b[0].untyped_storage() for b in blocks] + [b[0].untyped_storage() for b in bias_blocks]

    def versions(self):
        return tuple(b[0]._version for b in self.blocks) + tuple(b[0]._version for b in self.bias_blocks)

    def jobs(self, first_block: int) -> Tuple[List[_lib.PackJob], int]:
        out = []
        blk = first_block
        P = self.planes
        for (w, r0, c0) in self.blocks:
            j = _lib.PackJob()
            j.src, j.src2, j.ld_r, j.ld_c = w.data_ptr(), None, w.stride(0), w.stride(1)
            j.rows, j.cols = w.shape
            j.dst = P.t.data_ptr() + 2 * (r0 * P.ld + c0)
            j.ld_out, j.plane_stride, j.first_block, j.kind = P.ld, P.plane_stride, blk, 0
            blk += ((w.shape[0] + 31) // 32) * ((w.shape[1] + 31) // 32)
            out.append(j)
        for (b0, b1, off) in self.bias_blocks:
            j = _lib.PackJob()
            j.src, j.src2 = b0.data_ptr(), (b1.data_ptr() if b1 is not None else None)
            j.ld_r, j.ld_c, j.rows, j.cols = 1, 1, b0.numel(), 1
            j.dst = self.bias.data_ptr() + 4 * off
            j.ld_out, j.plane_stride, j.first_block, j.kind = 0, 0, blk, 1
            blk += (b0.numel() + 1023) // 1024
            out.append(j)
        return out, blk


_packs: Dict[Tuple, WeightPack] = {}
_pack_epoch = 0
_pack_table = None        # (device uint8 tensor of get_pack_job[], n_jobs, total_blocks, [packs]) or None when stale
_retired_tables = []      # superseded job tables, kept alive for the graphs that captured them
_PACK_MAX = 1024


def weights_updated(*_a, **_k):
    """Invalidate every pack (optimizer post-step hook; fused optimizers do not move parameter version counters)."""
    global _pack_epoch
    _pack_epoch += 1


def begin_step_capture():
    """A captured training step re-packs the weights inside the graph (they differ at every replay)."""
    global _pack_epoch
    _pack_epoch += 1


def _view_key(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()))


def _run_jobs(jobs: Sequence[_lib.PackJob], blocks: int):
    arr = (_lib.PackJob * len(jobs))(*jobs)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    _lib.check(_lib.load().get_pack_planes_multi(raw.data_ptr(), len(jobs), blocks, _stream()), "get_pack_planes_multi")
    return raw


def prepare_pack_table():
    """(Re)build the device job table covering every pack. Host -> device copy: must run OUTSIDE a stream capture."""
    global _pack_table
    if _pack_table is not None or not _packs:
        return
    packs = list(_packs.values())
    jobs, blk = [], 0
    for pk in packs:
        js, blk = pk.jobs(blk)
        jobs += js
    arr = (_lib.PackJob * len(jobs))(*jobs)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(packs[0].planes.t.device)
    _pack_table = (raw, len(jobs), blk, packs)


def _refresh_all():
    tab, n, blocks, packs = _pack_table
    _lib.check(_lib.load().get_pack_planes_multi(tab.data_ptr(), n, blocks, _stream()), "get_pack_planes_multi")
    for pk in packs:
        pk.state = (pk.versions(), _pack_epoch)


def refresh_packs() -> bool:
    """Bring every pack up to date NOW on the current stream (one launch). True when the cache is populated and current
    afterwards -- the condition under which independent branches may run on side streams without racing on packs."""
    if not _packs:
        return False
    if any(pk.state is None or pk.state[1] != _pack_epoch for pk in _packs.values()):
        if _pack_table is None:
            if torch.cuda.is_current_stream_capturing():
                return False
            prepare_pack_table()
        _refresh_all()
    return all(pk.state is not None and pk.state[1] == _pack_epoch for pk in _packs.values())


def get_pack(key, build) -> WeightPack:
    """Cached pack for `key`; `build()` -> (rows, cols, blocks, bias_len, bias_blocks) is called on first use only."""
    global _pack_table
    pk = _packs.get(key)
    if pk is None:
        rows, cols, blocks, bias_len, bias_blocks = build()
        dev = blocks[0][0].device
        pk = WeightPack(key, rows, cols, blocks, bias_len, bias_blocks, dev)
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("get_b200: a weight pack was first used inside a stream capture (run one eager step first)")
        js, blk = pk.jobs(0)
        _run_jobs(js, blk)
        pk.state = (pk.versions(), _pack_epoch)
        _packs[key] = pk
        # Packs and job tables are never freed: captured CUDA graphs hold their device addresses (a graph captured before
        # this pack existed keeps replaying with the table it was captured with). The number of distinct packs is bounded by
        # (#weights x #tile configurations).
        if _pack_table is not None:
            _retired_tables.append(_pack_table)
        _pack_table = None
        return pk
    if pk.state == (pk.versions(), _pack_epoch):
        return pk
    if pk.state[1] != _pack_epoch:
        capturing = torch.cuda.is_current_stream_capturing()
        if _pack_table is None and not capturing:
            prepare_pack_table()
        if _pack_table is not None:
            _refresh_all()
            if pk.state == (pk.versions(), _pack_epoch):
                return pk
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("get_b200: stale weight pack inside a stream capture")
    js, blk = pk.jobs(0)            # a single tensor edited in place (version counter moved)
    _run_jobs(js, blk)
    pk.state = (pk.versions(), _pack_epoch)
    return pk


def pack_of(w: torch.Tensor, bias0: Optional[torch.Tensor] = None, bias1: Optional[torch.Tensor] = None) -> WeightPack:
    """Pack of ONE logical (N, K) weight view (any strides: a transposed view packs into its k-contiguous copy)."""
    assert w.dim() == 2 and w.dtype == torch.float32 and w.is_cuda
    key = ("w", _view_key(w), None if bias0 is None else _view_key(bias0), None if bias1 is None else _view_key(bias1))

    def build():
        N, K = w.shape
        bb = [(bias0, bias1, 0)] if bias0 is not None else []
        return N, K, [(w, 0, 0)], (round_up(N, 8) if bias0 is not None else 0), bb
    return get_pack(key, build)


# =====================================================================================================================
# the contraction
# =====================================================================================================================
def tile_n(M: int, N: int, mode: int) -> int:
    return int(_lib.load().get_gemm_bp_tile_n(int(M), int(N), int(mode)))


def gemm_bp(segments: Sequence[Tuple[Planes, Planes, int]], M: int, N: int, *, mode: int, epilogue: int = BPE_STORE,
            C: Optional[torch.Tensor] = None, out1: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
            aux0: Optional[torch.Tensor] = None, aux1: Optional[torch.Tensor] = None, planes_out: Optional[Planes] = None,
            planes_out_n: int = 0, pad_one: bool = False, accumulate: bool = False, group_rows: int = 0,
            zr: Optional[Tuple[int, int]] = None, drop_out: Optional[Tuple[float, int]] = None, tn: int = 0,
            split_k: int = 1, workspace: Optional[torch.Tensor] = None, kblock: int = 0, rowdot=None):
    """acc[m,n] = sum_s A_s(m,:) . B_s(n,:) on the tensor cores; see include/get_b200.h (get_gemm_bp) for the epilogues."""
    lib = _lib.load()
    d = _lib.GemmBpDesc()
    d.nseg = len(segments)
    for s, (a, b, k) in enumerate(segments):
        for dst, t in ((d.A[s], a), (d.B[s], b)):
            dst.ptr, dst.ld, dst.plane_stride, dst.planes, dst.trans = t.ptr, t.ld, t.plane_stride, t.nplanes, t.trans
        d.K[s] = int(k)
    d.M, d.N, d.mode, d.tile_n, d.epilogue, d.accumulate = int(M), int(N), int(mode), int(tn), int(epilogue), int(accumulate)

    def f32(t, name, ld_field=None):
        if t is None:
            return
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("get_b200.gemm_bp: %s must be a float32 CUDA tensor" % name)
        setattr(d, name, t.data_ptr())
        if ld_field:
            assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), name
            setattr(d, ld_field, t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))
    f32(C, "C", "ldc"); f32(out1, "out1", "ld_out1"); f32(bias, "bias")
    f32(aux0, "aux0", "ld_aux0"); f32(aux1, "aux1", "ld_aux1")
    if planes_out is not None:
        d.planes_out, d.ld_planes_out, d.planes_out_stride = planes_out.ptr, planes_out.ld, planes_out.plane_stride
        d.planes_out_n = int(planes_out_n or planes_out.nplanes)
        d.planes_out_pad_one = int(pad_one)
    d.group_rows = int(group_rows)
    if zr is not None:
        d.zr_group_stride, d.zr_cols = int(zr[0]), int(zr[1])
    if drop_out is not None and drop_out[0] > 0:
        d.drop_out_p, d.drop_out_seed = float(drop_out[0]), int(drop_out[1]) & 0xFFFFFFFF
    d.split_k, d.kblock = int(split_k), int(kblock)
    if rowdot is not None:           # (w (N,), out (parts, M), p, seed): TANH_BLEND row-dot by-product
        rw, ro, rp, rs = rowdot
        d.rowdot_w, d.rowdot_out, d.rowdot_p, d.rowdot_seed = rw.data_ptr(), ro.data_ptr(), float(rp), int(rs) & 0xFFFFFFFF
        d.rowdot_out = None
        parts = int(lib.get_gemm_bp_rowdot_parts(ct.byref(d)))
        assert parts == ro.shape[0] and ro.shape[1] == M and ro.is_contiguous() and rw.is_contiguous(), (parts, tuple(ro.shape))
        d.rowdot_out = ro.data_ptr()
    if workspace is not None:
        d.workspace, d.workspace_floats = workspace.data_ptr(), workspace.numel()
    _lib.check(lib.get_gemm_bp(ct.byref(d), _stream()), "get_gemm_bp")
    return d


def _wgrad_tile_n(M: int, N: int, kblocks: int) -> int:
    """N tile of a weight-gradient contraction: few large output tiles, the long contraction split over the SMs. Cost model:
    rounds x k blocks per item x bytes per k block (the MN-major B operand is loaded in 64-column boxes)."""
    best, best_cost = 16, None
    ntm = (M + 127) // 128
    for bn in range(16, 257, 16):
        nt = (N + bn - 1) // bn
        if (nt - 1) * bn >= N:
            continue
        tiles = ntm * nt
        splits = max(1, min(kblocks, _SM_COUNT // tiles if tiles <= _SM_COUNT else 1))
        rounds = (tiles * splits + _SM_COUNT - 1) // _SM_COUNT
        cost = rounds * ((kblocks + splits - 1) // splits) * (128 + 64 * ((bn + 63) // 64) + 24)
        if best_cost is None or cost < best_cost:
            best, best_cost = bn, cost
    return best


def wgrad_bp(a: Planes, b: Planes, M: int, N: int, K: int, mode: int, dsts, accumulate: bool, tn: int = 0):
    """Weight gradient  W[m,n] = sum_k a(m,k) b(n,k)  with both operands MN-major (stored (K, .) row-major), split over k
    across the SMs; `dsts` = [(dst fp32 2-D view or 1-D vector, row0, nrows, col0, ncols)] receive blocks of the result
    (weight gradients and, through a ones column in b, bias gradients) straight from the fixed-order reduction."""
    lib = _lib.load()
    assert a.trans == 1 and b.trans == 1
    d = _lib.GemmBpDesc()
    d.nseg = 1
    for dst, t in ((d.A[0], a), (d.B[0], b)):
        dst.ptr, dst.ld, dst.plane_stride, dst.planes, dst.trans = t.ptr, t.ld, t.plane_stride, t.nplanes, t.trans
    d.K[0], d.M, d.N, d.mode, d.tile_n, d.epilogue = int(K), int(M), int(N), int(mode), int(tn), BPE_STORE
    kblocks = (K + 31) // 32
    bn = tn or _wgrad_tile_n(M, N, kblocks)
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    want = max(1, _SM_COUNT // tiles)
    d.split_k = min(kblocks, max(2, min(want, max(1, kblocks // 4))))      # >= 2 splits: one code path (partials + reduce)
    d.tile_n = bn
    d.workspace, d.workspace_floats = 16, 1 << 60
    splits = int(lib.get_gemm_bp_splits(ct.byref(d)))
    ws_ld = int(lib.get_gemm_bp_ws_ld(ct.byref(d)))
    if splits < 1 or ws_ld < 1:
        _lib.check(-1, "get_gemm_bp (planning a weight gradient)")
    ws = torch.empty((splits * M * ws_ld,), dtype=torch.float32, device=a.t.device)
    d.workspace, d.workspace_floats = ws.data_ptr(), ws.numel()
    if splits == 1:       # a contraction of a single k block: plain store in the workspace layout
        d.C, d.ldc = ws.data_ptr(), ws_ld
    _lib.check(lib.get_gemm_bp(ct.byref(d), _stream()), "get_gemm_bp (weight gradient)")
    arr = (_lib.BpDst * len(dsts))()
    for i, (t, r0, nr, c0, nc) in enumerate(dsts):
        assert t.dtype == torch.float32 and t.is_cuda
        arr[i].dst = t.data_ptr()
        arr[i].ld = (t.stride(0) if t.dim() == 2 else 1)
        arr[i].row0, arr[i].nrows, arr[i].col0, arr[i].ncols = int(r0), int(nr), int(c0), int(nc)
    _lib.check(lib.get_bp_splitk_reduce(ws.data_ptr(), splits, M, ws_ld, arr, len(dsts), int(accumulate), _stream()),
               "get_bp_splitk_reduce")
