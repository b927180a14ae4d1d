"""ctypes binding of libget_b200.so (the C-ABI declared in include/get_b200.h).

There is no fallback: if the shared library is missing and cannot be built, importing raises.
Every wrapper raises RuntimeError when the native call returns a non-zero status.
"""
import ctypes as C
import os

from . import build as _build

GEMM_MAX_SEG = 3

EPI_STORE, EPI_SIGMOID, EPI_TANH_BLEND, EPI_TANH_ROWGROUP, EPI_DGATE_R, EPI_DROPOUT_OUT, EPI_TANH = range(7)


class GemmOperand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("trans", C.c_int32), ("_pad", C.c_int32),
                ("rowidx", C.c_void_p)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", GemmOperand * GEMM_MAX_SEG), ("B", GemmOperand * GEMM_MAX_SEG), ("K", C.c_int32 * GEMM_MAX_SEG),
        ("nseg", C.c_int32), ("M", C.c_int32), ("N", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("alpha", C.c_float), ("accumulate", C.c_int32),
        ("epilogue", C.c_int32),
        ("bias0", C.c_void_p), ("bias1", C.c_void_p),
        ("aux0", C.c_void_p), ("ld_aux0", C.c_int64), ("aux1", C.c_void_p), ("ld_aux1", C.c_int64),
        ("out1", C.c_void_p), ("ld_out1", C.c_int64),
        ("group_rows", C.c_int32),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("drop_cols", C.c_int32), ("_pad0", C.c_int32),
        ("drop_out_p", C.c_float), ("drop_out_seed", C.c_uint32),
        ("split_k", C.c_int32), ("_pad1", C.c_int32), ("workspace", C.c_void_p),
    ]


class BpTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("plane_stride", C.c_int64), ("planes", C.c_int32),
                ("trans", C.c_int32)]


BPE_STORE, BPE_ZR, BPE_TANH_BLEND, BPE_TANH_ROWGROUP, BPE_DGATE_R, BPE_TANH = range(6)
BP_MAX_DST = 12


class GemmBpDesc(C.Structure):
    _fields_ = [
        ("A", BpTensor * GEMM_MAX_SEG), ("B", BpTensor * GEMM_MAX_SEG), ("K", C.c_int32 * GEMM_MAX_SEG),
        ("nseg", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("mode", C.c_int32), ("tile_n", C.c_int32),
        ("epilogue", C.c_int32), ("accumulate", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("out1", C.c_void_p), ("ld_out1", C.c_int64),
        ("bias", C.c_void_p), ("aux0", C.c_void_p), ("ld_aux0", C.c_int64), ("aux1", C.c_void_p), ("ld_aux1", C.c_int64),
        ("planes_out", C.c_void_p), ("ld_planes_out", C.c_int64), ("planes_out_stride", C.c_int64),
        ("planes_out_n", C.c_int32), ("planes_out_pad_one", C.c_int32),
        ("group_rows", C.c_int32), ("zr_group_stride", C.c_int32), ("zr_cols", C.c_int32),
        ("drop_out_p", C.c_float), ("drop_out_seed", C.c_uint32),
        ("split_k", C.c_int32), ("kblock", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_floats", C.c_int64),
        ("rowdot_w", C.c_void_p), ("rowdot_out", C.c_void_p), ("rowdot_p", C.c_float), ("rowdot_seed", C.c_uint32),
    ]


class BpDst(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("ld", C.c_int64), ("row0", C.c_int32), ("nrows", C.c_int32), ("col0", C.c_int32),
                ("ncols", C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src2", C.c_void_p), ("ld_r", C.c_int64), ("ld_c", C.c_int64), ("rows", C.c_int32),
                ("cols", C.c_int32), ("dst", C.c_void_p), ("ld_out", C.c_int64), ("plane_stride", C.c_int64),
                ("first_block", C.c_int64), ("kind", C.c_int32), ("_pad", C.c_int32)]


# name -> (restype, argtypes); must list every symbol of include/get_b200.h (tests/test_abi.py checks it)
_P, _I, _L, _F, _U = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32
SIGNATURES = {
    "get_gemm_f32": (_I, [C.POINTER(GemmDesc), _P]),
    "get_gemm_f32_launches": (_I, [C.POINTER(GemmDesc)]),
    "get_gemm_bp": (_I, [C.POINTER(GemmBpDesc), _P]),
    "get_gemm_bp_tile_n": (_I, [_I, _I, _I]),
    "get_gemm_bp_ws_ld": (_L, [C.POINTER(GemmBpDesc)]),
    "get_gemm_bp_splits": (_I, [C.POINTER(GemmBpDesc)]),
    "get_gemm_bp_rowdot_parts": (_I, [C.POINTER(GemmBpDesc)]),
    "get_bp_splitk_reduce": (_I, [_P, _I, _I, _L, C.POINTER(BpDst), _I, _I, _P]),
    "get_to_planes_bf16": (_I, [_P, _L, _I, _I, _P, _L, _L, _I, _I, _P]),
    "get_pack_planes_multi": (_I, [_P, _I, _L, _P]),
    "get_graph_aggregate_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "get_gsl_fused_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _U, _U, _P, _P, _P, _P]),
    "get_graph_aggregate_bp": (_I, [_P, _P, _P, _P, _P, _L, _L, _I, _I, _I, _I, _I, _I, _I, _P]),
    "get_gsl_fused_bp": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _U, _U, _P, _P, _P, _P, _L, _L, _I, _P]),
    "get_neighbor_lists_rowptr_pitch": (_I, [_I]),
    "get_neighbor_lists_entry_capacity": (_I, [_I]),
    "get_build_neighbor_lists": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "get_graph_gather": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _L, _I, _I, _I, _I, _I, _I, _P]),
    "get_gsl_gather": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _F, _U, _P, _P, _P, _P, _L, _L, _I, _P]),
    "get_rowdot_f32": (_I, [_P, _P, _L, _I, _F, _U, _P, _P]),
    "get_gsl_mask_adj_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "get_att_pool_fwd_f32": (_I, [_P, _P, _L, _P, _P, _I, _I, _I, _I, _I, _P, _P, _L, _P]),
    "get_att_pool_bwd_f32": (_I, [_P, _P, _L, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _L, _I, _P]),
    "get_att_pool_bwd_bp": (_I, [_P, _P, _L, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _P, _P, _L, _L, _I, _P, _P, _L, _I, _P]),
    "get_ggnn_gate_bwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _L, _P, _P, _P, _P]),
    "get_colsum_f32": (_I, [_P, _L, _I, _I, _P, _P, _P]),
    "get_colsum_workspace_floats": (_L, [_I, _I]),
    "get_rows_gather_f32": (_I, [_P, _L, _P, _I, _I, _P, _L, _P]),
    "get_rows_scatter_f32": (_I, [_P, _L, _P, _I, _I, _P, _L, _P]),
    "get_segment_sum_f32": (_I, [_P, _L, _P, _I, _I, _P, _L, _P]),
    "get_segments_i32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "get_ids_mask_u8": (_I, [_P, _I, _L, _I, _P, _P]),
    "get_embedding_rows_fwd_f32": (_I, [_P, _I, _I, _P, _I, _P, _L, _P]),
    "get_embedding_rows_bwd_f32": (_I, [_P, _L, _P, _I, _I, _I, _P, _I, _P]),
    "get_masked_mean_fwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "get_masked_mean_bwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "get_dropout_mask_f32": (_I, [_P, _L, _F, _U, _P]),
    "get_rows_gather_dropout_f32": (_I, [_P, _L, _P, _I, _I, _F, _U, _P, _L, _P]),
    "get_ggnn_gate_bwd_bp": (_I, [_P, _P, _P, _P, _I, _I, _P, _L, _L, _I, _I, _I, _P, _P]),
    "get_rows_gather_dropout_bp": (_I, [_P, _L, _I, _P, _I, _I, _F, _U, _P, _L, _L, _I, _P]),
    "get_dropout_salt_set": (_I, [_U, _P]),
    "get_dropout_salt_advance": (_I, [_P]),
    "get_dropout_salt_get": (_I, [C.POINTER(C.c_uint32)]),
    "get_adam_flat_f32": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _P, _P]),
    "get_build_word_graphs": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "get_cross_entropy_f32": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "get_b200_abi_version": (_I, []),
    "get_b200_last_error": (C.c_char_p, []),
    "get_b200_launch_count": (_L, []),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if necessary) the native library. Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not _build.is_current():
        # missing, or built from other sources than the ones next to it: rebuild (raises RuntimeError when nvcc is
        # unavailable and there is no library at all; a stale library without nvcc is refused rather than loaded)
        try:
            path = _build.build()
        except RuntimeError:
            if os.path.exists(path):
                raise RuntimeError("libget_b200.so is stale (sources changed since it was built) and nvcc is not available "
                                   "to rebuild it: run `python -m get_b200.build` where nvcc exists")
            raise
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.get_b200_abi_version() != 2:
        raise RuntimeError("libget_b200.so ABI version mismatch")
    # allocate the dropout salt word now (never inside a stream capture); fails harmlessly on a box without a GPU
    lib.get_dropout_salt_get(C.byref(C.c_uint32(0)))
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().get_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (status %d): %s" % (what, status, msg))


def launch_count() -> int:
    return int(load().get_b200_launch_count())
