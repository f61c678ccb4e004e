"""ctypes binding of ``libsnuffy_b200.so`` — the C-ABI boundary of the hot path.

Every entry point is declared in ``include/snuffy_b200.h``; this module mirrors that header one to one.
There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNUFFY_B200_LIB", os.path.join(_HERE, "libsnuffy_b200.so"))

# activation ids shared with csrc/common.cuh (snuffy.py:216-221 names)
ACT_IDS = {"none": 0, "relu": 1, "gelu": 2, "leakyrelu": 3, "selu": 4, "tanh": 5}

P = c_void_p
I = c_int64


class PlaneJob(ctypes.Structure):
    """snuffy_plane_job_t (include/snuffy_b200.h)."""
    _fields_ = [("src", c_void_p), ("ld", c_int64), ("rows", c_int64), ("cols", c_int64), ("kind", ctypes.c_int32),
                ("plane_rc", ctypes.c_int32), ("dst_row0", c_int64), ("dst_k0", c_int64), ("k_total", c_int64), ("dst", c_void_p), ("plane_stride", c_int64)]


MAX_PLANE_JOBS = 16
# name -> (restype, argtypes).  Order follows include/snuffy_b200.h.
_SIGNATURES = {
    "snuffy_version": (c_int, []),
    "snuffy_last_error": (c_char_p, []),
    "snuffy_sm_count": (c_int, []),
    "snuffy_launch_count": (ctypes.c_longlong, []),
    "snuffy_scores_fwd": (c_int, [P, P, P, P, I, I, I, P]),
    "snuffy_select_topk": (c_int, [P, I, I, I, I, P, P, P]),
    "snuffy_select_random": (c_int, [P, I, I, I, c_uint64, c_uint64, P, P]),
    "snuffy_select_gather": (c_int, [P, P, I, I, I, I, I, c_uint64, c_uint64, P, P, P, P, P]),
    "snuffy_compact_flags": (c_int, [P, I, I, I, P, P, P]),
    "snuffy_gather_rows": (c_int, [P, P, I, I, I, I, P, P]),
    "snuffy_build_row_map": (c_int, [P, I, I, I, P, P]),
    "snuffy_ln_rows_fwd": (c_int, [P, P, P, P, P, I, I, c_int, P, P, I, c_int, P, P]),
    "snuffy_scores_ln_planes_fwd": (c_int, [P, P, P, I, I, I, P, P, I, P, P]),
    "snuffy_ln_rows_scatter_planes": (c_int, [P, P, I, I, I, I, P, P, c_int, P, I, P]),
    "snuffy_ln_mean_head_chunks": (c_int64, [I, I]),
    "snuffy_ln_mean_head_fwd": (c_int, [P, P, P, P, P, I, I, I, I, P, P, P, P, P, P]),
    "snuffy_gemm_f32": (c_int, [P, I, c_int, P, I, c_int, P, I, I, I, I, c_float, P, c_int, P, I, P, P, P, c_float, c_uint64,
                               c_uint64, P]),
    "snuffy_gemm_tc_block_n": (c_int, [I]),
    "snuffy_plane_elems": (c_int64, [I, I, c_int]),
    "snuffy_gemm_tc": (c_int, [P, I, P, I, I, I, I, c_int, P, c_int, P, I, P, P, P, I, P, P, I, c_float, c_uint64, c_uint64,
                              P]),
    "snuffy_gemm_tc_auto_ksplit": (c_int64, [I, I, I]),
    "snuffy_gemm_tc_splitk_workspace": (c_int64, [I, I, I]),
    "snuffy_gemm_tc_splitk": (c_int, [P, I, P, I, I, I, I, c_int, I, P, P, I, P]),
    "snuffy_gemm_tc_awindow": (c_int, [P, I, I, I, P, I, I, I, I, c_int, P, I, P]),
    "snuffy_gemm_tc_splitk_rows": (c_int, [P, I, P, I, I, I, I, c_int, I, P, P, I, P]),
    "snuffy_planes_zero_rows": (c_int, [P, I, I, c_int, I, I, P]),
    "snuffy_plane_job_bytes": (c_int, []),
    "snuffy_weight_planes_batch": (c_int, [ctypes.POINTER(PlaneJob), I, P]),
    "snuffy_planes_t_fwd": (c_int, [P, I, I, I, c_int, c_int, P, P, P, P, P, c_int, c_float, c_uint64, c_uint64, P, I, P]),
    "snuffy_sparse_attn_workspace": (c_int64, [I, I, I, I, I]),
    "snuffy_sparse_attn_fwd": (c_int, [P, I, P, I, P, I, I, I, I, I, c_float, c_uint64, c_uint64, P, P, P, P, I, P]),
    "snuffy_sparse_attn_tc_workspace": (c_int64, [I, I, I, I, I]),
    "snuffy_sparse_attn_tc_fwd": (c_int, [P, I, I, I, I, P, I, I, I, I, I, c_float, c_uint64, c_uint64, P, P, P, P, P, I, P]),
    "snuffy_dsmil_workspace": (c_int64, [I, I, I]),
    "snuffy_dsmil_pool_fwd": (c_int, [P, P, P, P, P, I, I, I, I, P, P, P, P, P, I, P]),
    "snuffy_select_topk_varlen": (c_int, [P, P, I, I, I, I, P, P, P]),
    "snuffy_select_random_varlen": (c_int, [P, P, I, I, I, c_uint64, c_uint64, P, P]),
    "snuffy_sparse_attn_tc_varlen_fwd": (c_int, [P, I, I, I, I, P, P, I, I, I, I, I, P, P, I, P]),
    "snuffy_ln_mean_head_varlen_fwd": (c_int, [P, P, P, P, P, P, I, I, I, I, P, P, P, P, P]),
    "snuffy_gemm_f32_batched_workspace": (c_int64, [I, I, I, I]),
    "snuffy_gemm_f32_auto_ksplit": (c_int64, [I, I, I, I]),
    "snuffy_gemm_f32_batched": (c_int, [P, I, c_int, P, I, c_int, P, I, I, I, I, c_float, P, I, I, I, I, I, I, I, I, I, P, I, P]),
    "snuffy_ln_rows_bwd_blocks": (c_int64, [I]),
    "snuffy_ln_rows_bwd": (c_int, [P, P, I, c_float, P, P, P, P, P, P, I, I, P, P, P, P]),
    "snuffy_fold_partials": (c_int, [P, I, I, P, P]),
    "snuffy_act_bwd": (c_int, [P, P, c_int, c_float, c_uint64, c_uint64, I, P, P, P]),
    "snuffy_residual_dropout": (c_int, [P, P, c_float, c_uint64, c_uint64, I, P, P]),
    "snuffy_colsum_chunks": (c_int64, [I]),
    "snuffy_colsum": (c_int, [P, I, P, I, I, I, P, P, P]),
    "snuffy_attn_rows_bwd": (c_int, [P, P, I, I, I, c_int, c_float, c_float, c_uint64, c_uint64, P, P, P]),
    "snuffy_scatter_add_rows": (c_int, [P, P, P, I, I, I, I, P]),
    "snuffy_block_diag_rows": (c_int, [P, I, I, I, P, P]),
    "snuffy_block_diag_extract": (c_int, [P, I, I, I, P, P]),
    "snuffy_gemm_tc_diag_ksplit": (c_int64, [I, I, I, c_int, I, I]),
    "snuffy_gemm_tc_splitk_blockdiag": (c_int, [P, I, P, I, c_int, I, I, I, c_int, I, I, I, P, P, I, P]),
    "snuffy_gemm_tc_blockdiag": (c_int, [P, I, I, I, P, I, c_int, I, I, I, c_int, I, I, P, I, P]),
    "snuffy_gemm_tc_relugrad": (c_int, [P, I, P, I, I, I, I, c_int, P, c_float, P, I, P, I, P, P, P]),
    "snuffy_gemm_tc_actgrad": (c_int, [P, I, P, I, I, I, I, c_int, P, I, c_int, c_float, c_uint64, c_uint64, P, I, P, I, P, P,
                                       P]),
    "snuffy_attn_seg_bwd": (c_int, [P, P, I, I, I, I, c_int, c_float, c_float, c_uint64, c_uint64, P, P, P, I, P]),
    "snuffy_sparse_attn_bwd_tc_workspace": (c_int64, [I, I, I, I, I]),
    "snuffy_sparse_attn_bwd_tc": (c_int, [P, I, I, I, I, P, P, P, I, I, I, I, I, c_float, P, P, P, P, I, P]),
    "snuffy_softmax_cols_bwd": (c_int, [P, P, I, I, c_float, P, P]),
    "snuffy_mil_loss": (c_int, [P, P, P, P, I, I, I, c_float, P, c_float, P, P, P, P, P, P, P, P]),
    "snuffy_rng_advance": (c_int, [P, c_uint64, P]),
    "snuffy_sumsq_blocks": (c_int64, [I]),
    "snuffy_sumsq": (c_int, [P, I, P, P, P]),
    "snuffy_pack_f32": (c_int, [P, P, P, I, P, P]),
    "snuffy_comm_alloc": (c_int, [I, ctypes.POINTER(c_void_p)]),
    "snuffy_comm_free": (c_int, [P]),
    "snuffy_comm_handle_bytes": (c_int, []),
    "snuffy_comm_export": (c_int, [P, P]),
    "snuffy_comm_import": (c_int, [P, ctypes.POINTER(c_void_p)]),
    "snuffy_comm_close": (c_int, [P]),
    "snuffy_comm_counter_bytes": (c_int, []),
    "snuffy_peer_allreduce": (c_int, [P, P, P, c_int, c_int, I, P]),
    "snuffy_adamw_flat": (c_int, [P, P, P, P, I, c_float, c_float, c_float, c_float, c_float, I, c_float, P, c_float, P]),
    "snuffy_adamw_flat_dev": (c_int, [P, P, P, P, I, c_float, c_float, c_float, c_float, c_float, P, P, P, c_float, P, c_float,
                                     c_float, c_float, P]),
    "snuffy_patch_probs": (c_int, [P, I, P, P]),
    "snuffy_froc_detections": (c_int, [P, I, P, P, I, I, c_float, c_int, c_int, P, P, P, P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class SnuffyLibraryError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"snuffy_b200: native library not found at {LIB_PATH}. Build it with ./build.sh "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
if lib.snuffy_plane_job_bytes() != ctypes.sizeof(PlaneJob):
    raise ImportError("snuffy_b200: PlaneJob does not match snuffy_plane_job_t of the loaded library (rebuild with ./build.sh)")


def last_error() -> str:
    msg = lib.snuffy_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise SnuffyLibraryError(f"{what or 'libsnuffy_b200'} failed (status {status}): {last_error()}")
