"""Tensor-level wrappers over the C ABI (one Python function per ``extern "C"`` entry point).

PyTorch is used for device memory and the current CUDA stream only; every computation happens inside
``libsnuffy_b200.so``.  All wrappers are asynchronous on ``torch.cuda.current_stream()`` and never
synchronise, so a whole forward can be captured into a CUDA graph.
"""
from __future__ import annotations

import functools
import math
from typing import Optional, Tuple

import torch

from ._lib import ACT_IDS, MAX_PLANE_JOBS, PlaneJob, check, lib


_U64 = 2**64 - 1


# ------------------------------------------------------------------ helpers
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    # current CUDA stream of the current device as a raw cudaStream_t (the fast path skips the Stream object)
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


# size queries into the library are pure functions of their integer arguments: cache them (the training step makes ~140
# launches per bag and is otherwise bound by host-side call overhead)
_plane_elems = functools.lru_cache(maxsize=None)(lambda rows, K, rc: lib.snuffy_plane_elems(rows, K, rc))
_block_n = functools.lru_cache(maxsize=None)(lambda n: lib.snuffy_gemm_tc_block_n(n))
_colsum_chunks = functools.lru_cache(maxsize=None)(lambda rows: lib.snuffy_colsum_chunks(rows))
_ln_bwd_blocks = functools.lru_cache(maxsize=None)(lambda rows: lib.snuffy_ln_rows_bwd_blocks(rows))
_tc_auto_ksplit = functools.lru_cache(maxsize=None)(lambda M, N, K: lib.snuffy_gemm_tc_auto_ksplit(M, N, K))
_f32_auto_ksplit = functools.lru_cache(maxsize=None)(lambda nb, M, N, K: lib.snuffy_gemm_f32_auto_ksplit(nb, M, N, K))


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"snuffy_b200: `{name}` must live on a CUDA device (there is no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"snuffy_b200: `{name}` must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Planes:
    """Split-bf16 operand planes of an fp32 [rows, K] matrix (layout: csrc/common.cuh)."""

    __slots__ = ("buf", "rows", "K", "rc", "stride")

    def __init__(self, rows: int, K: int, rc: int, device, zero: bool = False):
        self.rows, self.K, self.rc = rows, K, rc
        self.stride = _plane_elems(rows, K, rc)
        alloc = torch.zeros if zero else torch.empty
        self.buf = alloc(2 * self.stride, dtype=torch.bfloat16, device=device)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    def row_window(self, row0: int, rows: int) -> "Planes":
        """View of `rows` rows starting at row0 (a multiple of rc = whole row tiles) sharing the storage: same K, same
        distance between the hi and lo plane."""
        if row0 % self.rc:
            raise ValueError(f"Planes.row_window: row0={row0} must be a multiple of {self.rc}")
        w = Planes.__new__(Planes)
        tile_elems = self.stride // ((self.rows + self.rc - 1) // self.rc)        # elements of one row tile in one plane
        w.buf, w.rows, w.K, w.rc, w.stride = self.buf[(row0 // self.rc) * tile_elems:], rows, self.K, self.rc, self.stride
        return w


# ------------------------------------------------------------------ a1: instance scores
def scores(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """c = x W^T + b over the last dim (snuffy.py:39-41).  x [..., d] -> c [..., C]."""
    x = _f32(x, "x")
    weight = _f32(weight, "weight")
    d = x.shape[-1]
    C = weight.shape[0]
    rows = x.numel() // d if d else 0
    c = torch.empty(*x.shape[:-1], C, dtype=torch.float32, device=x.device)
    check(lib.snuffy_scores_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), c.data_ptr(), rows, d, C, _stream()),
          "snuffy_scores_fwd")
    return c


# ------------------------------------------------------------------ a6 / a12: selection
def select_topk(c: torch.Tensor, k: int, flags: Optional[torch.Tensor] = None) -> torch.Tensor:
    """c [B, N, C] -> idx [B, C, k] int64: per (bag, class) the k best rows in descending score order
    (ties: lower index first).  `flags` [B, N] uint8 (zeroed by the caller) receives 1 at every winner."""
    c = _f32(c, "c")
    B, N, C = c.shape
    idx = torch.empty(B, C, k, dtype=torch.int64, device=c.device)
    check(lib.snuffy_select_topk(c.data_ptr(), B, N, C, k, idx.data_ptr(), _ptr(flags), _stream()), "snuffy_select_topk")
    return idx


def select_random(flags: torch.Tensor, k: int, seed: int, offset: int) -> torch.Tensor:
    """flags [B, N] uint8 -> idx [B, k] int64, k distinct un-flagged rows, uniform without replacement."""
    B, N = flags.shape
    idx = torch.empty(B, k, dtype=torch.int64, device=flags.device)
    check(lib.snuffy_select_random(flags.data_ptr(), B, N, k, seed & _U64, offset & _U64, idx.data_ptr(),
                                   _stream()), "snuffy_select_random")
    return idx


def compact_flags(flags: torch.Tensor, cap: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ascending indices of flagged rows per bag (torch.unique order): (idx [B, cap], counts [B] int32)."""
    B, N = flags.shape
    out = torch.zeros(B, cap, dtype=torch.int64, device=flags.device)
    counts = torch.empty(B, dtype=torch.int32, device=flags.device)
    check(lib.snuffy_compact_flags(flags.data_ptr(), B, N, cap, out.data_ptr(), counts.data_ptr(), _stream()),
          "snuffy_compact_flags")
    return out, counts


def gather_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x [B, N, d], idx [B, K] -> [B, K, d] (raw rows, snuffy.py:131,145-147)."""
    x = _f32(x, "x")
    B, N, d = x.shape
    K = idx.shape[1]
    out = torch.empty(B, K, d, dtype=torch.float32, device=x.device)
    check(lib.snuffy_gather_rows(x.data_ptr(), idx.data_ptr(), B, N, K, d, out.data_ptr(), _stream()), "snuffy_gather_rows")
    return out


def build_row_map(idx: torch.Tensor, N: int) -> torch.Tensor:
    """idx [B, K] -> row_map [B*N] int32: -1 or the slot b*K+k of a selected row."""
    B, K = idx.shape
    row_map = torch.empty(B * N, dtype=torch.int32, device=idx.device)
    check(lib.snuffy_build_row_map(idx.data_ptr(), B, N, K, row_map.data_ptr(), _stream()), "snuffy_build_row_map")
    return row_map


# ------------------------------------------------------------------ LayerNorm
def ln_rows(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], *,
            row_map: Optional[torch.Tensor] = None, alt: Optional[torch.Tensor] = None, apply_ln: bool = True,
            want_f32: bool = False, want_planes: bool = False, plane_rc: int = 128, want_stats: bool = False,
            zero_planes: bool = False, affine: bool = True):
    """LN over the rows of y = x-with-mapped-rows-replaced.  Returns (fp32 or None, Planes or None, stats or None).
    apply_ln=False: plain convert (a given `gamma` scales the columns); affine=False: normalise only (z)."""
    x = _f32(x, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    out = torch.empty(rows, d, dtype=torch.float32, device=x.device) if want_f32 else None
    planes = Planes(rows, d, plane_rc, x.device, zero=zero_planes) if want_planes else None
    stats = torch.empty(rows, 2, dtype=torch.float32, device=x.device) if want_stats else None
    check(lib.snuffy_ln_rows_fwd(x.data_ptr(), _ptr(row_map), _ptr(alt), _ptr(gamma), _ptr(beta), rows, d,
                                 (1 if affine else 2) if apply_ln else 0, _ptr(out), planes.ptr if planes else None,
                                 planes.stride if planes else 0, plane_rc, _ptr(stats), _stream()), "snuffy_ln_rows_fwd")
    return out, planes, stats


def scores_ln_planes(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]):
    """One pass over x [..., d]: (c [..., C] instance scores, Planes of the normalised rows).  Layer-0 fusion."""
    x = _f32(x, "x")
    weight = _f32(weight, "weight")
    d = x.shape[-1]
    rows = x.numel() // d
    C = weight.shape[0]
    c = torch.empty(*x.shape[:-1], C, dtype=torch.float32, device=x.device)
    planes = Planes(rows, d, 128, x.device)
    check(lib.snuffy_scores_ln_planes_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), rows, d, C, c.data_ptr(), planes.ptr,
                                          planes.stride, None, _stream()), "snuffy_scores_ln_planes_fwd")
    return c, planes


def ln_rows_scatter_planes(src: torch.Tensor, idx: torch.Tensor, N: int, planes: Planes, gamma=None, beta=None,
                           apply_ln: bool = True, affine: bool = False) -> None:
    """planes rows b*N + idx[b, k] <- split(LN(src[b*K + k]))  (in place; src [B*K, d], idx [B, K])."""
    src = _f32(src, "src")
    B, K = idx.shape
    mode = (1 if affine else 2) if apply_ln else 0
    check(lib.snuffy_ln_rows_scatter_planes(src.data_ptr(), idx.data_ptr(), B, N, K, src.shape[-1], _ptr(gamma), _ptr(beta),
                                            mode, planes.ptr, planes.stride, _stream()), "snuffy_ln_rows_scatter_planes")


_HEAD_WS = {}


def ln_mean_head(x: torch.Tensor, gamma, beta, w_head, b_head, *, want_stats: bool = False, want_pooled: bool = False):
    """bag[B, C] = head(mean_n LN_f(x[b]))  (snuffy.py:86,71).  x [B, N, d]."""
    x = _f32(x, "x")
    B, N, d = x.shape
    C = w_head.shape[0]
    chunks = lib.snuffy_ln_mean_head_chunks(B, N)
    key = (x.device.index, B)
    tickets = _HEAD_WS.get(key)
    if tickets is None:                      # zeroed once; the kernel leaves it zeroed
        tickets = _HEAD_WS[key] = torch.zeros(B, dtype=torch.int32, device=x.device)
    partials = torch.empty(B * chunks * d, dtype=torch.float32, device=x.device)
    stats = torch.empty(B * N, 2, dtype=torch.float32, device=x.device) if want_stats else None
    pooled = torch.empty(B, d, dtype=torch.float32, device=x.device) if want_pooled else None
    bag = torch.empty(B, C, dtype=torch.float32, device=x.device)
    check(lib.snuffy_ln_mean_head_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w_head.data_ptr(), _ptr(b_head),
                                      B, N, d, C, partials.data_ptr(), tickets.data_ptr(), _ptr(stats), _ptr(pooled),
                                      bag.data_ptr(), _stream()), "snuffy_ln_mean_head_fwd")
    return bag, stats, pooled


# ------------------------------------------------------------------ GEMMs
def gemm_f32(a: torch.Tensor, b: torch.Tensor, *, a_kc: bool = True, b_kc: bool = True, M: int, N: int, K: int,
             bias=None, act: str = "none", resid=None, row_map=None, resid_alt=None, alpha: float = 1.0,
             want_preact: bool = False, lda: Optional[int] = None, ldb: Optional[int] = None,
             out: Optional[torch.Tensor] = None, drop: Tuple[float, int, int] = (0.0, 0, 0)):
    """fp32 SIMT GEMM  C[M,N] = act(alpha * A.B^T + bias) + resid  (csrc/gemm_simt.cu)."""
    a = _f32(a, "a")
    b = _f32(b, "b")
    lda = lda if lda is not None else a.shape[-1]
    ldb = ldb if ldb is not None else b.shape[-1]
    c = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=a.device)
    pre = torch.empty(M, N, dtype=torch.float32, device=a.device) if want_preact else None
    ldr = resid.shape[-1] if resid is not None else 0
    check(lib.snuffy_gemm_f32(a.data_ptr(), lda, 1 if a_kc else 0, b.data_ptr(), ldb, 1 if b_kc else 0, c.data_ptr(), N,
                              M, N, K, alpha, _ptr(bias), ACT_IDS[act], _ptr(resid), ldr, _ptr(row_map), _ptr(resid_alt),
                              _ptr(pre), float(drop[0]), drop[1] & _U64, drop[2] & _U64, _stream()), "snuffy_gemm_f32")
    return (c, pre) if want_preact else c


def linear_f32(x: torch.Tensor, weight: torch.Tensor, bias=None, act: str = "none", resid=None,
               drop: Tuple[float, int, int] = (0.0, 0, 0)) -> torch.Tensor:
    """y = dropout(act(x W^T + b)) (+ resid) for x [rows, K], weight [N, K]."""
    rows, K = x.shape
    return gemm_f32(x, weight, M=rows, N=weight.shape[0], K=K, bias=bias, act=act, resid=resid, drop=drop)


def weight_planes(weight: torch.Tensor, col_gain: Optional[torch.Tensor] = None) -> Planes:
    """Split an nn.Linear weight [N, K] into B-operand planes (zero padded to the kernel's BLOCK_N).  `col_gain` [K]
    scales the columns first (W * gamma: a LayerNorm gain folded into the weight that consumes the normalised rows)."""
    weight = _f32(weight, "weight")
    n = weight.shape[0]
    rc = _block_n(n)
    _, planes, _ = ln_rows(weight, col_gain, None, apply_ln=False, want_planes=True, plane_rc=rc, zero_planes=True)
    return planes


def gemm_tc(a: Planes, b: Planes, *, M: int, N: int, K: int, passes: int = 3, bias=None, act: str = "none",
            resid=None, row_map=None, resid_alt=None, want_out: bool = True, want_preact: bool = False,
            want_planes: bool = False, device=None, drop: Tuple[float, int, int] = (0.0, 0, 0)):
    """tcgen05 split-bf16 GEMM (csrc/gemm_tc.cu).  Returns (out fp32 or None, preact or None, Planes or None)."""
    device = device if device is not None else a.buf.device
    out = torch.empty(M, N, dtype=torch.float32, device=device) if want_out else None
    pre = torch.empty(M, N, dtype=torch.float32, device=device) if want_preact else None
    op = Planes(M, N, 128, device) if want_planes else None
    ldr = resid.shape[-1] if resid is not None else 0
    check(lib.snuffy_gemm_tc(a.ptr, a.stride, b.ptr, b.stride, M, N, K, passes, _ptr(bias), ACT_IDS[act], _ptr(resid),
                             ldr, _ptr(row_map), _ptr(resid_alt), _ptr(out), N, _ptr(pre), op.ptr if op else None,
                             op.stride if op else 0, float(drop[0]), drop[1] & _U64, drop[2] & _U64, _stream()),
          "snuffy_gemm_tc")
    return out, pre, op


def gemm_tc_actgrad(a: Planes, b: Planes, gate: torch.Tensor, act: str, *, M: int, N: int, K: int, passes: int = 3,
                    drop: Tuple[float, int, int] = (0.0, 0, 0), want_out: bool = True, want_planes: bool = True,
                    want_colsum: bool = False):
    """(A . B^T) * act'(gate) * dropout mask in the GEMM epilogue -> (fp32 [M, N] or None, A-operand Planes or None[, column
    sums [N] of the result when want_colsum])."""
    gate = _f32(gate, "gate")
    device = a.buf.device
    out = torch.empty(M, N, dtype=torch.float32, device=device) if want_out else None
    op = Planes(M, N, 128, device) if want_planes else None
    cs = torch.empty(N, dtype=torch.float32, device=device) if want_colsum else None
    csp = torch.empty((M + 127) // 128 * 4 * N, dtype=torch.float32, device=device) if want_colsum else None
    check(lib.snuffy_gemm_tc_actgrad(a.ptr, a.stride, b.ptr, b.stride, M, N, K, passes, gate.data_ptr(), gate.shape[-1],
                                     ACT_IDS[act], float(drop[0]), drop[1] & _U64, drop[2] & _U64, _ptr(out), N,
                                     op.ptr if op else None, op.stride if op else 0, _ptr(cs), _ptr(csp), _stream()),
          "snuffy_gemm_tc_actgrad")
    return (out, op, cs) if want_colsum else (out, op)


def gemm_tc_relugrad(a: Planes, b: Planes, act_planes: Planes, dropout_p: float, *, M: int, N: int, K: int, passes: int = 3,
                     want_out: bool = False, want_planes: bool = True, want_colsum: bool = False):
    """(A . B^T) * (act > 0 ? 1 / (1 - p) : 0) with `act_planes` = the forward's dropout(relu(h)) operand planes [M, N]:
    ReLU backward with neither the saved pre-activation nor the dropout draw.  -> (fp32 or None, Planes or None, colsum or None)."""
    if act_planes.rc != 128 or act_planes.K != N or act_planes.rows < M:
        raise ValueError("gemm_tc_relugrad: act_planes are the 128-row planes of an [M, N] activation")
    device = a.buf.device
    out = torch.empty(M, N, dtype=torch.float32, device=device) if want_out else None
    op = Planes(M, N, 128, device) if want_planes else None
    cs = torch.empty(N, dtype=torch.float32, device=device) if want_colsum else None
    csp = torch.empty((M + 127) // 128 * 4 * N, dtype=torch.float32, device=device) if want_colsum else None
    check(lib.snuffy_gemm_tc_relugrad(a.ptr, a.stride, b.ptr, b.stride, M, N, K, passes, act_planes.ptr, float(dropout_p),
                                      _ptr(out), N, op.ptr if op else None, op.stride if op else 0, _ptr(cs), _ptr(csp),
                                      _stream()), "snuffy_gemm_tc_relugrad")
    return out, op, cs


# ------------------------------------------------------------------ a9: sparse attention
def _rows_view(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D fp32 CUDA tensor whose last dim is contiguous (rows may be strided, e.g. a column slice of Q|V)."""
    if not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"snuffy_b200: `{name}` must be a float32 CUDA tensor")
    if t.dim() != 2 or t.stride(1) != 1:
        t = t.reshape(-1, t.shape[-1]).contiguous()
    return t


def sparse_attn(q: torch.Tensor, v: torch.Tensor, kp: torch.Tensor, B: int, N: int, Ksel: int, h: int, *,
                want_probs: bool = True, want_stats: bool = False, dropout_p: float = 0.0, seed: int = 0,
                offset: int = 0):
    """q, v [B*N, d] (row-strided views allowed), kp [B*Ksel, d] -> O [B*Ksel, d], P [B, h, N, Ksel] or None,
    stats [B, h, N, 2] = (row max, 1/row sum) or None.   snuffy.py:160-168."""
    q, v = _rows_view(q, "q"), _rows_view(v, "v")
    kp = _f32(kp, "kp")
    d = kp.shape[-1]
    dev = q.device
    ws_bytes = lib.snuffy_sparse_attn_workspace(B, N, Ksel, h, d)
    if ws_bytes < 0:
        raise ValueError(f"d_model={d} is not divisible by h={h}")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    o = torch.empty(B * Ksel, d, dtype=torch.float32, device=dev)
    probs = torch.empty(B, h, N, Ksel, dtype=torch.float32, device=dev) if want_probs else None
    stats = torch.empty(B, h, N, 2, dtype=torch.float32, device=dev) if want_stats else None
    check(lib.snuffy_sparse_attn_fwd(q.data_ptr(), q.stride(0), v.data_ptr(), v.stride(0), kp.data_ptr(), B, N, Ksel, h,
                                     d, float(dropout_p), seed & _U64, offset & _U64, o.data_ptr(), _ptr(probs),
                                     _ptr(stats), ws.data_ptr(), ws_bytes, _stream()), "snuffy_sparse_attn_fwd")
    return o, probs, stats


def sparse_attn_tc_supported(B: int, N: int, Ksel: int, h: int, d: int) -> bool:
    return d % 32 == 0 and lib.snuffy_sparse_attn_tc_workspace(B, N, Ksel, h, d) >= 0


def sparse_attn_tc(qv_planes: Planes, kp: torch.Tensor, B: int, N: int, Ksel: int, h: int, d: int, *,
                   want_probs: bool = True, want_stats: bool = False, dropout_p: float = 0.0, seed: int = 0,
                   offset: int = 0, want_mask: bool = False):
    """Tensor-core sparse attention on the Q|V planes written by the projection GEMM (planes over [B*N, 2d]).
    want_mask (with dropout): also returns the keep bits of the draw, [B, h, N, ceil(Ksel / 8)] uint8, for the fused backward."""
    kp = _f32(kp, "kp")
    dev = kp.device
    ws_bytes = lib.snuffy_sparse_attn_tc_workspace(B, N, Ksel, h, d)
    if ws_bytes < 0:
        raise ValueError("shape not served by the tensor-core attention kernel")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    o = torch.empty(B * Ksel, d, dtype=torch.float32, device=dev)
    probs = torch.empty(B, h, N, Ksel, dtype=torch.float32, device=dev) if want_probs else None
    stats = torch.empty(B, h, N, 2, dtype=torch.float32, device=dev) if want_stats else None
    mask = torch.empty(B, h, N, (Ksel + 7) // 8, dtype=torch.uint8, device=dev) if want_mask and dropout_p > 0 else None
    check(lib.snuffy_sparse_attn_tc_fwd(qv_planes.ptr, qv_planes.stride, qv_planes.K, 0, d, kp.data_ptr(), B, N, Ksel, h, d,
                                        float(dropout_p), seed & _U64, offset & _U64, o.data_ptr(), _ptr(probs),
                                        _ptr(stats), _ptr(mask), ws.data_ptr(), ws_bytes, _stream()), "snuffy_sparse_attn_tc_fwd")
    if want_mask:
        return o, probs, stats, mask
    return o, probs, stats


# ------------------------------------------------------------------ a15: DSMIL pooling
def dsmil_pool(q: torch.Tensor, q_max: torch.Tensor, v: torch.Tensor, w_fcc: torch.Tensor, b_fcc):
    """A[N,C], Bm[C,d], logits[C] (dsmil.py:83-91)."""
    q, q_max, v, w_fcc = _f32(q, "q"), _f32(q_max, "q_max"), _f32(v, "v"), _f32(w_fcc, "w_fcc")
    N, dq = q.shape
    C = q_max.shape[0]
    d = v.shape[1]
    dev = q.device
    ws_bytes = lib.snuffy_dsmil_workspace(N, d, C)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    A = torch.empty(N, C, dtype=torch.float32, device=dev)
    Bm = torch.empty(C, d, dtype=torch.float32, device=dev)
    logits = torch.empty(C, dtype=torch.float32, device=dev)
    check(lib.snuffy_dsmil_pool_fwd(q.data_ptr(), q_max.data_ptr(), v.data_ptr(), w_fcc.data_ptr(), _ptr(b_fcc), N, d, dq,
                                    C, A.data_ptr(), Bm.data_ptr(), logits.data_ptr(), None, ws.data_ptr(), ws_bytes,
                                    _stream()), "snuffy_dsmil_pool_fwd")
    return A, Bm, logits


# ------------------------------------------------------------------ backward kernels (csrc/backward.cu, gemm_simt.cu)
def gemm_f32_batched(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int,
                     ldc: int, a_kc: bool = True, b_kc: bool = True, alpha: float = 1.0, bias=None, nb_outer: int = 1,
                     nb_inner: int = 1, sa=(0, 0), sb=(0, 0), sc=(0, 0), ksplit: Optional[int] = None) -> torch.Tensor:
    """c_z[M, N] = alpha * a_z . b_z^T (+ bias) for every batch z = (outer, inner); operands are addressed from the
    tensors' data pointers with element strides (outer, inner).  ksplit None = pick one that fills the machine."""
    nbatch = nb_outer * nb_inner
    if ksplit is None:
        ksplit = _f32_auto_ksplit(nbatch, M, N, K)
    ws_bytes = lib.snuffy_gemm_f32_batched_workspace(nbatch, M, N, ksplit)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device) if ws_bytes else None
    check(lib.snuffy_gemm_f32_batched(a.data_ptr(), lda, 1 if a_kc else 0, b.data_ptr(), ldb, 1 if b_kc else 0, c.data_ptr(),
                                      ldc, M, N, K, alpha, _ptr(bias), nb_outer, nb_inner, sa[0], sa[1], sb[0], sb[1], sc[0],
                                      sc[1], ksplit, _ptr(ws), ws_bytes, _stream()), "snuffy_gemm_f32_batched")
    return c


def matmul_nt(a: torch.Tensor, b: torch.Tensor, *, resid=None, alpha: float = 1.0) -> torch.Tensor:
    """a [M, K] . b [N, K]^T -> [M, N]  (dX = dY . W uses matmul_nn)."""
    M, K = a.shape
    return gemm_f32(a, b, M=M, N=b.shape[0], K=K, resid=resid, alpha=alpha)


def matmul_nn(a: torch.Tensor, b: torch.Tensor, *, resid=None) -> torch.Tensor:
    """a [M, K] . b [K, N] -> [M, N]   (dX = dY . W with W = nn.Linear.weight [out, in])."""
    M, K = a.shape
    return gemm_f32(a, b, M=M, N=b.shape[1], K=K, b_kc=False, resid=resid)


def matmul_tn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [R, M]^T . b [R, N] -> [M, N], split over the long R axis   (dW = dY^T X)."""
    a, b = _f32(a, "a"), _f32(b, "b")
    R, M = a.shape
    N = b.shape[1]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    return gemm_f32_batched(a, b, c, M=M, N=N, K=R, lda=M, ldb=N, ldc=N, a_kc=False, b_kc=False)


def ln_rows_bwd(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, *, dy: Optional[torch.Tensor] = None,
                dy_bcast: Optional[torch.Tensor] = None, rows_per_bag: int = 1, bscale: float = 1.0, row_map=None,
                alt=None, add=None, want_dx: bool = True, defer_fold: bool = False):
    """LayerNorm backward.  Returns (dx [rows, d] or None, dgamma [d], dbeta [d]); with defer_fold (dx, fold) where fold()
    -> (dgamma, dbeta) sums the per-CTA partials later, on whatever stream is current then."""
    x = _f32(x, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    dev = x.device
    dx = torch.empty(rows, d, dtype=torch.float32, device=dev) if want_dx else None
    gb = torch.empty(2, d, dtype=torch.float32, device=dev)
    blocks = _ln_bwd_blocks(rows)
    partials = torch.empty(blocks * 2 * d, dtype=torch.float32, device=dev)
    if dy is not None:
        dy = _f32(dy, "dy")
    if dy_bcast is not None:
        dy_bcast = _f32(dy_bcast, "dy_bcast")
    if add is not None:
        add = _f32(add, "add")
    check(lib.snuffy_ln_rows_bwd(_ptr(dy), _ptr(dy_bcast), rows_per_bag, float(bscale), x.data_ptr(), _ptr(row_map), _ptr(alt),
                                 stats.data_ptr(), gamma.data_ptr(), _ptr(add), rows, d, _ptr(dx),
                                 None if defer_fold else gb.data_ptr(), partials.data_ptr(), _stream()), "snuffy_ln_rows_bwd")
    if defer_fold:
        def fold():
            out = torch.empty(2, d, dtype=torch.float32, device=dev)
            check(lib.snuffy_fold_partials(partials.data_ptr(), blocks, 2 * d, out.data_ptr(), _stream()), "snuffy_fold_partials")
            return out[0], out[1]
        fold.partials = partials
        return dx, fold
    return dx, gb[0], gb[1]


def act_bwd(hpre: Optional[torch.Tensor], da: Optional[torch.Tensor], act: str = "none",
            drop: Tuple[float, int, int] = (0.0, 0, 0), want_dh: bool = True, want_a: bool = False):
    """(dh, a): dh = da * mask * act'(hpre); a = act(hpre) * mask.  hpre None = dropout mask only."""
    ref = hpre if hpre is not None else da
    ref = _f32(ref, "hpre/da")
    dh = torch.empty_like(ref) if want_dh else None
    a = torch.empty_like(ref) if want_a else None
    check(lib.snuffy_act_bwd(_ptr(hpre), _ptr(da), ACT_IDS[act], float(drop[0]), drop[1] & _U64, drop[2] & _U64, ref.numel(),
                             _ptr(dh), _ptr(a), _stream()), "snuffy_act_bwd")
    return dh, a


def residual_dropout(x: torch.Tensor, y: torch.Tensor, drop: Tuple[float, int, int] = (0.0, 0, 0)) -> torch.Tensor:
    """x + dropout(y) (snuffy.py:108,110); the mask is the one act_bwd(None, g, drop=drop) applies in the backward."""
    x, y = _f32(x, "x"), _f32(y, "y")
    if x.shape != y.shape:
        raise ValueError(f"residual_dropout: shapes differ ({tuple(x.shape)} vs {tuple(y.shape)})")
    out = torch.empty_like(x)
    check(lib.snuffy_residual_dropout(x.data_ptr(), y.data_ptr(), float(drop[0]), drop[1] & _U64, drop[2] & _U64, x.numel(),
                                      out.data_ptr(), _stream()), "snuffy_residual_dropout")
    return out


def colsum(x: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [rows, d] (row stride allowed), w [rows, C] or None -> [C, d]: out[c] = sum_rows w[row, c] * x[row]."""
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.reshape(-1, x.shape[-1]).contiguous()
    rows, d = x.shape
    C = 1 if w is None else w.shape[-1]
    if w is not None:
        w = _f32(w.reshape(rows, C), "w")
    out = torch.empty(C, d, dtype=torch.float32, device=x.device)
    partials = torch.empty(_colsum_chunks(rows) * C * d, dtype=torch.float32, device=x.device)
    check(lib.snuffy_colsum(x.data_ptr(), x.stride(0), _ptr(w), rows, d, C, out.data_ptr(), partials.data_ptr(), _stream()),
          "snuffy_colsum")
    return out


def attn_rows_bwd(s: torch.Tensor, stats: torch.Tensor, N: int, mode: int, scale: float,
                  drop: Tuple[float, int, int] = (0.0, 0, 0), pd: Optional[torch.Tensor] = None,
                  g: Optional[torch.Tensor] = None) -> None:
    nrows, ksel = s.shape
    check(lib.snuffy_attn_rows_bwd(s.data_ptr(), stats.data_ptr(), nrows, ksel, N, mode, float(scale), float(drop[0]),
                                   drop[1] & _U64, drop[2] & _U64, _ptr(pd), _ptr(g), _stream()), "snuffy_attn_rows_bwd")


def scatter_add_rows(dx: torch.Tensor, idx: torch.Tensor, src: torch.Tensor) -> None:
    """dx [B, N, d] (in place) += src [B*K, d] at rows idx [B, K]."""
    B, N, d = dx.shape
    K = idx.shape[1]
    check(lib.snuffy_scatter_add_rows(dx.data_ptr(), idx.data_ptr(), src.data_ptr(), B, N, K, d, _stream()),
          "snuffy_scatter_add_rows")


def sparse_attn_bwd(q: torch.Tensor, v: torch.Tensor, kp: torch.Tensor, d_o: torch.Tensor, stats: torch.Tensor, B: int,
                    N: int, Ksel: int, h: int, drop: Tuple[float, int, int] = (0.0, 0, 0)):
    """Backward of O_j = softmax_keys(Q_j Kp_j^T / sqrt(dk))^T V_j  (snuffy.py:160-168).

    q, v: [B*N, d] (row-strided views allowed); kp, d_o: [B*Ksel, d]; stats [B, h, N, 2] from the forward.
    Returns (dQ, dV (column halves of dQV), dKp [B*Ksel, d], dQV [B*N, 2d]).  The [B, h, N, Ksel] score tensors are materialised here
    (recompute, never saved by the forward) and every contraction is one head-batched SIMT GEMM launch."""
    q, v = _rows_view(q, "q"), _rows_view(v, "v")
    kp, d_o = _f32(kp, "kp"), _f32(d_o, "d_o")
    d = kp.shape[-1]
    dk = d // h
    dev = q.device
    ldq, ldv = q.stride(0), v.stride(0)
    scale = math.sqrt(dk)
    S = torch.empty(B * h * N, Ksel, dtype=torch.float32, device=dev)
    Pd = torch.empty_like(S)
    G = torch.empty_like(S)
    dqv = torch.empty(B * N, 2 * d, dtype=torch.float32, device=dev)      # dQ | dV side by side: one operand for dWqv / du1
    dq, dv = dqv[:, :d], dqv[:, d:]
    dkp = torch.empty(B * Ksel, d, dtype=torch.float32, device=dev)
    hb = dict(nb_outer=B, nb_inner=h)
    s_str = (h * N * Ksel, N * Ksel)                    # [B, h, N, Ksel] matrices
    # S = Q_j Kp_j^T / sqrt(dk)
    gemm_f32_batched(q, kp, S, M=N, N=Ksel, K=dk, lda=ldq, ldb=d, ldc=Ksel, alpha=1.0 / scale,
                     sa=(N * ldq, dk), sb=(Ksel * d, dk), sc=s_str, ksplit=1, **hb)
    attn_rows_bwd(S, stats, N, 0, scale, drop, pd=Pd)
    # dV_j = P~_j dO_j          [N, dk]   (B operand dO_j stored [key, c]: k-major rows)
    gemm_f32_batched(Pd, d_o, dv, M=N, N=dk, K=Ksel, lda=Ksel, ldb=d, ldc=2 * d, b_kc=False,
                     sa=s_str, sb=(Ksel * d, dk), sc=(N * 2 * d, dk), ksplit=1, **hb)
    # G = V_j dO_j^T            [N, Ksel]
    gemm_f32_batched(v, d_o, G, M=N, N=Ksel, K=dk, lda=ldv, ldb=d, ldc=Ksel,
                     sa=(N * ldv, dk), sb=(Ksel * d, dk), sc=s_str, ksplit=1, **hb)
    attn_rows_bwd(S, stats, N, 1, scale, drop, g=G)     # G <- dS
    # dQ_j = dS_j Kp_j          [N, dk]
    gemm_f32_batched(G, kp, dq, M=N, N=dk, K=Ksel, lda=Ksel, ldb=d, ldc=2 * d, b_kc=False,
                     sa=s_str, sb=(Ksel * d, dk), sc=(N * 2 * d, dk), ksplit=1, **hb)
    # dKp_j = dS_j^T Q_j        [Ksel, dk], contraction over the N patches -> split-K
    gemm_f32_batched(G, q, dkp, M=Ksel, N=dk, K=N, lda=Ksel, ldb=ldq, ldc=d, a_kc=False, b_kc=False,
                     sa=s_str, sb=(N * ldq, dk), sc=(Ksel * d, dk), **hb)
    return dq, dv, dkp, dqv


def softmax_cols_bwd(a: torch.Tensor, da: torch.Tensor, scale: float) -> torch.Tensor:
    """dS for A = softmax over dim 0 of S / scale (dsmil.py:83-86).  a, da [N, C]."""
    a, da = _f32(a, "a"), _f32(da, "da")
    ds = torch.empty_like(a)
    check(lib.snuffy_softmax_cols_bwd(a.data_ptr(), da.data_ptr(), a.shape[0], a.shape[1], float(scale), ds.data_ptr(),
                                      _stream()), "snuffy_softmax_cols_bwd")
    return ds


# ------------------------------------------------------------------ tensor-core backward operands (csrc/norm.cu, gemm_tc.cu)
def planes_t(x: torch.Tensor, rc: int, *, mode: int = 0, stats=None, gamma=None, beta=None, row_map=None, alt=None,
             act: str = "none", drop: Tuple[float, int, int] = (0.0, 0, 0)) -> Planes:
    """Operand planes of x^T for x [R, C] fp32 (row-strided views allowed): plane rows = columns of x, k = rows of x.
    mode 1: LayerNorm(x through row_map) from saved stats; mode 2: dropout(act(x))."""
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.reshape(-1, x.shape[-1]).contiguous()
    R, C = x.shape
    out = Planes(C, R, rc, x.device)
    check(lib.snuffy_planes_t_fwd(x.data_ptr(), x.stride(0), R, C, rc, mode, _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(row_map),
                                  _ptr(alt), ACT_IDS[act], float(drop[0]), drop[1] & _U64, drop[2] & _U64, out.ptr, out.stride,
                                  _stream()), "snuffy_planes_t_fwd")
    return out


def weight_planes_t(weight: torch.Tensor) -> Planes:
    """B-operand planes of W^T for W [out, in]: rows = input features, k = output features (dX = dY . W)."""
    weight = _f32(weight, "weight")
    return planes_t(weight, _block_n(weight.shape[1]))


def gemm_tc_splitk_rows_supported(M: int, N: int) -> bool:
    return M >= 128 and N >= 128 and M % 128 == 0 and N % _block_n(N) == 0


def gemm_tc_splitk_rows(dy: Planes, x: Planes, *, M: int, N: int, R: int, passes: int = 3) -> torch.Tensor:
    """dW [M, N] = dY^T X from the ROW planes of dY [R, M] and X [R, N] (rc = 128), contracted over the rows through
    MN-major descriptors: no transposed operand copies."""
    if dy.rc != 128 or x.rc != 128 or dy.K != M or x.K != N or dy.rows < R or x.rows < R:
        raise ValueError("gemm_tc_splitk_rows: operands are 128-row planes of [R, M] and [R, N]")
    dev = dy.buf.device
    if R % 16:
        for pl in (dy, x):
            check(lib.snuffy_planes_zero_rows(pl.ptr, pl.stride, pl.K, pl.rc, R, (R + 15) // 16 * 16, _stream()),
                  "snuffy_planes_zero_rows")
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    ks = _tc_auto_ksplit(M, N, R)
    ws_bytes = lib.snuffy_gemm_tc_splitk_workspace(M, N, ks)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    check(lib.snuffy_gemm_tc_splitk_rows(dy.ptr, dy.stride, x.ptr, x.stride, M, N, R, passes, ks, out.data_ptr(), ws.data_ptr(),
                                         ws_bytes, _stream()), "snuffy_gemm_tc_splitk_rows")
    return out


def weight_planes_batch(jobs) -> None:
    """One launch for a list of conversions (csrc/norm.cu weight_planes_batch_kernel).  Each job is a tuple
    (kind, src, dst, dst_row0[, dst_k0]): kind "planes" (dst Planes rows [dst_row0, dst_row0 + src rows) = src), "planes_t"
    (dst Planes rows [dst_row0, ...) = columns of src, k from dst_k0) or "copy" (dst fp32 tensor, dst_row0 = element
    offset)."""
    kinds = {"planes": 0, "planes_t": 1, "copy": 2}
    for start in range(0, len(jobs), MAX_PLANE_JOBS):
        chunk = jobs[start:start + MAX_PLANE_JOBS]
        arr = (PlaneJob * len(chunk))()
        for slot, (kind, src, dst, row0, *k0) in zip(arr, chunk):
            src = _f32(src, "src")
            if src.dim() != 2 or src.stride(1) != 1:
                raise ValueError("weight_planes_batch: sources are row-major fp32 matrices")
            slot.src, slot.ld, slot.rows, slot.cols, slot.kind, slot.dst_row0 = src.data_ptr(), src.stride(0), src.shape[0], \
                src.shape[1], kinds[kind], row0
            if kind == "copy":
                slot.plane_rc, slot.dst, slot.plane_stride, slot.dst_k0, slot.k_total = 0, dst.data_ptr(), 0, 0, 0
            else:
                slot.plane_rc, slot.dst, slot.plane_stride = dst.rc, dst.ptr, dst.stride
                slot.dst_k0, slot.k_total = (k0[0] if k0 else 0), dst.K
        check(lib.snuffy_weight_planes_batch(arr, len(chunk), _stream()), "snuffy_weight_planes_batch")


def gemm_tc_splitk(a: Planes, b: Planes, *, M: int, N: int, K: int, passes: int = 3) -> torch.Tensor:
    """out [M, N] = A . B^T over a long contraction (weight gradients), split over CTAs with a deterministic fold."""
    dev = a.buf.device
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    ks = _tc_auto_ksplit(M, N, K)
    ws_bytes = lib.snuffy_gemm_tc_splitk_workspace(M, N, ks)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    check(lib.snuffy_gemm_tc_splitk(a.ptr, a.stride, b.ptr, b.stride, M, N, K, passes, ks, out.data_ptr(), ws.data_ptr(),
                                    ws_bytes, _stream()), "snuffy_gemm_tc_splitk")
    return out


def select_gather(c: torch.Tensor, x: torch.Tensor, k_top: int, k_rand: int, seed: int = 0, offset: int = 0):
    """Binary-path selection in one launch: c [B, N, 1] scores, x [B, N, d] ->
    (sel [B, k_top + k_rand] int64, flags [B, N] uint8, row_map [B * N] int32, xs [B * Ksel, d])."""
    c = _f32(c, "c")
    x = _f32(x, "x")
    B, N, d = x.shape
    ksel = k_top + k_rand
    dev = x.device
    sel = torch.empty(B, ksel, dtype=torch.int64, device=dev)
    flags = torch.empty(B, N, dtype=torch.uint8, device=dev)
    row_map = torch.empty(B * N, dtype=torch.int32, device=dev)
    xs = torch.empty(B * ksel, d, dtype=torch.float32, device=dev)
    check(lib.snuffy_select_gather(c.data_ptr(), x.data_ptr(), B, N, d, k_top, k_rand, seed & _U64, offset & _U64, sel.data_ptr(),
                                   flags.data_ptr(), row_map.data_ptr(), xs.data_ptr(), _stream()), "snuffy_select_gather")
    return sel, flags, row_map, xs


# ------------------------------------------------------------------ packed variable-length bags (inference)
def select_topk_varlen(c: torch.Tensor, cu: torch.Tensor, B: int, max_n: int, k: int, flags: Optional[torch.Tensor] = None):
    """c [T, C] packed scores -> idx [B, C, k] int64 GLOBAL rows (per bag and class, descending score)."""
    c = _f32(c, "c")
    C = c.shape[-1]
    idx = torch.empty(B, C, k, dtype=torch.int64, device=c.device)
    check(lib.snuffy_select_topk_varlen(c.data_ptr(), cu.data_ptr(), B, max_n, C, k, idx.data_ptr(), _ptr(flags), _stream()),
          "snuffy_select_topk_varlen")
    return idx


def select_random_varlen(flags: torch.Tensor, cu: torch.Tensor, B: int, max_n: int, k: int, seed: int, offset: int):
    idx = torch.empty(B, k, dtype=torch.int64, device=flags.device)
    check(lib.snuffy_select_random_varlen(flags.data_ptr(), cu.data_ptr(), B, max_n, k, seed & _U64, offset & _U64,
                                          idx.data_ptr(), _stream()), "snuffy_select_random_varlen")
    return idx


def sparse_attn_tc_varlen(qv_planes: Planes, kp: torch.Tensor, cu: torch.Tensor, B: int, max_n: int, Ksel: int, h: int, d: int):
    kp = _f32(kp, "kp")
    ws_bytes = lib.snuffy_sparse_attn_tc_workspace(B, max_n, Ksel, h, d)
    if ws_bytes < 0:
        raise ValueError("shape not served by the tensor-core attention kernel")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=kp.device)
    o = torch.empty(B * Ksel, d, dtype=torch.float32, device=kp.device)
    check(lib.snuffy_sparse_attn_tc_varlen_fwd(qv_planes.ptr, qv_planes.stride, qv_planes.K, 0, d, kp.data_ptr(), cu.data_ptr(), B,
                                               max_n, Ksel, h, d, o.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
          "snuffy_sparse_attn_tc_varlen_fwd")
    return o


def ln_mean_head_varlen(x: torch.Tensor, cu: torch.Tensor, B: int, max_n: int, gamma, beta, w_head, b_head) -> torch.Tensor:
    x = _f32(x, "x")
    d = x.shape[-1]
    C = w_head.shape[0]
    chunks = lib.snuffy_ln_mean_head_chunks(B, max_n)
    key = (x.device.index, B)
    tickets = _HEAD_WS.get(key)
    if tickets is None:
        tickets = _HEAD_WS[key] = torch.zeros(B, dtype=torch.int32, device=x.device)
    partials = torch.empty(B * chunks * d, dtype=torch.float32, device=x.device)
    bag = torch.empty(B, C, dtype=torch.float32, device=x.device)
    check(lib.snuffy_ln_mean_head_varlen_fwd(x.data_ptr(), cu.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w_head.data_ptr(),
                                             _ptr(b_head), B, max_n, d, C, partials.data_ptr(), tickets.data_ptr(), None,
                                             bag.data_ptr(), _stream()), "snuffy_ln_mean_head_varlen_fwd")
    return bag


# ------------------------------------------------------------------ attention backward on tensor cores (head-block operands)
def gemm_tc_awindow(a: Planes, a_col0: int, b: Planes, *, M: int, N: int, K: int, passes: int = 3,
                    out: Optional[torch.Tensor] = None, ldc: Optional[int] = None) -> torch.Tensor:
    """out [M, N] = A[:, a_col0 : a_col0 + K] . B^T with A a column window of a wider plane set."""
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.buf.device)
        ldc = N
    check(lib.snuffy_gemm_tc_awindow(a.ptr, a.stride, a.K, a_col0, b.ptr, b.stride, M, N, K, passes, out.data_ptr(), ldc,
                                     _stream()), "snuffy_gemm_tc_awindow")
    return out


def gemm_tc_blockdiag(a: Planes, a_col0: int, b: Planes, *, M: int, N: int, K: int, group_n: int, group_k: int, passes: int = 3,
                      out: Optional[torch.Tensor] = None, ldc: Optional[int] = None) -> torch.Tensor:
    """out [M, N] = A[:, a_col0 : a_col0 + K] . B^T where B is block diagonal (B[n, k] != 0 only for n // group_n == k // group_k):
    each column tile skips the k-blocks that only meet zeros."""
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.buf.device)
        ldc = N
    check(lib.snuffy_gemm_tc_blockdiag(a.ptr, a.stride, a.K, a_col0, b.ptr, b.stride, b.rc, M, N, K, passes, group_n, group_k,
                                       out.data_ptr(), ldc, _stream()), "snuffy_gemm_tc_blockdiag")
    return out


_diag_ksplit = functools.lru_cache(maxsize=None)(lambda M, N, K, rc, dm, dn: lib.snuffy_gemm_tc_diag_ksplit(M, N, K, rc, dm, dn))


def gemm_tc_splitk_blockdiag(a: Planes, b: Planes, *, M: int, N: int, K: int, diag_m: int, diag_n: int, passes: int = 3) -> torch.Tensor:
    """Split-K out [M, N] = A . B^T of which only the blocks with row // diag_m == col // diag_n are computed (the rest of the
    returned tensor is unspecified)."""
    dev = a.buf.device
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    ks = _diag_ksplit(M, N, K, b.rc, diag_m, diag_n)
    ws_bytes = lib.snuffy_gemm_tc_splitk_workspace(M, N, ks)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    check(lib.snuffy_gemm_tc_splitk_blockdiag(a.ptr, a.stride, b.ptr, b.stride, b.rc, M, N, K, passes, ks, diag_m, diag_n,
                                              out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "snuffy_gemm_tc_splitk_blockdiag")
    return out


def block_diag_rows(src: torch.Tensor, h: int) -> torch.Tensor:
    """src [Ksel, d] -> [h*Ksel, d]: row (j, k) = src[k] on head j's columns, zeros elsewhere."""
    src = _f32(src, "src")
    ksel, d = src.shape
    out = torch.empty(h * ksel, d, dtype=torch.float32, device=src.device)
    check(lib.snuffy_block_diag_rows(src.data_ptr(), ksel, h, d, out.data_ptr(), _stream()), "snuffy_block_diag_rows")
    return out


_bwd_fused_ws = functools.lru_cache(maxsize=None)(lambda B, N, Ksel, h, d: lib.snuffy_sparse_attn_bwd_tc_workspace(B, N, Ksel, h, d))


def sparse_attn_bwd_fused_supported(B: int, N: int, Ksel: int, h: int, d: int) -> bool:
    return d % 32 == 0 and _bwd_fused_ws(B, N, Ksel, h, d) >= 0


def sparse_attn_bwd_fused(qvp: Planes, kp: torch.Tensor, d_o: torch.Tensor, stats: torch.Tensor, B: int, N: int, Ksel: int,
                          h: int, d: int, dropout_p: float = 0.0, mask: Optional[torch.Tensor] = None):
    """The attention backward as ONE tcgen05 kernel (csrc/attn_bwd_tc.cu) on the forward's Q|V planes, saved statistics and
    (with dropout) keep bits.  Returns (dQ, dV (column halves of dQV), dKp [B*Ksel, d], dQV [B*N, 2d]) like sparse_attn_bwd."""
    kp, d_o, stats = _f32(kp, "kp"), _f32(d_o, "d_o"), _f32(stats, "stats")
    if dropout_p > 0 and (mask is None or mask.dtype != torch.uint8 or not mask.is_contiguous()
                          or mask.numel() != B * h * N * ((Ksel + 7) // 8)):
        raise ValueError("sparse_attn_bwd_fused: dropout needs the forward's keep-bit mask [B, h, N, ceil(Ksel / 8)] uint8")
    dev = kp.device
    ws_bytes = _bwd_fused_ws(B, N, Ksel, h, d)
    if ws_bytes < 0:
        raise ValueError("shape not served by the fused attention backward")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    dqv = torch.empty(B * N, 2 * d, dtype=torch.float32, device=dev)
    dkp = torch.empty(B * Ksel, d, dtype=torch.float32, device=dev)
    check(lib.snuffy_sparse_attn_bwd_tc(qvp.ptr, qvp.stride, qvp.K, 0, d, kp.data_ptr(), d_o.data_ptr(), stats.data_ptr(), B, N,
                                        Ksel, h, d, float(dropout_p), _ptr(mask) if dropout_p > 0 else None, dqv.data_ptr(),
                                        dkp.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "snuffy_sparse_attn_bwd_tc")
    return dqv[:, :d], dqv[:, d:], dkp, dqv


def sparse_attn_bwd_tc_supported(B: int, N: int, Ksel: int, h: int, d: int) -> bool:
    return d % 32 == 0 and (d // h) % 4 == 0 and (h * Ksel) % 8 == 0 and h * Ksel <= 4096


def sparse_attn_bwd_tc(qvp: Planes, qv: torch.Tensor, kp: torch.Tensor, d_o: torch.Tensor, stats: torch.Tensor, B: int, N: int,
                       Ksel: int, h: int, d: int, drop: Tuple[float, int, int] = (0.0, 0, 0), passes: int = 3):
    """Tensor-core backward of O_j = softmax_keys(Q_j Kp_j^T / sqrt(dk))^T V_j for all heads of a bag at once.

    qvp: the Q|V planes of the forward ([B*N, 2d]); qv: the same in fp32 (for the transposed operand of dKp);
    Returns (dQ, dV (column halves of dQV), dKp [B*Ksel, d], dQV [B*N, 2d]) like sparse_attn_bwd."""
    kp, d_o = _f32(kp, "kp"), _f32(d_o, "d_o")
    dev = kp.device
    dk = d // h
    hk = h * Ksel
    scale = math.sqrt(dk)
    dqv = torch.empty(B * N, 2 * d, dtype=torch.float32, device=dev)
    dkp = torch.empty(B * Ksel, d, dtype=torch.float32, device=dev)
    for b in range(B):
        rows = slice(b * N, (b + 1) * N)
        kbd = block_diag_rows(kp[b * Ksel:(b + 1) * Ksel], h)                     # [h*Ksel, d]
        obd = block_diag_rows(d_o[b * Ksel:(b + 1) * Ksel], h)
        st = stats[b]                                                             # [h, N, 2]
        if (b * N) % 128 == 0:
            a = qvp.row_window(b * N, N)                                          # this bag's rows (whole 128-row tiles)
        else:                                                                     # bag starts inside a plane tile: re-split its rows
            _, a, _ = ln_rows(qv[rows], None, None, apply_ln=False, want_planes=True)
        # the head-block operands are block diagonal (head j's keys x head j's columns): every product below skips the k-blocks
        # that only meet their zeros (gemm_tc_blockdiag), which removes most of the 8x redundant FLOPs of the dense formulation
        S = gemm_tc_blockdiag(a, 0, weight_planes(kbd), M=N, N=hk, K=d, group_n=Ksel, group_k=dk, passes=passes)   # raw Q_j . Kp_j^T
        fused = Ksel % 8 == 0 and Ksel <= 256 and hk % 32 == 0     # the row kernel writes the next product's operand planes itself
        if fused:
            Pd, pdp = None, Planes(N, hk, 128, dev)
            check(lib.snuffy_attn_seg_bwd(S.data_ptr(), st.data_ptr(), N, h, Ksel, b, 0, scale, float(drop[0]), drop[1] & _U64,
                                          drop[2] & _U64, None, None, pdp.ptr, pdp.stride, _stream()), "snuffy_attn_seg_bwd")
        else:
            Pd = torch.empty_like(S)
            check(lib.snuffy_attn_seg_bwd(S.data_ptr(), st.data_ptr(), N, h, Ksel, b, 0, scale, float(drop[0]), drop[1] & _U64,
                                          drop[2] & _U64, Pd.data_ptr(), None, None, 0, _stream()), "snuffy_attn_seg_bwd")
            _, pdp, _ = ln_rows(Pd, None, None, apply_ln=False, want_planes=True)
        # dV = P~ . dObd            [N, d]
        gemm_tc_blockdiag(pdp, 0, planes_t(obd, 128), M=N, N=d, K=hk, group_n=dk, group_k=Ksel, passes=passes,
                          out=dqv[rows, d:], ldc=2 * d)
        del pdp, Pd
        # G = V . dObd^T            [N, h*Ksel]
        G = gemm_tc_blockdiag(a, d, weight_planes(obd), M=N, N=hk, K=d, group_n=Ksel, group_k=dk, passes=passes)
        dsp = Planes(N, hk, 128, dev) if fused else None
        check(lib.snuffy_attn_seg_bwd(S.data_ptr(), st.data_ptr(), N, h, Ksel, b, 1, scale, float(drop[0]), drop[1] & _U64,
                                      drop[2] & _U64, None, G.data_ptr(), dsp.ptr if fused else None, dsp.stride if fused else 0,
                                      _stream()), "snuffy_attn_seg_bwd")                                        # G <- dS
        del S
        # dQ = dS . Kbd             [N, d]
        if not fused:
            _, dsp, _ = ln_rows(G, None, None, apply_ln=False, want_planes=True)
        gemm_tc_blockdiag(dsp, 0, planes_t(kbd, 128), M=N, N=d, K=hk, group_n=dk, group_k=Ksel, passes=passes,
                          out=dqv[rows, :d], ldc=2 * d)
        del dsp
        # dKbd = dS^T . Q           [h*Ksel, d], contraction over the N patches -> split-K; its diagonal blocks are dKp
        dkbd = gemm_tc_splitk_blockdiag(planes_t(G, 128), planes_t(qv[rows, :d], 128), M=hk, N=d, K=N, diag_m=Ksel, diag_n=dk,
                                        passes=passes)
        check(lib.snuffy_block_diag_extract(dkbd.data_ptr(), Ksel, h, d, dkp[b * Ksel:(b + 1) * Ksel].data_ptr(), _stream()),
              "snuffy_block_diag_extract")
    return dqv[:, :d], dqv[:, d:], dkp, dqv


# ------------------------------------------------------------------ patch-level outputs
def patch_probs(scores: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sigmoid of the instance scores (train.py:913-916), optionally written into a caller-owned slice `out`
    (same number of elements, contiguous) of an epoch-wide buffer."""
    s = _f32(scores, "scores")
    if out is None:
        out = torch.empty_like(s)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == s.numel()):
        raise ValueError("patch_probs: `out` must be a contiguous float32 CUDA tensor with scores.numel() elements")
    check(lib.snuffy_patch_probs(s.data_ptr(), s.numel(), out.data_ptr(), _stream()), "snuffy_patch_probs")
    return out


def froc_detections(probs: torch.Tensor, positions: torch.Tensor, threshold: float,
                    cu_seqlens: Optional[torch.Tensor] = None, tile: int = 512, half: int = 256
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Stable per-slide compaction of the patches with probability > threshold (train.py:342-345, 138-141).

    probs [T] or [T, C] (column 0 is used, like the binary caller), positions int32 [T, 2], cu_seqlens int32 [B+1] or None
    (one slide).  Returns (det_prob [T], det_xy [T, 2] int32, count [B] int32): slide b's detections are rows
    cu_seqlens[b] .. cu_seqlens[b] + count[b] of det_*; rows beyond the count are unspecified."""
    p = _f32(probs, "probs")
    T = p.shape[0]
    stride = 1 if p.dim() == 1 else int(p.shape[1])
    if positions.dtype != torch.int32 or not positions.is_cuda or tuple(positions.shape) != (T, 2):
        raise ValueError("froc_detections: `positions` must be an int32 CUDA tensor of shape [T, 2]")
    positions = positions.contiguous()
    if cu_seqlens is not None:
        if cu_seqlens.dtype != torch.int32 or not cu_seqlens.is_cuda or cu_seqlens.dim() != 1 or cu_seqlens.numel() < 2:
            raise ValueError("froc_detections: `cu_seqlens` must be an int32 CUDA tensor [B+1]")
        cu_seqlens = cu_seqlens.contiguous()
        slides = cu_seqlens.numel() - 1
    else:
        slides = 1
    det_prob = torch.empty(T, dtype=torch.float32, device=p.device)
    det_xy = torch.empty(T, 2, dtype=torch.int32, device=p.device)
    count = torch.empty(slides, dtype=torch.int32, device=p.device)
    check(lib.snuffy_froc_detections(p.data_ptr(), stride, positions.data_ptr(), _ptr(cu_seqlens), slides, T, float(threshold),
                                     int(tile), int(half), det_prob.data_ptr(), det_xy.data_ptr(), count.data_ptr(), _stream()),
          "snuffy_froc_detections")
    return det_prob, det_xy, count
