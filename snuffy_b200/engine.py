"""Host-side orchestration of the aggregator hot path: which kernel runs when, on what buffers.

The math contract is SURVEY.md Appendix A (reference: snuffy.py:126-157, 160-205, 224-225, 68-86;
snuffy_multiclass.py:130-171).  Nothing here computes on tensors with PyTorch ops — every step is a call
into ``libsnuffy_b200.so`` through :mod:`snuffy_b200.ops`; PyTorch only owns the buffers and the stream.

Data layout in HBM (all fp32 unless noted):
  x        [B*N, d]      layer input, never modified (the reference clones before its scatter)
  sel      [B, Ksel] i64 selected rows S = T ++ R;   row_map [B*N] i32  (-1 | slot)
  xs       [B*Ksel, d]   raw selected rows (the attention KEYS);  xs_new = xs + attention output
  planes   split-bf16 operand planes (hi, lo), pre-tiled for tcgen05 (csrc/common.cuh)
  qv       [B*N, 2d]     Q | V projections of LN1(x)
"""
from __future__ import annotations

import contextlib
import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops
from .ops import Planes

PRECISIONS = ("bf16x3", "fp32", "bf16x1")
#: use the tcgen05 attention kernel when the shape allows (SNUFFY_B200_ATTN=simt forces the fp32 SIMT kernel)
ATTN_TC = os.environ.get("SNUFFY_B200_ATTN", "tc") != "simt"
#: share one set of normalised operand planes between LN1 and LN2 in inference (SNUFFY_B200_SHARE_Z=0 disables)
SHARE_Z = os.environ.get("SNUFFY_B200_SHARE_Z", "1") != "0"


def default_precision() -> str:
    p = os.environ.get("SNUFFY_B200_PRECISION", "bf16x3")
    if p not in PRECISIONS:
        raise ValueError(f"SNUFFY_B200_PRECISION must be one of {PRECISIONS}, got {p!r}")
    return p


# ------------------------------------------------------------------ random-patch stream
class _RandomStream:
    """(seed, offset) pairs for the device sampler: seed follows torch.manual_seed, offset counts draws."""

    def __init__(self):
        self._seed = None
        self._offset = 0
        self._counter = None          # device int64[1] while a CUDA graph is being captured (dp.DataParallelTrainer)
        self._delta = 0

    def begin_indirect(self, counter: torch.Tensor) -> None:
        """Until end_indirect(): draws are (seed | bit 63 | k << 32, address of `counter`) — resolved on the device at replay time
        as (seed & 0xFFFFFFFF, counter + k), include/snuffy_b200.h "Random draws under CUDA-graph replay"."""
        if counter.dtype != torch.int64 or not counter.is_cuda or counter.numel() != 1:
            raise ValueError("begin_indirect: counter must be a CUDA int64 tensor with one element")
        self._counter, self._delta = counter, 0

    def end_indirect(self) -> int:
        n, self._counter, self._delta = self._delta, None, 0
        return n

    def next(self) -> Tuple[int, int]:
        if self._counter is not None:
            self._delta += 1
            if self._delta >= 1 << 31:
                raise RuntimeError("too many random draws in one captured step")
            return (1 << 63) | (self._delta << 32) | (torch.initial_seed() & 0xFFFFFFFF), self._counter.data_ptr()
        seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF               # bit 63 is the library's "indirect draw" flag
        if seed != self._seed:
            self._seed, self._offset = seed, 0
        self._offset += 1
        return seed, self._offset


_RANDOM = _RandomStream()


def k_top_of(big_lambda: int, random_patch_share: float) -> int:
    # Python doubles, exactly as the reference writes it (snuffy.py:124,129)
    return math.ceil(big_lambda * (1.0 - random_patch_share))


def k_rand_of(big_lambda: int, random_patch_share: float, n: int) -> int:
    # snuffy.py:137-140
    return min(int(big_lambda * random_patch_share), max(0, n - k_top_of(big_lambda, random_patch_share)))


# ------------------------------------------------------------------ selection
def select_binary(c: torch.Tensor, big_lambda: int, random_patch_share: float, random_mode: str = "device",
                  top: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Binary-path selection (snuffy.py:128-147) for c [B, N, 1].  Returns (S [B, Ksel], T [B, Ktop], flags).

    T and the taken-flags depend only on c, so callers pass them back in for layers > 0; the random part is
    re-drawn on every call like the reference.  random_mode: "device" (Philox on the GPU, no sync) or "numpy"
    (the reference's NumPy global-RNG stream: bit-identical indices, costs a D2H + H2D round trip)."""
    B, N, C = c.shape
    kt = min(k_top_of(big_lambda, random_patch_share), N)
    kr = k_rand_of(big_lambda, random_patch_share, N)
    if top is None:
        flags = torch.zeros(B, N, dtype=torch.uint8, device=c.device) if kr > 0 else None
        top = ops.select_topk(c, kt, flags).view(B, kt)
    if kr == 0:
        return top, top, flags
    if random_mode == "numpy" and _RANDOM._counter is not None:
        raise RuntimeError("random_mode='numpy' draws on the host and cannot be captured in a CUDA graph; use 'device'")
    if random_mode == "numpy":
        rnd = []
        for b in range(B):                                   # host round trip, reference stream
            taken = top[b].tolist()
            remaining = list(set(range(N)) - set(taken))
            rnd.append(torch.from_numpy(np.random.choice(remaining, kr, replace=False)))
        rand = torch.stack(rnd).to(device=c.device, dtype=torch.int64)
    elif random_mode == "device":
        seed, offset = _RANDOM.next()
        rand = ops.select_random(flags, kr, seed, offset)
    else:
        raise ValueError(f"random_mode must be 'device' or 'numpy', got {random_mode!r}")
    return torch.cat((top, rand), dim=1), top, flags


def select_multiclass(c: torch.Tensor, big_lambda: int, random_patch_share: float, random_mode: str = "device",
                      cache: Optional[dict] = None) -> torch.Tensor:
    """Multiclass selection (snuffy_multiclass.py:130-160) for c [B, N, C] -> S [B, 2*ref] int64."""
    B, N, C = c.shape
    kt = min(k_top_of(big_lambda, random_patch_share), N)
    if cache is not None and "uniq" in cache:
        uniq, ref, flags = cache["uniq"], cache["ref"], cache["flags"]
    else:
        flags = torch.zeros(B, N, dtype=torch.uint8, device=c.device)
        ops.select_topk(c, kt, flags)
        uniq, counts = ops.compact_flags(flags, C * kt)
        if C == 1:
            ref = biggest = kt                               # one class: no duplicates, no sync needed
        else:                                                # shapes depend on it: one sync per forward (both numbers at once)
            ref, biggest = (int(v) for v in torch.stack((counts.min(), counts.max())).tolist())
        ref = min(ref, N - ref)
        if N - biggest < ref:
            # a bag whose union of per-class top sets leaves fewer than `ref` other rows: the reference raises from
            # np.random.choice(..., replace=False) (snuffy_multiclass.py:152-155); never sample flagged rows silently
            raise ValueError(f"snuffy_multiclass: cannot draw {ref} random patches outside a top set of {biggest} rows "
                             f"in a bag of {N} (np.random.choice without replacement would fail in the reference)")
        if cache is not None:
            cache.update(uniq=uniq, ref=ref, flags=flags)
    if ref <= 0:
        raise ValueError(f"snuffy_multiclass: empty selection (N={N}, top share={kt})")
    top = uniq[:, :ref]
    if random_mode == "numpy" and _RANDOM._counter is not None:
        raise RuntimeError("random_mode='numpy' draws on the host and cannot be captured in a CUDA graph; use 'device'")
    if random_mode == "numpy":
        rnd = []
        fl = flags.cpu().numpy()
        for b in range(B):
            remaining = np.nonzero(fl[b] == 0)[0]
            rnd.append(torch.from_numpy(np.random.choice(remaining, ref, replace=False)))
        rand = torch.stack(rnd).to(device=c.device, dtype=torch.int64)
    elif random_mode == "device":
        seed, offset = _RANDOM.next()
        rand = ops.select_random(flags, ref, seed, offset)
    else:
        raise ValueError(f"random_mode must be 'device' or 'numpy', got {random_mode!r}")
    return torch.cat((top, rand), dim=1).contiguous()


# ------------------------------------------------------------------ per-layer weights
@dataclass
class LayerWeights:
    """Views of one EncoderLayer's parameters plus derived operands (fused Q|V weight, tcgen05 planes)."""
    wq: torch.Tensor; bq: torch.Tensor
    wk: torch.Tensor; bk: torch.Tensor
    wv: torch.Tensor; bv: torch.Tensor
    wo: torch.Tensor; bo: torch.Tensor
    w1: torch.Tensor; b1: torch.Tensor
    w2: torch.Tensor; b2: torch.Tensor
    g1: torch.Tensor; be1: torch.Tensor
    g2: torch.Tensor; be2: torch.Tensor
    wqv: Optional[torch.Tensor] = None
    bqv: Optional[torch.Tensor] = None
    wqv_planes: Optional[Planes] = None
    w1_planes: Optional[Planes] = None
    w2_planes: Optional[Planes] = None
    wk_planes: Optional[Planes] = None
    wo_planes: Optional[Planes] = None
    # LayerNorm affine folded into the consumers of the shared z planes:  LN(y) W^T + b = z (W * gamma)^T + (b + W beta)
    wqv_fold_planes: Optional[Planes] = None
    bqv_fold: Optional[torch.Tensor] = None
    w1_fold_planes: Optional[Planes] = None
    b1_fold: Optional[torch.Tensor] = None
    # transposed weights as B operands of the backward dX = dY . W products
    w2t_planes: Optional[Planes] = None
    w1t_planes: Optional[Planes] = None
    wqvt_planes: Optional[Planes] = None
    wot_planes: Optional[Planes] = None
    wkt_planes: Optional[Planes] = None

    def prepare(self, precision: str) -> None:
        if self.wqv is None:
            # one [2d, d] operand so LN1(x) is read once for both projections
            self.wqv = torch.cat((self.wq.detach(), self.wv.detach()), dim=0).contiguous()
            self.bqv = torch.cat((self.bq.detach(), self.bv.detach()), dim=0).contiguous()
        if precision != "fp32" and self.wqv_planes is None:
            self.wqv_planes = ops.weight_planes(self.wqv)
            self.w1_planes = ops.weight_planes(self.w1.detach())
            self.w2_planes = ops.weight_planes(self.w2.detach())
            self.wk_planes = ops.weight_planes(self.wk.detach())
            self.wo_planes = ops.weight_planes(self.wo.detach())

    def prepare_train(self) -> bool:
        """Everything a tensor-core training step derives from the parameters (prepare + prepare_backward) in one launch;
        False when the shapes need the per-operand path (Q and V halves that do not start on a row tile)."""
        if self.wqv_planes is not None and self.w2t_planes is not None:
            return True
        d, dff, dev = self.wq.shape[1], self.w1.shape[0], self.wq.device
        rc_qv, rc_d, rc_ff = ops._block_n(2 * d), ops._block_n(d), ops._block_n(dff)
        if d % rc_qv or d % 32 or dff % 8 or self.wqv_planes is not None or self.w2t_planes is not None:
            return False
        self.bqv = torch.empty(2 * d, dtype=torch.float32, device=dev)
        self.wqv_planes = Planes(2 * d, d, rc_qv, dev)
        self.w1_planes, self.w2_planes = Planes(dff, d, rc_ff, dev), Planes(d, dff, rc_d, dev)
        self.wk_planes, self.wo_planes = Planes(d, d, rc_d, dev), Planes(d, d, rc_d, dev)
        self.wqvt_planes = Planes(d, 2 * d, rc_d, dev)
        self.w1t_planes, self.w2t_planes = Planes(d, dff, rc_d, dev), Planes(dff, d, rc_ff, dev)
        self.wkt_planes, self.wot_planes = Planes(d, d, rc_d, dev), Planes(d, d, rc_d, dev)
        ops.weight_planes_batch([
            ("planes", self.wq, self.wqv_planes, 0), ("planes", self.wv, self.wqv_planes, d),
            ("planes", self.w1, self.w1_planes, 0), ("planes", self.w2, self.w2_planes, 0),
            ("planes", self.wk, self.wk_planes, 0), ("planes", self.wo, self.wo_planes, 0),
            ("planes_t", self.w1, self.w1t_planes, 0), ("planes_t", self.w2, self.w2t_planes, 0),
            ("planes_t", self.wk, self.wkt_planes, 0), ("planes_t", self.wo, self.wot_planes, 0),
            # (Wq|Wv)^T: plane row = input feature, k = the 2d outputs: the k blocks of Wv^T follow those of Wq^T
            ("planes_t", self.wq, self.wqvt_planes, 0, 0), ("planes_t", self.wv, self.wqvt_planes, 0, d),
            ("copy", self.bq.view(1, d), self.bqv, 0), ("copy", self.bv.view(1, d), self.bqv, d)])
        return True

    def prepare_backward(self) -> None:
        if self.w2t_planes is None:
            self.w2t_planes = ops.weight_planes_t(self.w2.detach())
            self.w1t_planes = ops.weight_planes_t(self.w1.detach())
            self.wqvt_planes = ops.weight_planes_t(self.wqv)
            self.wot_planes = ops.weight_planes_t(self.wo.detach())
            self.wkt_planes = ops.weight_planes_t(self.wk.detach())

    def prepare_folded(self) -> None:
        if self.wqv_fold_planes is None:
            d = self.wq.shape[1]
            self.wqv_fold_planes = ops.weight_planes(self.wqv, col_gain=self.g1)
            self.bqv_fold = ops.gemm_f32(self.be1.view(1, d), self.wqv, M=1, N=self.wqv.shape[0], K=d, bias=self.bqv).view(-1)
            self.w1_fold_planes = ops.weight_planes(self.w1.detach(), col_gain=self.g2)
            self.b1_fold = ops.gemm_f32(self.be2.view(1, d), self.w1.detach(), M=1, N=self.w1.shape[0], K=d,
                                        bias=self.b1).view(-1)


@dataclass
class LayerTape:
    """What one layer's forward leaves behind for the backward pass."""
    sel: torch.Tensor
    row_map: torch.Tensor
    xs: torch.Tensor
    xs_new: torch.Tensor
    kp: torch.Tensor
    qv: torch.Tensor
    o: torch.Tensor
    ln1_stats: torch.Tensor
    ln2_stats: torch.Tensor
    attn_stats: torch.Tensor
    h_pre: torch.Tensor
    x_in: torch.Tensor
    drop: Tuple[float, int, int]                 # attention dropout (p, seed, offset)   snuffy.py:166-167
    drop_enc1: Tuple[float, int, int] = (0.0, 0, 0)   # sublayer[0] dropout on the attention output   snuffy.py:108
    drop_ff: Tuple[float, int, int] = (0.0, 0, 0)     # feed-forward hidden dropout                     snuffy.py:225
    drop_enc2: Tuple[float, int, int] = (0.0, 0, 0)   # sublayer[1] dropout on the FFN output           snuffy.py:110
    qvp: Optional[Planes] = None                      # Q|V as operand planes (tensor-core attention backward)
    attn_mask: Optional[torch.Tensor] = None          # keep bits of the attention dropout as the forward drew them
    # row planes the forward GEMMs consumed: LN1(x), LN2(y), dropout(act(h)): X operands of the weight-gradient products
    u1_planes: Optional[Planes] = None
    u2_planes: Optional[Planes] = None
    a_planes: Optional[Planes] = None


#: weight gradients of the three all-row projections straight from row planes (MN-major descriptors); "0" = from transposed copies
DW_BY_ROWS = os.environ.get("SNUFFY_B200_DW_BY_ROWS", "1") != "0"


#: the key projection of a layer on a second stream, beside the Q|V projection; "0" = one stream
FWD_SIDE_STREAM = os.environ.get("SNUFFY_B200_FWD_SIDE_STREAM", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


#: the stream a caller drew the selection on and has not joined yet (set by EncoderLayer.forward around its call)
_SEL_PENDING = None
#: the binary path's selection as one launch (top-k + random-k + row map + key gather); "0" = four launches
FUSED_SELECT = os.environ.get("SNUFFY_B200_FUSED_SELECT", "1") != "0"
FUSED_SELECT_MAX_K = 4096
#: (sel, xs, row_map) of a fused selection, picked up by the encoder-layer call that follows it
_GATHERED = None


def selection_stream(precision: str, device: torch.device):
    """Context for drawing a layer's selection beside LayerNorm 1 / the Q|V projection: forks the second stream and returns
    (stream or None, context manager).  The caller stores the stream in `_SEL_PENDING` around its encoder-layer call."""
    if not FWD_SIDE_STREAM or precision == "fp32" or device.type != "cuda":
        return None, contextlib.nullcontext()
    side = _side_stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    return side, torch.cuda.stream(side)


def join_pending_selection(device: torch.device) -> None:
    global _SEL_PENDING
    if _SEL_PENDING is not None:
        torch.cuda.current_stream(device).wait_stream(_SEL_PENDING)
        _SEL_PENDING = None


def dw_by_rows(d: int, dff: int) -> bool:
    """The training tape keeps the forward's operand planes and the backward contracts them over the rows (backward.py)."""
    return DW_BY_ROWS and ops.gemm_tc_splitk_rows_supported(d, dff) and ops.gemm_tc_splitk_rows_supported(dff, d)


def tc_supported(d: int) -> bool:
    return d % 8 == 0


def fused_scores_available(milnet, x: torch.Tensor) -> bool:
    """True when layer 0 can take its normalised planes from the scoring pass: inference, tensor-core precision, shared
    z planes, a plain FCLayer scorer.  (The planes are only valid for the encoder's first layer: x is its input.)"""
    from ._modules import FCLayer
    layers = milnet.b_classifier.encoder.layers
    if not SHARE_Z or len(layers) == 0 or not isinstance(milnet.i_classifier, FCLayer):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in milnet.parameters())):
        return False
    if not x.is_cuda or x.dtype != torch.float32 or x.shape[-1] % 8 != 0 or x.shape[-1] > 4096:
        return False
    return layers[0]._effective_precision() != "fp32"


def encoder_layer_forward(x: torch.Tensor, B: int, N: int, sel: torch.Tensor, w: LayerWeights, heads: int,
                          activation: str, precision: str, want_probs: bool, save: bool = False,
                          attn_dropout: float = 0.0, enc_dropout: float = 0.0, ff_dropout: float = 0.0,
                          varlen: Optional[Tuple[torch.Tensor, int, int, int]] = None, zplanes: Optional[Planes] = None
                          ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[LayerTape]]:
    """One EncoderLayer (snuffy.py:126-157) on x [B*N, d] with the selection sel [B, Ksel].

    The three dropout rates are 0 in eval mode; in train mode they are the reference's (attention probabilities
    snuffy.py:166-167, both sublayer outputs 108/110, FFN hidden 225), drawn from counter-based masks that the backward
    regenerates.  Returns (x_next [B*N, d], P [B, h, N, Ksel] or None, tape or None).

    varlen = (cu_seqlens [bags+1] int64 on the device, bags, keys per bag, longest bag): x is the packed concatenation of
    variable-length bags, passed as B = 1, N = T with `sel` [1, bags*Ksel] holding GLOBAL rows.  Every row-wise kernel is
    unchanged; only the attention reduction needs the bag boundaries (inference only)."""
    d = x.shape[1]
    rows = B * N
    Ksel = sel.shape[1]
    if precision != "fp32" and not tc_supported(d):
        precision = "fp32"                                   # tcgen05 planes need d % 8 == 0; still the CUDA path
    if varlen is not None:
        cu, bags, ksel_bag, max_n = varlen
        if save or want_probs or precision == "fp32" or attn_dropout > 0 or \
                not ops.sparse_attn_tc_supported(bags, max_n, ksel_bag, heads, d):
            raise NotImplementedError("packed variable-length bags: inference on the tensor-core path only "
                                      "(eval mode, return_attn = False, head size a multiple of 32 and <= 128)")
    if not (save and precision != "fp32" and w.prepare_train()):
        w.prepare(precision)
    passes = 1 if precision == "bf16x1" else 3
    # Inference shares ONE set of normalised planes z = (x - mean) * rstd between LN1 and LN2 (their affines are folded
    # into the weights); the training tape keeps the two LayerNorms separate (their statistics are saved per sub-layer).
    share_z = SHARE_Z and precision != "fp32" and not save

    small_tc = precision != "fp32"
    kside = _side_stream(x.device) if small_tc and FWD_SIDE_STREAM and x.is_cuda else None
    if kside is None:
        join_pending_selection(x.device)                   # the selection was drawn on the second stream: wait for it here
    else:
        kside.wait_stream(torch.cuda.current_stream(x.device))
    global _GATHERED
    fused, _GATHERED = _GATHERED, None
    with torch.cuda.stream(kside) if kside is not None else contextlib.nullcontext():
        if fused is not None and varlen is None and fused[0].data_ptr() == sel.data_ptr() and fused[1].shape == (B * Ksel, d):
            xs, row_map = fused[1], fused[2]                                # written by the selection launch itself
        else:
            xs = ops.gather_rows(x.view(B, N, d), sel).view(B * Ksel, d)    # raw keys (App. B-1)
            row_map = ops.build_row_map(sel, N)

    # The key projection of the Ksel selected rows does not depend on the Q|V projection over all N rows: issued on a second
    # stream (forked here, joined before the attention kernel, inside a captured graph too) it runs in the SMs the big product
    # leaves idle instead of after it.
    kp = None
    if kside is not None:
        with torch.cuda.stream(kside):
            _, xsp, _ = ops.ln_rows(xs, None, None, apply_ln=False, want_planes=True)
            kp, _, _ = ops.gemm_tc(xsp, w.wk_planes, M=B * Ksel, N=d, K=d, passes=passes, bias=w.bk)
            del xsp

    # --- attention sub-layer: u = LN1(x); Q,V over all N rows; keys from the raw selected rows
    attn_tc = False
    if precision == "fp32":
        u, _, ln1_stats = ops.ln_rows(x, w.g1, w.be1, want_f32=True, want_stats=save)
        qv = ops.gemm_f32(u, w.wqv, M=rows, N=2 * d, K=d, bias=w.bqv)
    else:
        if share_z:
            w.prepare_folded()
            if zplanes is not None and zplanes.rows == rows and zplanes.K == d:
                up, ln1_stats = zplanes, None            # layer 0: produced together with the instance scores
            else:
                _, up, ln1_stats = ops.ln_rows(x, None, None, want_planes=True, affine=False)
            wqv_p, bqv = w.wqv_fold_planes, w.bqv_fold
        else:
            _, up, ln1_stats = ops.ln_rows(x, w.g1, w.be1, want_planes=True, want_stats=save)
            wqv_p, bqv = w.wqv_planes, w.bqv
        attn_tc = varlen is not None or (ATTN_TC and ops.sparse_attn_tc_supported(B, N, Ksel, heads, d))
        # the projection writes Q|V straight as the planes the tensor-core attention consumes (fp32 only if saved)
        qv, _, qvp = ops.gemm_tc(up, wqv_p, M=rows, N=2 * d, K=d, passes=passes, bias=bqv,
                                 want_out=save or not attn_tc, want_planes=attn_tc)
    # [B*Ksel, d] key / output projections: a handful of tcgen05 tiles beat the SIMT kernel even for one bag (200 rows:
    # 8 SIMT CTAs looping over K take ~90 us, four tcgen05 CTAs ~10 us)
    if kside is not None:
        global _SEL_PENDING
        torch.cuda.current_stream(x.device).wait_stream(kside)     # selection, gathered keys, row map, key projection
        _SEL_PENDING = None
    elif small_tc:
        _, xsp, _ = ops.ln_rows(xs, None, None, apply_ln=False, want_planes=True)
        kp, _, _ = ops.gemm_tc(xsp, w.wk_planes, M=B * Ksel, N=d, K=d, passes=passes, bias=w.bk)
    else:
        kp = ops.linear_f32(xs, w.wk, w.bk)
    def draw(p):
        return (float(p),) + _RANDOM.next() if p > 0.0 else (0.0, 0, 0)
    drop, drop_enc1, drop_ff, drop_enc2 = draw(attn_dropout), draw(enc_dropout), draw(ff_dropout), draw(enc_dropout)
    attn_mask = None
    if varlen is not None:
        o, probs, attn_stats = ops.sparse_attn_tc_varlen(qvp, kp, cu, bags, max_n, ksel_bag, heads, d), None, None
    elif attn_tc:
        o, probs, attn_stats, attn_mask = ops.sparse_attn_tc(qvp, kp, B, N, Ksel, heads, d, want_probs=want_probs,
                                                             want_stats=save, dropout_p=drop[0], seed=drop[1], offset=drop[2],
                                                             want_mask=True)
    else:
        o, probs, attn_stats = ops.sparse_attn(qv[:, :d], qv[:, d:], kp, B, N, Ksel, heads, want_probs=want_probs,
                                               want_stats=save, dropout_p=drop[0], seed=drop[1], offset=drop[2])
    if small_tc:                                                             # X_S' = X_S + W3 O + b3
        _, opl, _ = ops.ln_rows(o, None, None, apply_ln=False, want_planes=True)
        xs_new, _, _ = ops.gemm_tc(opl, w.wo_planes, M=B * Ksel, N=d, K=d, passes=passes, bias=w.bo, resid=xs,
                                   drop=drop_enc1)
    else:
        xs_new = ops.linear_f32(o, w.wo, w.bo, resid=xs, drop=drop_enc1)

    # --- feed-forward sub-layer over y = x with rows S replaced by X_S' (read through row_map)
    h_pre = None
    if precision == "fp32":
        u2, _, ln2_stats = ops.ln_rows(x, w.g2, w.be2, row_map=row_map, alt=xs_new, want_f32=True, want_stats=save)
        res = ops.gemm_f32(u2, w.w1, M=rows, N=w.w1.shape[0], K=d, bias=w.b1, act=activation, want_preact=save,
                           drop=drop_ff)
        hdn, h_pre = res if save else (res, None)
        x_next = ops.gemm_f32(hdn, w.w2, M=rows, N=d, K=w.w1.shape[0], bias=w.b2, resid=x, row_map=row_map,
                              resid_alt=xs_new, drop=drop_enc2)
    else:
        if share_z:
            ops.ln_rows_scatter_planes(xs_new, sel, N, up)          # only the Ksel rows changed: y[S] = X_S'
            yp, w1_p, b1, ln2_stats = up, w.w1_fold_planes, w.b1_fold, None
        else:
            _, yp, ln2_stats = ops.ln_rows(x, w.g2, w.be2, row_map=row_map, alt=xs_new, want_planes=True, want_stats=save)
            w1_p, b1 = w.w1_planes, w.b1
        dff = w.w1.shape[0]
        # ReLU training: the backward reads its gate off these activated planes (a > 0), so the fp32 pre-activation is not kept
        relu_gate = save and activation == "relu" and not share_z and dw_by_rows(d, dff)
        _, h_pre, hp = ops.gemm_tc(yp, w1_p, M=rows, N=dff, K=d, passes=passes, bias=b1, act=activation,
                                   want_out=False, want_preact=save and not relu_gate, want_planes=True, drop=drop_ff)
        x_next, _, _ = ops.gemm_tc(hp, w.w2_planes, M=rows, N=d, K=dff, passes=passes, bias=w.b2, resid=x,
                                   row_map=row_map, resid_alt=xs_new, drop=drop_enc2)
    tape = None
    if save:
        tape = LayerTape(sel=sel, row_map=row_map, xs=xs, xs_new=xs_new, kp=kp, qv=qv, o=o, ln1_stats=ln1_stats,
                         ln2_stats=ln2_stats, attn_stats=attn_stats, h_pre=h_pre, x_in=x, drop=drop, drop_enc1=drop_enc1,
                         drop_ff=drop_ff, drop_enc2=drop_enc2, qvp=qvp if precision != "fp32" else None, attn_mask=attn_mask)
        if precision != "fp32" and not share_z:
            tape.u1_planes, tape.u2_planes, tape.a_planes = up, yp, hp
    return x_next, probs, tape
