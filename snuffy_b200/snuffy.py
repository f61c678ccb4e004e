"""Drop-in replacement for the reference's ``snuffy`` module (binary Snuffy aggregator, B = 1, C = 1).

Same seven class names, constructor signatures and ``state_dict`` keys as /root/reference/snuffy.py:34-238;
``train.py`` can ``import snuffy`` from ``dropin/`` and run unchanged.  All math runs in libsnuffy_b200.so.
"""
from __future__ import annotations

import torch

from . import engine, ops
from ._modules import (BClassifier, Encoder, EncoderLayerBase, FCLayer, IClassifier, MILNet,  # noqa: F401
                       MultiHeadedAttention, PositionwiseFeedForward, SublayerConnection, attention, clones)

# the reference exposes a module-global `device` (snuffy.py:31); kept for callers that read it
device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class EncoderLayer(EncoderLayerBase):
    "Encoder is made up of self-attn and feed forward (snuffy.py:113-157)"

    multiclass = False

    def __init__(self, size, self_attn, feed_forward, dropout, big_lambda, random_patch_share):
        super().__init__()
        self._init_common(size, self_attn, feed_forward, dropout, big_lambda, random_patch_share)

    def _select(self, x, c, state):
        B, N, _ = x.shape
        if c.dim() != 3 or c.shape[0] != 1 or c.shape[2] != 1 or B != 1:
            # the reference breaks at .squeeze()/index_select for B > 1 or C > 1 (snuffy.py:129-131)
            raise ValueError(f"binary snuffy.EncoderLayer needs x [1, N, d] and c [1, N, 1]; got x {tuple(x.shape)}, "
                             f"c {tuple(c.shape)} (use snuffy_multiclass or forward_bags for batches)")
        sel = self.select_fused(c, state, x)
        return sel if sel is not None else self.select(c, state)

    def select_fused(self, c, state, x):
        """The same selection as `select` in ONE launch (csrc/select.cu select_gather_kernel) when the layer input x [B, N, d]
        is at hand and the random rows come from the device sampler: top-k, random-k, the row map and the gathered key rows,
        which the encoder-layer call that follows picks up.  Returns S, or None when this form does not apply (the caller then
        uses `select`): a `select` patched on the instance, a forced selection, the NumPy sampler."""
        if not engine.FUSED_SELECT or "select" in self.__dict__ or self.forced_selection is not None:
            return None
        if not (c.dim() == 3 and c.shape[-1] == 1 and x.dim() == 3 and x.is_cuda and x.dtype == torch.float32
                and x.is_contiguous() and c.shape[:2] == x.shape[:2]):
            return None
        n = c.shape[1]
        kt = min(engine.k_top_of(self.big_lambda, self.random_patch_share), n)
        kr = engine.k_rand_of(self.big_lambda, self.random_patch_share, n)
        if not (1 <= kt <= engine.FUSED_SELECT_MAX_K and kr <= engine.FUSED_SELECT_MAX_K and (kr == 0 or self.random_mode == "device")):
            return None
        seed, offset = engine._RANDOM.next() if kr > 0 else (0, 0)
        sel, flags, row_map, xs = ops.select_gather(c.detach(), x.detach(), kt, kr, seed, offset)
        if state is not None:
            state["top"], state["flags"] = sel[:, :kt], flags
        engine._GATHERED = (sel, xs, row_map)
        return sel

    def select(self, c, state=None):
        """S = T ++ R for c [B, N, 1] (batched extension of snuffy.py:128-147).  T is computed once per forward
        (c is the same for every layer, App. B-10) and cached in `state`; R is re-drawn per layer."""
        state = state if state is not None else {}
        sel, top, flags = engine.select_binary(c.detach(), self.big_lambda, self.random_patch_share, self.random_mode,
                                               top=state.get("top"), flags=state.get("flags"))
        state["top"], state["flags"] = top, flags
        return sel


def forward_bags(milnet: MILNet, x: torch.Tensor):
    """Batched extension: run `milnet` on B same-sized bags x [B, N, d] in one pass (one launch sequence fills
    148 SMs far better than B separate forwards).  Returns (classes [B, N, 1], bag [B, 1], A [B, h, N, Ksel])."""
    zplanes = None
    if x.dim() == 3 and engine.fused_scores_available(milnet, x):
        from . import ops
        lin = milnet.i_classifier.fc[0]
        feats = x.contiguous()
        classes, zplanes = ops.scores_ln_planes(feats, lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None)
    else:
        feats, classes = milnet.i_classifier(x)
    enc = milnet.b_classifier.encoder
    state = {}
    attn = None
    h = feats
    from .autograd import encoder_layer_fn
    for layer in enc.layers:
        # the selection is drawn on the second stream, beside LayerNorm 1 and the Q|V projection (engine.selection_stream)
        side, on_side = engine.selection_stream("fp32" if layer.forced_selection is not None else layer._effective_precision(),
                                                h.device)
        with on_side:
            sel = layer.select_fused(classes, state, h)
            if sel is None:
                sel = layer.forced_selection if layer.forced_selection is not None else layer.select(classes, state)
            sel = sel.to(device=h.device, dtype=torch.int64).contiguous()
        engine._SEL_PENDING = side
        try:
            h, attn = encoder_layer_fn(layer, h, sel, zplanes=zplanes)
        finally:
            engine.join_pending_selection(h.device)
        zplanes = None
    from .autograd import ln_mean_head_fn
    b = milnet.b_classifier
    bag = ln_mean_head_fn(h, enc.norm.weight, enc.norm.bias, b.linear.weight, b.linear.bias)
    return classes, bag, attn


def forward_packed(milnet: MILNet, x: torch.Tensor, cu_seqlens):
    """Packed variable-length extension (BASELINE configs[3]): `x` [T, d] is the concatenation of B bags of different
    lengths, `cu_seqlens` their B + 1 row offsets (list or tensor).  One launch sequence for all bags: every row-wise kernel
    runs on the packed rows; selection, the attention reduction and the mean-pool head are per bag.  Inference only.
    Returns (classes [T, 1], bag [B, 1]).  Every bag needs at least big_lambda patches (shorter bags: call the model)."""
    if x.dim() != 2:
        raise ValueError(f"forward_packed expects the packed bags as [T, d], got {tuple(x.shape)}")
    if torch.is_grad_enabled() and any(p.requires_grad for p in milnet.parameters()):
        raise NotImplementedError("forward_packed is an inference path: wrap the call in torch.no_grad()")
    cu_host = [int(v) for v in (cu_seqlens.tolist() if torch.is_tensor(cu_seqlens) else cu_seqlens)]
    B = len(cu_host) - 1
    lens = [cu_host[i + 1] - cu_host[i] for i in range(B)]
    if B < 1 or cu_host[0] != 0 or cu_host[-1] != x.shape[0] or min(lens) < 1:
        raise ValueError("cu_seqlens must start at 0, end at T and describe non-empty bags")
    T = x.shape[0]
    max_n = max(lens)
    cu = torch.tensor(cu_host, dtype=torch.int64, device=x.device)
    from . import ops
    from .autograd import scores_fn
    lin = milnet.i_classifier.fc[0]
    h = x.detach().contiguous()
    zplanes = None
    if engine.fused_scores_available(milnet, h):
        classes, zplanes = ops.scores_ln_planes(h, lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None)
    else:
        classes = scores_fn(x, lin.weight, lin.bias)                                    # [T, 1]
    enc = milnet.b_classifier.encoder
    top = flags = None
    for layer in enc.layers:
        kt = engine.k_top_of(layer.big_lambda, layer.random_patch_share)
        kr = int(layer.big_lambda * layer.random_patch_share)
        if min(lens) < kt + kr:
            raise ValueError(f"forward_packed: every bag needs >= {kt + kr} patches (shortest has {min(lens)})")
        if layer.forced_selection is not None:                                           # tests / parity: LOCAL rows [B, Ksel]
            sel = layer.forced_selection.to(device=x.device, dtype=torch.int64).view(B, -1) + cu[:-1, None]
        else:
            if top is None:                                                              # c is the same for every layer
                flags = torch.zeros(T, dtype=torch.uint8, device=x.device)
                top = ops.select_topk_varlen(classes.view(T, 1), cu, B, max_n, kt, flags).view(B, kt)
            sel = top
            if kr > 0:
                seed, offset = engine._RANDOM.next()
                sel = torch.cat((top, ops.select_random_varlen(flags, cu, B, max_n, kr, seed, offset)), dim=1)
        ksel = sel.shape[1]
        h, _, _ = engine.encoder_layer_forward(h, 1, T, sel.reshape(1, B * ksel).contiguous(), layer.layer_weights(),
                                               layer.self_attn.h, layer.feed_forward.activation_name,
                                               layer._effective_precision(), want_probs=False,
                                               varlen=(cu, B, ksel, max_n), zplanes=zplanes)
        zplanes = None
    b = milnet.b_classifier
    bag = ops.ln_mean_head_varlen(h, cu, B, max_n, enc.norm.weight.detach(), enc.norm.bias.detach(), b.linear.weight.detach(),
                                  b.linear.bias.detach())
    return classes, bag
