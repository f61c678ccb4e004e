"""Drop-in replacement for the reference's ``snuffy`` module (binary Snuffy aggregator, B = 1, C = 1).

Same seven class names, constructor signatures and ``state_dict`` keys as /root/reference/snuffy.py:34-238;
``train.py`` can ``import snuffy`` from ``dropin/`` and run unchanged.  All math runs in libsnuffy_b200.so.
"""
from __future__ import annotations

import torch

from . import engine
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
        return self.select(c, state)

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
    feats, classes = milnet.i_classifier(x)
    enc = milnet.b_classifier.encoder
    state = {}
    attn = None
    h = feats
    for layer in enc.layers:
        sel = layer.forced_selection if layer.forced_selection is not None else layer.select(classes, state)
        from .autograd import encoder_layer_fn
        h, attn = encoder_layer_fn(layer, h, sel.to(device=h.device, dtype=torch.int64).contiguous())
    from .autograd import ln_mean_head_fn
    b = milnet.b_classifier
    bag = ln_mean_head_fn(h, enc.norm.weight, enc.norm.bias, b.linear.weight, b.linear.bias)
    return classes, bag, attn
