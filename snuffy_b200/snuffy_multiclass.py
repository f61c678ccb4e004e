"""Drop-in replacement for the reference's ``snuffy_multiclass`` module (B >= 1 bags, C >= 1 classes).

Mirrors /root/reference/snuffy_multiclass.py:34-253: per-class top-K -> unique (ascending) -> ref = min |U_b|,
min(ref, N - ref) -> T_b = U_b[:ref], R_b = ref random rows outside U_b (drawn even when r = 0, App. B-18).
"""
from __future__ import annotations

import torch

from . import engine
from ._modules import (BClassifier, Encoder, EncoderLayerBase, FCLayer, IClassifier, MILNet,  # noqa: F401
                       MultiHeadedAttention, PositionwiseFeedForward, SublayerConnection, attention, clones)

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class EncoderLayer(EncoderLayerBase):
    "snuffy_multiclass.py:116-171 (note `num_class` before `dropout` in the signature)"

    multiclass = True

    def __init__(self, size, self_attn, feed_forward, num_class, dropout, big_lambda, random_patch_share):
        super().__init__()
        self._init_common(size, self_attn, feed_forward, dropout, big_lambda, random_patch_share)
        self.num_classes = num_class

    def _select(self, x, c, state):
        if c.dim() != 3 or c.shape[0] != x.shape[0] or c.shape[1] != x.shape[1]:
            raise ValueError(f"snuffy_multiclass.EncoderLayer needs c [B, N, C] matching x; got {tuple(c.shape)}")
        return engine.select_multiclass(c.detach(), self.big_lambda, self.random_patch_share, self.random_mode,
                                        cache=state)
