"""Differentiable entry points used by the nn.Modules (forward kernels + their backward kernels).

Each function is a ``torch.autograd.Function`` whose forward AND backward are calls into
``libsnuffy_b200.so``; PyTorch's autograd engine only sequences them (train.py:259 ``loss.backward()``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, ops


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# ------------------------------------------------------------------ instance scores (snuffy.py:39-41)
def scores_fn(feats: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    if _needs_grad(feats, weight, bias):
        from .backward import ScoresFunction
        return ScoresFunction.apply(feats, weight, bias)
    return ops.scores(feats, weight.detach(), None if bias is None else bias.detach())


# ------------------------------------------------------------------ LayerNorm (Encoder.forward's final norm)
def layer_norm_fn(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    if _needs_grad(x, gamma, beta):
        from .backward import LayerNormFunction
        return LayerNormFunction.apply(x, gamma, beta)
    out, _, _ = ops.ln_rows(x.reshape(-1, x.shape[-1]), gamma.detach(), beta.detach(), want_f32=True)
    return out.view(x.shape)


# ------------------------------------------------------------------ final LN + mean + head (snuffy.py:86,71)
def ln_mean_head_fn(x, gamma, beta, w_head, b_head) -> torch.Tensor:
    if _needs_grad(x, gamma, beta, w_head, b_head):
        from .backward import LnMeanHeadFunction
        return LnMeanHeadFunction.apply(x, gamma, beta, w_head, b_head)
    bag, _, _ = ops.ln_mean_head(x, gamma.detach(), beta.detach(), w_head.detach(),
                                 None if b_head is None else b_head.detach())
    return bag


# ------------------------------------------------------------------ one encoder layer (snuffy.py:126-157)
def encoder_layer_fn(layer, x: torch.Tensor, sel: torch.Tensor, zplanes=None):
    """x [B, N, d], sel [B, Ksel] -> (x_next [B, N, d], A [B, h, N, Ksel] or None)."""
    B, N, d = x.shape
    precision = layer._effective_precision()
    heads = layer.self_attn.h
    act = layer.feed_forward.activation_name
    training = layer.training
    p_attn = float(layer.self_attn.dropout.p) if training else 0.0
    p_enc = float(layer.sublayer[0].dropout.p) if training else 0.0
    p_ff = float(layer.feed_forward.dropout.p) if training else 0.0
    if _needs_grad(x, *layer._params()):
        from .backward import EncoderLayerFunction
        return EncoderLayerFunction.apply(layer, x, sel, *layer._params())
    w = layer.layer_weights()
    xc = x.detach().contiguous().view(B * N, d)
    x_next, probs, _ = engine.encoder_layer_forward(xc, B, N, sel, w, heads, act, precision,
                                                    want_probs=layer.return_attn, attn_dropout=p_attn,
                                                    enc_dropout=p_enc, ff_dropout=p_ff, zplanes=zplanes)
    return x_next.view(B, N, d), probs
