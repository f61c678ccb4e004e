"""Deterministic synthetic weights / bags shared by the golden generator, the
tests and bench.py (TEST INFRASTRUCTURE — see oracle/snuffy_oracle.py header).

Everything is drawn from ``numpy.random.RandomState`` (the legacy MT19937
stream, which NumPy guarantees to be stable across versions), so a fixture only
has to store (config, seed) and the reference's outputs — not the multi-MB
weights.  Key names are the reference's ``state_dict`` keys.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _xavier_normal(rs: np.random.RandomState, out_f: int, in_f: int, extra: int = 1) -> np.ndarray:
    fan_in, fan_out = in_f * extra, out_f * extra
    std = np.sqrt(2.0 / (fan_in + fan_out))
    shape = (out_f, in_f) if extra == 1 else (out_f, in_f, extra)
    return (rs.standard_normal(shape) * std).astype(np.float32)


def _bias(rs, n, zero):
    return np.zeros(n, np.float32) if zero else (0.1 * rs.standard_normal(n)).astype(np.float32)


def make_snuffy_params(d: int, depth: int, num_classes: int = 1, mlp_multiplier: int = 4,
                       seed: int = 0, realistic: bool = False) -> Dict[str, np.ndarray]:
    """state_dict-shaped weights for snuffy / snuffy_multiclass MILNet.

    realistic=True reproduces train.py's init (xavier-normal 2-D weights, zero
    biases, LayerNorm (1,0)); False perturbs biases and LN affine so that tests
    exercise every term.
    """
    rs = np.random.RandomState(seed)
    p: Dict[str, np.ndarray] = {}
    p["i_classifier.fc.0.weight"] = _xavier_normal(rs, num_classes, d)
    p["i_classifier.fc.0.bias"] = _bias(rs, num_classes, realistic)
    dff = d * mlp_multiplier
    for l in range(depth):
        pre = f"b_classifier.encoder.layers.{l}."
        for i in range(4):
            p[pre + f"self_attn.linears.{i}.weight"] = _xavier_normal(rs, d, d)
            p[pre + f"self_attn.linears.{i}.bias"] = _bias(rs, d, realistic)
        p[pre + "feed_forward.w_1.weight"] = _xavier_normal(rs, dff, d)
        p[pre + "feed_forward.w_1.bias"] = _bias(rs, dff, realistic)
        p[pre + "feed_forward.w_2.weight"] = _xavier_normal(rs, d, dff)
        p[pre + "feed_forward.w_2.bias"] = _bias(rs, d, realistic)
        for s in range(2):
            if realistic:
                p[pre + f"sublayer.{s}.norm.weight"] = np.ones(d, np.float32)
                p[pre + f"sublayer.{s}.norm.bias"] = np.zeros(d, np.float32)
            else:
                p[pre + f"sublayer.{s}.norm.weight"] = (1 + 0.1 * rs.standard_normal(d)).astype(np.float32)
                p[pre + f"sublayer.{s}.norm.bias"] = (0.1 * rs.standard_normal(d)).astype(np.float32)
    if realistic:
        p["b_classifier.encoder.norm.weight"] = np.ones(d, np.float32)
        p["b_classifier.encoder.norm.bias"] = np.zeros(d, np.float32)
    else:
        p["b_classifier.encoder.norm.weight"] = (1 + 0.1 * rs.standard_normal(d)).astype(np.float32)
        p["b_classifier.encoder.norm.bias"] = (0.1 * rs.standard_normal(d)).astype(np.float32)
    p["b_classifier.linear.weight"] = _xavier_normal(rs, num_classes, d)
    p["b_classifier.linear.bias"] = _bias(rs, num_classes, realistic)
    return p


def make_dsmil_params(d: int, num_classes: int = 1, nonlinear: bool = True,
                      passing_v: bool = False, seed: int = 0) -> Dict[str, np.ndarray]:
    """state_dict-shaped weights for dsmil.MILNet(FCLayer, BClassifier)."""
    rs = np.random.RandomState(seed)
    p: Dict[str, np.ndarray] = {}
    p["i_classifier.fc.0.weight"] = _xavier_normal(rs, num_classes, d)
    p["i_classifier.fc.0.bias"] = _bias(rs, num_classes, False)
    if nonlinear:
        p["b_classifier.q.0.weight"] = _xavier_normal(rs, 128, d)
        p["b_classifier.q.0.bias"] = _bias(rs, 128, False)
        p["b_classifier.q.2.weight"] = _xavier_normal(rs, 128, 128)
        p["b_classifier.q.2.bias"] = _bias(rs, 128, False)
    else:
        p["b_classifier.q.weight"] = _xavier_normal(rs, 128, d)
        p["b_classifier.q.bias"] = _bias(rs, 128, False)
    if passing_v:
        p["b_classifier.v.1.weight"] = _xavier_normal(rs, d, d)
        p["b_classifier.v.1.bias"] = _bias(rs, d, False)
    p["b_classifier.fcc.weight"] = _xavier_normal(rs, num_classes, num_classes, d)
    p["b_classifier.fcc.bias"] = _bias(rs, num_classes, False)
    return p


def make_bag(n: int, d: int, seed: int = 1234, batch: int = 1, l2: bool = False) -> np.ndarray:
    """Synthetic bag(s) x ~ N(0,1), fp32, [batch, n, d]."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((batch, n, d)).astype(np.float32)
    if l2:
        x /= np.linalg.norm(x, axis=-1, keepdims=True)
    return x
