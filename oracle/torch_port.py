"""CPU port of the reference's binary Snuffy forward in PyTorch CPU ops — the timed CPU baseline.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/snuffy_oracle.py header): imported by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs and by tests; never by ``snuffy_b200``.

``/root/reference`` does not exist on the GPU box, so the reference module itself cannot be timed there.
This port keeps the reference's ATen op sequence and its costs: a full descending sort of
the N scores (snuffy.py:128), the selection + gather done twice (131-147 and 103-106), LayerNorm over all N
tokens, three Linear projections, materialised [h, N, K] scores / softmax, P^T V (160-168), the N x d clone
+ index_put scatter (152-155), the FFN over all N tokens (224-225), final LayerNorm, mean, head (86, 71).
It is validated against the golden fixtures in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_ACT = {
    "relu": F.relu,
    "gelu": F.gelu,
    "leakyrelu": lambda t: F.leaky_relu(t, 0.01),
    "selu": F.selu,
}


def _lin(x, params, key):
    return F.linear(x, params[key + ".weight"], params[key + ".bias"])


def forward(x: torch.Tensor, params: Dict[str, torch.Tensor], heads: int, big_lambda: int, random_patch_share: float,
            depth: int, activation: str = "relu", selections: Optional[Sequence[np.ndarray]] = None):
    """x [1, N, d] fp32 CPU tensor -> (classes [1, N, 1], bag [1, 1], A [1, h, N, Ksel])."""
    act = _ACT[activation]
    n, d = x.shape[1], x.shape[2]
    dk = d // heads
    c = _lin(x, params, "i_classifier.fc.0")
    k_top = math.ceil(big_lambda * (1.0 - random_patch_share))
    attn = None
    for l in range(depth):
        pre = f"b_classifier.encoder.layers.{l}."
        if selections is not None:
            sel = torch.as_tensor(np.asarray(selections[l]), dtype=torch.int64)
        else:
            _, order = torch.sort(c, 1, descending=True)
            top = order[:, 0:k_top, :].squeeze()
            if top.dim() == 0:
                top = top.unsqueeze(0)
            k_rand = min(int(big_lambda * random_patch_share), max(0, n - k_top))
            if k_rand:
                remaining = list(set(range(n)) - set(top.tolist()))
                rnd = torch.from_numpy(np.random.choice(remaining, k_rand, replace=False))
                sel = torch.hstack((top, rnd))
            else:
                sel = top
        keys_raw = torch.index_select(x, 1, sel)
        keys_raw = torch.index_select(x, 1, sel)                       # the reference gathers twice
        u = F.layer_norm(x, (d,), params[pre + "sublayer.0.norm.weight"], params[pre + "sublayer.0.norm.bias"])
        q = _lin(u, params, pre + "self_attn.linears.0").view(1, -1, heads, dk).transpose(1, 2)
        kp = _lin(keys_raw, params, pre + "self_attn.linears.1").view(1, -1, heads, dk).transpose(1, 2)
        v = _lin(u, params, pre + "self_attn.linears.2").view(1, -1, heads, dk).transpose(1, 2)
        scores = torch.matmul(q, kp.transpose(-2, -1)) / math.sqrt(dk)
        attn = scores.softmax(dim=-1)
        o = torch.matmul(attn.transpose(-2, -1), v)
        o = o.transpose(1, 2).contiguous().view(1, -1, heads * dk)
        x_sel = keys_raw + _lin(o, params, pre + "self_attn.linears.3")
        y = x.clone()
        y[:, sel, :] = x_sel
        u2 = F.layer_norm(y, (d,), params[pre + "sublayer.1.norm.weight"], params[pre + "sublayer.1.norm.bias"])
        x = y + _lin(act(_lin(u2, params, pre + "feed_forward.w_1")), params, pre + "feed_forward.w_2")
    z = F.layer_norm(x, (d,), params["b_classifier.encoder.norm.weight"], params["b_classifier.encoder.norm.bias"])
    bag = _lin(z.mean(dim=1), params, "b_classifier.linear")
    return c, bag, attn
