"""Drop-in replacement for the reference's ``dsmil`` module (/root/reference/dsmil.py:28-106).

FCLayer / IClassifier / BClassifier / MILNet with the same constructor signatures and ``state_dict`` keys
(``b_classifier.q.{0,2}.*``, ``b_classifier.v.1.*``, ``b_classifier.fcc.*``).  2-D inputs [N, d], no batch dim.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._modules import FCLayer, IClassifier, _require_cuda  # noqa: F401  (same classes as snuffy's)


class BClassifier(nn.Module):
    """Critical-instance attention pooling (dsmil.py:53-92)."""

    def __init__(self, input_size, output_class, dropout_v=0.0, nonlinear=True, passing_v=False):
        super().__init__()
        if nonlinear:
            self.q = nn.Sequential(nn.Linear(input_size, 128), nn.ReLU(), nn.Linear(128, 128), nn.Tanh())
        else:
            self.q = nn.Linear(input_size, 128)
        if passing_v:
            self.v = nn.Sequential(nn.Dropout(dropout_v), nn.Linear(input_size, input_size), nn.ReLU())
        else:
            self.v = nn.Identity()
        self.fcc = nn.Conv1d(output_class, output_class, kernel_size=input_size)

    #: "bf16x3" (default: the q / v MLPs on the tcgen05 split-bf16 GEMM, fp32 to ~2^-16) | "fp32" (SIMT, exact)
    precision = "bf16x3"

    def _planes(self, lin: nn.Linear):
        """B-operand planes of a Linear weight, cached until the parameter is replaced or updated in place."""
        cache = self.__dict__.setdefault("_wplanes", {})
        key = (lin.weight.data_ptr(), lin.weight._version)
        hit = cache.get(id(lin))
        if hit is None or hit[0] != key:
            cache[id(lin)] = hit = (key, ops.weight_planes(lin.weight.detach()))
        return hit[1]

    def _tc(self, feats: torch.Tensor) -> bool:
        # worth it (and exercised) for real bags only: thousands of rows, feature size a multiple of the 32-column plane block
        return self.precision != "fp32" and feats.shape[1] % 32 == 0 and feats.shape[0] >= 1024

    def _mlp(self, x, layers):
        """[(Linear, activation name), ...] applied to x [rows, K]: tcgen05 GEMMs with the activation in the epilogue (each
        layer hands its output to the next as operand planes), or the exact-fp32 SIMT GEMM."""
        if not self._tc(x):
            for lin, act in layers:
                x = ops.linear_f32(x, lin.weight.detach(), lin.bias.detach(), act=act)
            return x
        rows = x.shape[0]
        _, xp, _ = ops.ln_rows(x, None, None, apply_ln=False, want_planes=True)
        out = None
        for i, (lin, act) in enumerate(layers):
            last = i == len(layers) - 1
            out, _, xp = ops.gemm_tc(xp, self._planes(lin), M=rows, N=lin.weight.shape[0], K=lin.weight.shape[1], passes=3,
                                     bias=lin.bias.detach(), act=act, want_out=last, want_planes=not last)
        return out

    def _q(self, feats: torch.Tensor) -> torch.Tensor:
        if isinstance(self.q, nn.Sequential):
            return self._mlp(feats, [(self.q[0], "relu"), (self.q[2], "tanh")])
        return self._mlp(feats, [(self.q, "none")])

    def forward(self, feats, c):  # N x K, N x C
        _require_cuda(feats, "dsmil.BClassifier")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .backward import dsmil_bclassifier_fn
            return dsmil_bclassifier_fn(self, feats, c)
        feats = feats.detach().contiguous()
        c = c.detach().contiguous()
        n, d = feats.shape
        ncls = c.shape[1]
        if isinstance(self.v, nn.Sequential):
            if self.training and self.v[0].p > 0:
                raise NotImplementedError("dsmil value dropout in train mode needs the autograd path")
            v = self._mlp(feats, [(self.v[1], "relu")])
        else:
            v = feats
        q = self._q(feats)
        # critical instance per class = first row of the descending sort (dsmil.py:78-81): top-1 selection
        crit = ops.select_topk(c.view(1, n, ncls), 1).view(1, ncls)
        # q_max = q(feats[crit]) (dsmil.py:80-82) is row crit of Q: the same MLP applied to the same row -> gather it
        q_max = ops.gather_rows(q.view(1, n, q.shape[1]), crit).view(ncls, q.shape[1])
        a, bm, logits = ops.dsmil_pool(q, q_max, v, self.fcc.weight.detach(), self.fcc.bias.detach())
        return logits.view(1, -1), a, bm.view(1, ncls, d)


class MILNet(nn.Module):
    """dsmil.py:95-106: returns (classes [N, C], prediction_bag [1, C], A [N, C])."""

    def __init__(self, i_classifier, b_classifier):
        super().__init__()
        self.i_classifier = i_classifier
        self.b_classifier = b_classifier

    def forward(self, x):
        x = x.view(-1, self.i_classifier.in_size)
        feats, classes = self.i_classifier(x)
        prediction_bag, A, B = self.b_classifier(feats, classes)
        return classes, prediction_bag, A
