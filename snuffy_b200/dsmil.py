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

    def _q(self, feats: torch.Tensor) -> torch.Tensor:
        if isinstance(self.q, nn.Sequential):
            hdn = ops.linear_f32(feats, self.q[0].weight.detach(), self.q[0].bias.detach(), act="relu")
            return ops.linear_f32(hdn, self.q[2].weight.detach(), self.q[2].bias.detach(), act="tanh")
        return ops.linear_f32(feats, self.q.weight.detach(), self.q.bias.detach())

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
            v = ops.linear_f32(feats, self.v[1].weight.detach(), self.v[1].bias.detach(), act="relu")
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
