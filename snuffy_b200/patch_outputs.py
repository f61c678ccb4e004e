"""Patch-level outputs of validation / test on the device (the step after the hot path).

The reference's `valid` loop (train.py:334-355) turns the instance scores of every bag into probabilities
(`torch.sigmoid(ins_prediction.view(-1, 1))`, train.py:913-916), copies them to the host one bag at a time
(`attentions.cpu().numpy()`: train.py:345, 354), builds FROC detection tuples `(prob, x*512+256, y*512+256)` per patch
in a Python list comprehension (train.py:342-345) and filters them by the ROC-optimal threshold in a multiprocessing pool
(`mp_thresholding`, train.py:138-141).  Here:

  parse_positions(strings)            the reference's position regex (train.py:312-320) -> int32 [n, 2] (host, once)
  PatchOutputCollector                probabilities of all bags of an epoch in ONE device buffer (one launch per bag,
                                      no per-bag sync), read back with one copy: the ROC / threshold inputs of
                                      `_calc_feats_metrics` (train.py:363-368)
  froc_detections(...)                the detection tuples above the threshold, compacted per slide on the device,
                                      returned in the reference's format {slide name: [(prob, x, y), ...]}

ROC / AUC / FROC scoring themselves (sklearn, froc.py) stay with the caller: out of scope.
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

_POSITION_RE = re.compile(r'[^\d]*(\d+)[^\d]*(\d+)[^\d]*')          # train.py:313


def parse_positions(positions: Sequence[str]) -> np.ndarray:
    """Patch file names / position strings -> int32 [n, 2]: the first two digit groups, like train.py:312-320."""
    out = np.empty((len(positions), 2), dtype=np.int32)
    for i, s in enumerate(positions):
        m = _POSITION_RE.search(str(s))
        if m is None:
            raise ValueError(f"parse_positions: no two integers in {s!r}")
        out[i, 0], out[i, 1] = int(m.group(1)), int(m.group(2))
    return out


class PatchOutputCollector:
    """Epoch-wide buffers for the per-patch probabilities and the per-bag predictions.

        col = PatchOutputCollector(total_rows, num_bags, num_classes, device)
        for bag in bags:
            classes, bag_logits, _ = milnet(bag)             # or dp.mil_loss(...)'s mixed prediction
            col.add(classes, prediction)                     # one launch, no sync
        probs, preds, cu = col.to_host()                     # ONE device->host copy each
    """

    def __init__(self, total_rows: int, num_bags: int, num_classes: int = 1, device="cuda"):
        self.device = torch.device(device)
        self.C = int(num_classes)
        self.probs = torch.empty(int(total_rows), self.C, dtype=torch.float32, device=self.device)
        self.preds = torch.empty(int(num_bags), self.C, dtype=torch.float32, device=self.device)
        self._offsets = [0]

    @property
    def num_bags(self) -> int:
        return len(self._offsets) - 1

    @property
    def rows(self) -> int:
        return self._offsets[-1]

    def add(self, classes: torch.Tensor, prediction: Optional[torch.Tensor] = None) -> torch.Tensor:
        """classes: instance scores [1, N, C] / [N, C] (logits).  prediction: the bag's mixed prediction [C] (device)."""
        c = classes.detach().reshape(-1, self.C)
        n, start, k = c.shape[0], self._offsets[-1], self.num_bags
        if start + n > self.probs.shape[0] or k >= self.preds.shape[0]:
            raise ValueError("PatchOutputCollector: capacity exceeded")
        view = self.probs[start:start + n]
        ops.patch_probs(c, out=view)
        if prediction is not None:
            self.preds[k].copy_(prediction.detach().reshape(-1), non_blocking=True)
        self._offsets.append(start + n)
        return view

    def cu_seqlens(self) -> torch.Tensor:
        return torch.tensor(self._offsets, dtype=torch.int32, device=self.device)

    def to_host(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        return (self.probs[:self.rows].cpu().numpy(), self.preds[:self.num_bags].cpu().numpy(),
                np.asarray(self._offsets, dtype=np.int64))


def froc_detections(probs: torch.Tensor, positions: torch.Tensor, cu_seqlens: Optional[torch.Tensor], threshold: float,
                    names: Optional[Sequence[str]] = None, tile: int = 512, half: int = 256):
    """{name: [(prob, x, y), ...]} (or a list per slide when `names` is None) — the `detections_dict` that the reference
    hands to its FROC scorer (train.py:384-390), with ONE device->host copy of the kept rows' buffers."""
    det_prob, det_xy, count = ops.froc_detections(probs, positions, threshold, cu_seqlens, tile, half)
    count_h = count.cpu().numpy()
    starts = cu_seqlens.cpu().numpy() if cu_seqlens is not None else np.array([0, probs.shape[0]])
    p_h, xy_h = det_prob.cpu().numpy(), det_xy.cpu().numpy()
    per_slide: List[List[Tuple[float, int, int]]] = []
    for b, k in enumerate(count_h):
        s = int(starts[b])
        per_slide.append([(float(p), int(x), int(y)) for p, (x, y) in zip(p_h[s:s + k], xy_h[s:s + k])])
    if names is None:
        return per_slide
    return {name: det for name, det in zip(names, per_slide)}
