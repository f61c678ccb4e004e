"""Binary bag store and pinned-memory prefetcher: the data formats either side of the hot path.

The reference keeps every slide's patch embeddings as a ``%.4f`` CSV written by compute_feats.py:256-266 and parses them with
pandas on every run (utils.py:138-241, minutes of ``read_csv`` for a dataset).  The store is the same content in the layout
the aggregator consumes: one flat little-endian fp32 file of all rows, a JSON index (per-bag row offsets, labels, slide names,
optional per-patch labels / positions), read through ``numpy.memmap`` straight into pinned staging buffers.

  write_store(path, bags, labels, names)           list of [N_i, d] float32 arrays -> <path>.bin + <path>.json
  csv_to_store(bags_csv, path, num_classes)        the reference's two-level CSV layout -> store (utils.py:138-183)
  BagStore(path)                                   len / bag(i) / label(i) / lengths
  PinnedPrefetcher(store, order, device)           a staging thread + copy stream keep the next bags in flight (memmap ->
                                                   pinned -> device) while the current one is computed; no host-side wait on
                                                   the consumer's kernels (the reference: pageable .to(device) per bag, train.py:255)

Host-side code: no CUDA kernels here; the prefetcher only issues cudaMemcpyAsync through torch.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = "snuffy_b200.bagstore.v1"


def write_store(path: str, bags: Sequence[np.ndarray], labels: Sequence, names: Optional[Sequence[str]] = None,
                patch_labels: Optional[Sequence[Optional[np.ndarray]]] = None,
                positions: Optional[Sequence[Optional[Sequence]]] = None) -> None:
    """bags[i]: [N_i, d] array (cast to float32); labels[i]: scalar or [C] vector."""
    if len(bags) != len(labels) or not bags:
        raise ValueError("write_store: need one label per bag and at least one bag")
    d = int(np.asarray(bags[0]).shape[1])
    offsets = [0]
    with open(path + ".bin", "wb") as f:
        for b in bags:
            a = np.ascontiguousarray(np.asarray(b), dtype="<f4")
            if a.ndim != 2 or a.shape[1] != d:
                raise ValueError(f"write_store: every bag must be [N, {d}], got {a.shape}")
            a.tofile(f)
            offsets.append(offsets[-1] + a.shape[0])
    index = {"magic": MAGIC, "d": d, "offsets": offsets,
             "labels": [np.asarray(l, dtype=np.float32).reshape(-1).tolist() for l in labels],
             "names": list(names) if names is not None else [str(i) for i in range(len(bags))]}
    if patch_labels is not None:
        index["patch_labels"] = [None if p is None else np.asarray(p, dtype=np.float32).tolist() for p in patch_labels]
    if positions is not None:
        index["positions"] = [None if p is None else [str(q) for q in p] for p in positions]
    with open(path + ".json", "w") as f:
        json.dump(index, f)


def csv_to_store(bags_csv: str, path: str, num_classes: int = 1, path_rewrite=None) -> int:
    """Convert the reference's dataset listing (column 0: per-bag CSV of embeddings, column 1: label; utils.py:138-183,
    WITHOUT its row shuffle) into a store.  Returns the number of bags.  Needs pandas (host-side, one-off)."""
    import pandas as pd
    listing = pd.read_csv(bags_csv)
    bags, labels, names, plabels, ppos = [], [], [], [], []
    for _, row in listing.iterrows():
        csv_path = str(row.iloc[0])
        if path_rewrite is not None:
            csv_path = path_rewrite(csv_path)
        df = pd.read_csv(csv_path)
        has = "position" in df and "label" in df
        feats = df.drop(columns=["label", "position"]) if has else df
        bags.append(feats.to_numpy(dtype=np.float32))
        label = np.zeros(num_classes, dtype=np.float32)
        if num_classes == 1:
            label[0] = row.iloc[1]
        elif int(row.iloc[1]) <= num_classes - 1:
            label[int(row.iloc[1])] = 1
        labels.append(label)
        names.append(os.path.splitext(os.path.basename(csv_path))[0])
        plabels.append(df["label"].to_numpy() if has else None)
        ppos.append(list(df["position"]) if has else None)
    write_store(path, bags, labels, names, plabels if any(p is not None for p in plabels) else None,
                ppos if any(p is not None for p in ppos) else None)
    return len(bags)


class BagStore:
    def __init__(self, path: str):
        with open(path + ".json") as f:
            self.index = json.load(f)
        if self.index.get("magic") != MAGIC:
            raise ValueError(f"{path}.json is not a {MAGIC} index")
        self.d = int(self.index["d"])
        self.offsets = np.asarray(self.index["offsets"], dtype=np.int64)
        rows = int(self.offsets[-1])
        size = os.path.getsize(path + ".bin")
        if size != rows * self.d * 4:
            raise ValueError(f"{path}.bin has {size} bytes, the index describes {rows * self.d * 4}")
        self.data = np.memmap(path + ".bin", dtype="<f4", mode="r", shape=(rows, self.d))

    def __len__(self) -> int:
        return len(self.offsets) - 1

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def bag(self, i: int) -> np.ndarray:
        return self.data[self.offsets[i]:self.offsets[i + 1]]

    def label(self, i: int) -> np.ndarray:
        return np.asarray(self.index["labels"][i], dtype=np.float32)

    def name(self, i: int) -> str:
        return self.index["names"][i]


class PinnedPrefetcher:
    """Iterates (slide id, bag [1, N, d] on the device, label [1, C] on the device) over `order`.

    A background thread stages the bags AHEAD of the consumer: memmap -> pinned buffer (host memcpy, GIL released by numpy),
    then cudaMemcpyAsync on a copy stream.  Nothing on the consumer's side ever blocks the host:
      * a pinned buffer is reused once ITS OWN previous H2D copy has completed (`_copied`, waited on by the staging thread —
        a DMA, not the consumer's compute);
      * a device buffer is reused once the consumer's kernels on it have run — a GPU-side `wait_event(_free)` on the copy
        stream, not a host synchronize;
      * the consumer's stream waits (GPU side) for `_ready` of the slot it is handed.
    So while the GPU computes on bag k, bags k+1 .. k+slots-1 are being read, staged and copied (the reference does a pageable
    `.to(device)` per bag on the training thread, train.py:255-256)."""

    def __init__(self, store: BagStore, order: Iterable[int], device, max_rows: Optional[int] = None, slots: int = 3):
        self.store, self.order, self.device = store, list(order), torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())      # the staging thread needs an explicit index
        if slots < 2:
            raise ValueError("PinnedPrefetcher needs at least 2 slots")
        cap = int(max_rows if max_rows is not None else (max(store.lengths[self.order]) if self.order else 0))
        c = len(store.index["labels"][0])
        self.slots = slots
        self._pinned = [torch.empty(cap, store.d, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self._pinned_lab = [torch.empty(1, c, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self._dev = [torch.empty(cap, store.d, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self._lab = [torch.empty(1, c, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self._copy = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in range(slots)]     # H2D of the slot done       (consumer stream waits)
        self._copied = [torch.cuda.Event() for _ in range(slots)]    # same instant, for the host (staging thread waits)
        self._free = [torch.cuda.Event() for _ in range(slots)]      # consumer's work on the slot enqueued (copy stream waits)
        self.h2d_bytes = 0

    def _stage(self, slot: int, i: int) -> int:
        bag = self.store.bag(i)
        n = bag.shape[0]
        self._copied[slot].synchronize()                     # pinned buffer: its previous H2D has finished (never the compute)
        self._pinned[slot][:n].numpy()[...] = bag            # memmap -> pinned (the only host copy)
        self._pinned_lab[slot].numpy()[...] = self.store.label(i).reshape(1, -1)
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(self._free[slot])          # device buffer: GPU-side wait for the consumer's kernels
            self._dev[slot][:n].copy_(self._pinned[slot][:n], non_blocking=True)
            self._lab[slot].copy_(self._pinned_lab[slot], non_blocking=True)
            self._ready[slot].record(self._copy)
            self._copied[slot].record(self._copy)
        self.h2d_bytes += n * self.store.d * 4
        return n

    def __iter__(self) -> Iterator[Tuple[int, torch.Tensor, torch.Tensor]]:
        import queue
        import threading
        if not self.order:
            return
        free_q: "queue.Queue" = queue.Queue()
        ready_q: "queue.Queue" = queue.Queue()
        for s in range(self.slots):
            free_q.put(s)
        stop = threading.Event()

        def stager():
            try:
                torch.cuda.set_device(self.device)
                for i in self.order:
                    slot = free_q.get()
                    if stop.is_set() or slot is None:
                        return
                    ready_q.put((i, slot, self._stage(slot, i)))
                ready_q.put(None)
            except BaseException as exc:                     # surfaces in the consumer
                ready_q.put(exc)

        worker = threading.Thread(target=stager, name="snuffy-prefetch", daemon=True)
        worker.start()
        cur = torch.cuda.current_stream(self.device)
        try:
            while True:
                item = ready_q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                i, slot, n = item
                cur.wait_event(self._ready[slot])
                yield i, self._dev[slot][:n].view(1, n, self.store.d), self._lab[slot]
                self._free[slot].record(cur)                 # work enqueued by the caller on `cur` has consumed the buffers
                free_q.put(slot)
        finally:
            stop.set()
            free_q.put(None)
            worker.join(timeout=30)
