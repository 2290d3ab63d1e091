"""Persistent, device-resident keyframe descriptor database.

Gives BASELINE config 5 (loop-closure query against a keyframe DB) a real producer: the reference has none
(place_recognition.py is empty, loop_closure.py:7-36 matches exactly two frames, todo.txt:7-8 asks for reuse of
previously computed descriptors).  Semantics are those of OpenCV's train collection
``matcher.add([des_kf0, des_kf1, ...]); matcher.knnMatch(q, k=2)``: one "image" per keyframe, results ordered
by ``(distance, imgIdx, trainIdx)`` == global row order, ``DMatch.imgIdx`` = keyframe, ``trainIdx`` = row inside it.

Descriptors are uploaded ONCE when a keyframe is added (append-only ``uint8[capacity, 32]`` CUDA tensor, grown
geometrically); a query then moves only the query descriptors to the device and 17 bytes per query back.
"""
from __future__ import annotations

import numpy as np

from .matcher import REFERENCE_RATIO, _as_desc, knn2


class KeyframeDB:
    def __init__(self, device: int = 0, capacity: int = 1 << 16):
        import torch
        self.device = torch.device("cuda", device)
        self._rows = torch.empty((capacity, 32), dtype=torch.uint8, device=self.device)
        self._n = 0
        self._offsets = [0]          # first global row of every keyframe, + total

    def __len__(self) -> int:
        return len(self._offsets) - 1

    @property
    def n_rows(self) -> int:
        return self._n

    def add(self, descriptors) -> int:
        """Append one keyframe's descriptors (numpy uint8[n,32] or CUDA tensor); returns its keyframe id."""
        import torch
        d = descriptors if hasattr(descriptors, "is_cuda") else torch.from_numpy(_as_desc(descriptors, "descriptors"))
        n = int(d.shape[0])
        if self._n + n > self._rows.shape[0]:
            cap = max(2 * self._rows.shape[0], self._n + n)
            grown = torch.empty((cap, 32), dtype=torch.uint8, device=self.device)
            grown[: self._n] = self._rows[: self._n]
            self._rows = grown
        self._rows[self._n:self._n + n].copy_(d, non_blocking=True)
        self._n += n
        self._offsets.append(self._n)
        return len(self._offsets) - 2

    def rows(self):
        """The resident collection, uint8[n_rows, 32] CUDA view."""
        return self._rows[: self._n]

    def locate(self, global_idx):
        """global train row -> (keyframe id, row inside the keyframe); -1 stays -1."""
        g = np.asarray(global_idx)
        off = np.asarray(self._offsets)
        kf = np.searchsorted(off, np.maximum(g, 0), side="right") - 1
        loc = g - off[kf]
        kf = np.where(g < 0, -1, kf)
        loc = np.where(g < 0, -1, loc)
        return kf, loc

    def query(self, q, ratio=REFERENCE_RATIO, cross_check: bool = False):
        """kNN-2 of ``q`` (numpy or CUDA uint8[nq,32]) against every stored descriptor.
        Returns ``idx`` (global rows), ``dist``, ``accept`` as numpy arrays for numpy queries, CUDA tensors otherwise."""
        import torch
        host = not hasattr(q, "is_cuda")
        qd = torch.from_numpy(_as_desc(q, "queryDescriptors")).to(self.device, non_blocking=True) if host else q
        idx, dist, acc = knn2(qd, self.rows(), ratio=ratio, cross_check=cross_check)
        if host:
            return idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy()
        return idx, dist, acc
