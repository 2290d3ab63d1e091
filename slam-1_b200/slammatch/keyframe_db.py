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


class ShardedKeyframeDB:
    """The keyframe DB spread over the ranks of one box: config 5's multi-GPU producer.

    SPMD: every rank calls ``add(descriptors)`` with the same keyframe, keyframe k is STORED only on rank ``k % world``
    (round-robin keeps the shards balanced while the map grows), and every rank calls ``query(q)`` with the same query
    batch.  A keyframe's global rows are its position in the overall append order -- exactly OpenCV's
    ``add([...])`` collection order -- so a rank holds several runs of global rows back to back in one local array.  Local
    order == global order inside a rank, hence the rank-local top-2 by (distance, local row) is the rank's top-2 by
    (distance, global row); the keys are then rebased segment by segment to global rows, exchanged like any other
    sharded query (NCCL all-gather + slm_merge_top2, or the NVLink exchange through slm_exchange_merge) and merged:
    the result is byte-identical to the single-GPU ``KeyframeDB``.

    ``rank`` / ``world`` default to the process group's; passing them explicitly gives the rank-local half of the
    query -- ``local_keys(q)`` -- for tests that merge the ranks by hand on one GPU.
    """

    def __init__(self, device: int = 0, capacity: int = 1 << 16, group=None, ratio=REFERENCE_RATIO, rank=None, world=None):
        import torch
        import torch.distributed as dist
        have_pg = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = int(world) if world is not None else (dist.get_world_size(group) if have_pg else 1)
        self.rank = int(rank) if rank is not None else (dist.get_rank(group) if have_pg else 0)
        self.ratio = ratio
        self.device = torch.device("cuda", device)
        self._rows = torch.empty((capacity, 32), dtype=torch.uint8, device=self.device)
        self._n_local = 0
        self._offsets = [0]              # first global row of every keyframe, + total (same on all ranks)
        self._seg_local = [0]            # this rank's segments: first local row ...
        self._seg_global = []            # ... and first global row of each
        self._seg_dev = None

    def __len__(self) -> int:
        return len(self._offsets) - 1

    @property
    def n_rows(self) -> int:
        return self._offsets[-1]

    @property
    def n_local_rows(self) -> int:
        return self._n_local

    def add(self, descriptors) -> int:
        import torch
        kf = len(self._offsets) - 1
        n = int(descriptors.shape[0])
        first = self._offsets[-1]
        self._offsets.append(first + n)
        if kf % self.world == self.rank and n > 0:
            d = descriptors if hasattr(descriptors, "is_cuda") else torch.from_numpy(_as_desc(descriptors, "descriptors"))
            if self._n_local + n > self._rows.shape[0]:
                cap = max(2 * self._rows.shape[0], self._n_local + n)
                grown = torch.empty((cap, 32), dtype=torch.uint8, device=self.device)
                grown[: self._n_local] = self._rows[: self._n_local]
                self._rows = grown
            self._rows[self._n_local:self._n_local + n].copy_(d, non_blocking=True)
            self._seg_global.append(first)
            self._n_local += n
            self._seg_local.append(self._n_local)
            self._seg_dev = None
        return kf

    def locate(self, global_idx):
        g = np.asarray(global_idx)
        off = np.asarray(self._offsets)
        kf = np.searchsorted(off, np.maximum(g, 0), side="right") - 1
        return np.where(g < 0, -1, kf), np.where(g < 0, -1, g - off[kf])

    def local_keys(self, qd):
        """This rank's packed top-2 keys ``int64[nq, 2]`` with GLOBAL row indices."""
        import torch
        from . import _lib
        nq = int(qd.shape[0])
        keys = torch.empty((nq, 2), dtype=torch.int64, device=self.device)
        ctx = _lib.context(self.device.index or 0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(ctx.lib.slm_knn2_keys(ctx.handle, qd.data_ptr(), nq, self._rows.data_ptr() if self._n_local else None,
                                         self._n_local, 0, keys.data_ptr(), stream))
        if self._n_local == 0:
            return keys
        if self._seg_dev is None:
            self._seg_dev = (torch.tensor(self._seg_local[:-1], dtype=torch.int64, device=self.device),
                             torch.tensor(self._seg_global, dtype=torch.int64, device=self.device))
        seg_local, seg_global = self._seg_dev
        # rebase local rows to global rows, segment by segment (missing neighbours stay SLM_KEY_NONE = -1)
        none = keys == -1
        row = keys & 0xFFFFFFFF
        seg = torch.searchsorted(seg_local, row, right=True) - 1
        glob = row - seg_local[seg] + seg_global[seg]
        return torch.where(none, keys, (keys & ~0xFFFFFFFF) | glob)

    def query(self, q):
        """kNN-2 of ``q`` against the whole distributed collection; every rank returns the same full result
        (numpy for numpy queries, CUDA tensors otherwise)."""
        import torch
        import torch.distributed as dist
        from . import _lib
        host = not hasattr(q, "is_cuda")
        qd = torch.from_numpy(_as_desc(q, "queryDescriptors")).to(self.device, non_blocking=True) if host else q
        nq = int(qd.shape[0])
        keys = self.local_keys(qd)
        gathered = torch.empty((self.world, nq, 2), dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_gather_into_tensor(gathered.view(self.world * nq, 2), keys.contiguous(), group=self.group)
        else:
            gathered[0] = keys
        idx = torch.empty((nq, 2), dtype=torch.int32, device=self.device)
        dist_ = torch.empty((nq, 2), dtype=torch.int32, device=self.device)
        acc = torch.empty((nq,), dtype=torch.uint8, device=self.device)
        num, den = self.ratio if self.ratio is not None else (0, 1)
        ctx = _lib.context(self.device.index or 0)
        _lib.check(ctx.lib.slm_merge_top2(ctx.handle, gathered.data_ptr(), self.world, nq, int(num), int(den), idx.data_ptr(),
                                          dist_.data_ptr(), acc.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        if host:
            return idx.cpu().numpy(), dist_.cpu().numpy(), acc.cpu().numpy()
        return idx, dist_, acc
