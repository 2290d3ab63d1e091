"""Train-set sharding over the GPUs of one box (BASELINE configs 4 and 5).

The keyframe-descriptor database (config 5) or the visual vocabulary (config 4) is split into contiguous
row blocks, one per rank; the query batch is replicated.  Every rank runs the single-GPU kernel on its
block with ``train_index_base`` = first global row and emits packed top-2 keys
``(distance << 32) | global index``; ONE all-gather of ``nq x 2`` keys per rank (NCCL over NVLink) is the
only exchange step, followed by a merge kernel that keeps the two smallest keys -- unsigned key order is
OpenCV's ``(distance, imgIdx, trainIdx)`` collection order, so the result is byte-identical for any
shard count.  Frame-to-frame / local-map matching (configs 1-3) never shards.

The reference has no counterpart (loop_closure.py:7-36 matches two frames; place_recognition.py is
empty): this module only defines how the existing kernel scales.
"""
from __future__ import annotations

from typing import Callable, Optional


def shard_bounds(n_rows: int, world_size: int):
    """Contiguous, near-equal row blocks: [(first, last_exclusive)] per rank (32-row aligned interior cuts)."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    cuts = [0]
    for r in range(1, world_size):
        c = (n_rows * r) // world_size
        c -= c % 32
        cuts.append(max(c, cuts[-1]))
    cuts.append(n_rows)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


class ShardedMatcher:
    """One rank's view of a sharded train set.

    ``local_keys(q) -> int64[nq, 2]`` and ``merge(gathered int64[W, nq, 2]) -> result`` default to the CUDA
    entry points (slm_knn2_keys / slm_merge_top2); tests inject CPU stand-ins to exercise the collective
    plumbing under gloo.
    """

    def __init__(self, train_shard, first_row: int, group=None, ratio=(7, 10), variant: Optional[str] = None,
                 local_keys: Optional[Callable] = None, merge: Optional[Callable] = None, exchange: str = "auto"):
        import torch.distributed as dist
        self.train = train_shard
        self.first_row = int(first_row)
        self.group = group
        self.ratio = ratio
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._local_keys = local_keys or self._cuda_local_keys
        self._merge = merge or self._cuda_merge
        self._variant = variant
        self._gather_bufs = {}
        # exchange of the per-rank keys: "nccl" = all-gather + merge kernel; "nvlink" = one kernel that stores
        # the keys straight into every peer's symmetric-memory buffer, flags them and merges (slm_exchange_merge);
        # "auto" tries nvlink for small query batches and falls back to nccl when peer mapping is unavailable
        # "a2a" = all-to-all of query slices + all-gather of the merged results: rank r merges only queries
        # [r * nq / W, (r + 1) * nq / W), so every rank receives ~2 x 16 B x nq instead of 16 B x nq x (W - 1);
        # (config 4: 1M descriptors -> 28 MB instead of 112 MB per rank at W = 8); opt-in, "auto" never picks it
        if exchange not in ("auto", "nccl", "nvlink", "a2a"):
            raise ValueError("exchange must be auto, nccl, nvlink or a2a")
        self.exchange = exchange if (exchange == "a2a" or (local_keys is None and merge is None and self.world > 1)) else "nccl"
        self._symm = None
        self._step = 0
        self.last_exchange = "none"     # what the last knn2() call actually used

    # -- CUDA implementations ---------------------------------------------------------------------
    def _ctx(self):
        from . import _lib
        ctx = _lib.context(self.train.device.index or 0)
        if self._variant is not None:
            ctx.set_variant(self._variant)
        return ctx

    def _cuda_local_keys(self, q, out=None):
        import torch
        from . import _lib
        ctx = self._ctx()
        nq, nt = q.shape[0], self.train.shape[0]
        keys = out if out is not None else torch.empty((nq, 2), dtype=torch.int64, device=q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(ctx.lib.slm_knn2_keys(ctx.handle, q.data_ptr(), nq, self.train.data_ptr() if nt else None, nt,
                                         self.first_row, keys.data_ptr(), stream))
        return keys

    def _cuda_merge(self, gathered):
        import torch
        from . import _lib
        ctx = self._ctx()
        w, nq, _ = gathered.shape
        dev = gathered.device
        idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        dist_ = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        acc = torch.empty((nq,), dtype=torch.uint8, device=dev)
        num, den = self.ratio if self.ratio is not None else (0, 1)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(ctx.lib.slm_merge_top2(ctx.handle, gathered.data_ptr(), w, nq, int(num), int(den), idx.data_ptr(),
                                          dist_.data_ptr(), acc.data_ptr(), stream))
        return idx, dist_, acc

    # -- NVLink exchange ----------------------------------------------------------------------------
    _NVLINK_MAX_NQ = 8192

    def _symm_setup(self, device):
        """Peer-mapped key buffers + flags through torch symmetric memory (NVLink / NVSwitch P2P)."""
        import ctypes
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        cap = self._NVLINK_MAX_NQ
        keys = symm_mem.empty((2, self.world, cap, 2), dtype=torch.int64, device=device)
        flags = symm_mem.empty((2 * self.world,), dtype=torch.int32, device=device)
        flags.zero_()
        hk = symm_mem.rendezvous(keys, group)
        hf = symm_mem.rendezvous(flags, group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                       # every rank's flags are zero before anybody publishes
        arr = ctypes.c_uint64 * self.world
        self._symm = dict(keys=keys, flags=flags, hk=hk, hf=hf, cap=cap, rank=dist.get_rank(group),
                          key_ptrs=arr(*[int(x) for x in hk.buffer_ptrs]),
                          flag_ptrs=arr(*[int(x) for x in hf.buffer_ptrs]))

    def _nvlink_ready(self, q) -> bool:
        if self.exchange in ("nccl", "a2a") or not (0 < q.shape[0] <= self._NVLINK_MAX_NQ) or not getattr(q, "is_cuda", False):
            return False
        if self._symm is None:
            try:
                self._symm_setup(q.device)
            except Exception as e:              # no peer mapping in this environment
                if self.exchange == "nvlink":
                    raise
                import warnings
                warnings.warn(f"slammatch: NVLink exchange unavailable ({e!r}); using the NCCL all-gather")
                self.exchange = "nccl"
                return False
        return True

    def _knn2_nvlink(self, q):
        import ctypes
        import torch
        from . import _lib
        s = self._symm
        ctx = self._ctx()
        nq, dev = q.shape[0], q.device
        idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        dist_ = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        acc = torch.empty((nq,), dtype=torch.uint8, device=dev)
        num, den = self.ratio if self.ratio is not None else (0, 1)
        self._step += 1
        nt = self.train.shape[0]
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(ctx.lib.slm_knn2_exchange(ctx.handle, q.data_ptr(), nq, self.train.data_ptr() if nt else None, nt,
                                             self.first_row, s["cap"], ctypes.cast(s["key_ptrs"], ctypes.c_void_p),
                                             ctypes.cast(s["flag_ptrs"], ctypes.c_void_p), s["rank"], self.world,
                                             self._step, int(num), int(den), idx.data_ptr(), dist_.data_ptr(),
                                             acc.data_ptr(), stream))
        return idx, dist_, acc

    # -- the sharded query ----------------------------------------------------------------------------
    def _gather(self, keys, slot, async_op=False):
        import torch
        import torch.distributed as dist
        shape = (self.world,) + tuple(keys.shape)
        buf = self._gather_bufs.get(slot)
        if buf is None or tuple(buf.shape) != shape or buf.device != keys.device:
            buf = torch.empty(shape, dtype=keys.dtype, device=keys.device)
            self._gather_bufs[slot] = buf
        # output laid out as the concatenation along dim 0 (the form every backend accepts)
        work = dist.all_gather_into_tensor(buf.view((self.world * keys.shape[0],) + tuple(keys.shape[1:])),
                                           keys.contiguous(), group=self.group, async_op=async_op)
        return buf, work

    def _knn2_a2a(self, q):
        """Exchange for large query sets.  The merge is associative, so nobody needs everybody's keys for every query:
        an all-to-all hands rank r all ranks' keys of ITS query slice, r merges + finalises the slice, and three small
        all-gathers (idx, dist, accept) replicate the result."""
        import torch
        import torch.distributed as dist
        W, nq, dev = self.world, q.shape[0], q.device
        per = -(-nq // W)
        pad = W * per
        send = self._gather_bufs.get("a2a_send")
        if send is None or send.shape[0] != pad or send.device != dev:
            send = torch.empty((pad, 2), dtype=torch.int64, device=dev)
            recv = torch.empty((W, per, 2), dtype=torch.int64, device=dev)
            self._gather_bufs["a2a_send"], self._gather_bufs["a2a_recv"] = send, recv
        recv = self._gather_bufs["a2a_recv"]
        if pad > nq:
            send[nq:] = -1                      # = SLM_KEY_NONE: padding queries have no neighbours anywhere
        if self._local_keys == self._cuda_local_keys:
            self._local_keys(q, out=send[:nq])
        else:
            send[:nq] = self._local_keys(q)
        dist.all_to_all_single(recv.view(pad, 2), send, group=self.group)
        idx_s, dist_s, acc_s = self._merge(recv)
        idx = torch.empty((pad, 2), dtype=idx_s.dtype, device=dev)
        dist_ = torch.empty((pad, 2), dtype=dist_s.dtype, device=dev)
        acc = torch.empty((pad,), dtype=acc_s.dtype, device=dev)
        dist.all_gather_into_tensor(idx, idx_s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(dist_, dist_s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(acc, acc_s.contiguous(), group=self.group)
        return idx[:nq], dist_[:nq], acc[:nq]

    def knn2(self, q, query_batch: int = 1 << 22):
        """Local top-2 keys -> all-gather -> merge.  Every rank returns the full, identical result.

        Query sets above ``query_batch`` rows are cut into batches so that the all-gather of batch b (16 bytes
        per query and rank) overlaps the search of batch b+1.  (Measured on 8 B200s for config 4 the extra
        launches cost more than the overlap saves, hence the high default.)"""
        import torch
        keys_fn, nq = self._local_keys, q.shape[0]
        if self.world == 1:
            keys = keys_fn(q)
            return self._merge(keys.reshape((1,) + tuple(keys.shape)))
        if self._nvlink_ready(q):
            self.last_exchange = "nvlink peer stores + flags (slm_exchange_merge)"
            return self._knn2_nvlink(q)
        if self.exchange == "a2a":      # opt-in until measured at 8 GPUs (neutral at 2: 4.71 vs 4.68 ms on c4)
            self.last_exchange = "nccl all-to-all of query slices + all-gather of merged results"
            return self._knn2_a2a(q)
        self.last_exchange = "nccl all-gather"
        if nq <= query_batch:
            if keys_fn == self._cuda_local_keys:
                # in-place all-gather: the search writes this rank's keys straight into its slot of the buffer
                import torch.distributed as dist
                shape = (self.world, nq, 2)
                buf = self._gather_bufs.get(0)
                if buf is None or tuple(buf.shape) != shape or buf.device != q.device:
                    buf = torch.empty(shape, dtype=torch.int64, device=q.device)
                    self._gather_bufs[0] = buf
                rank = dist.get_rank(self.group)
                keys_fn(q, out=buf[rank])
                dist.all_gather_into_tensor(buf.view(self.world * nq, 2), buf[rank], group=self.group)
                return self._merge(buf)
            buf, _ = self._gather(keys_fn(q), 0)
            return self._merge(buf)
        outs, pending = [], None
        for b, s in enumerate(range(0, nq, query_batch)):
            keys = keys_fn(q[s:s + query_batch])
            buf, work = self._gather(keys, b & 1, async_op=True)
            if pending is not None:
                pending[1].wait()
                outs.append(self._merge(pending[0]))
            pending = (buf, work, keys)          # keep `keys` alive until its gather has been consumed
        pending[1].wait()
        outs.append(self._merge(pending[0]))
        return tuple(torch.cat([o[i] for o in outs], dim=0) for i in range(3))
