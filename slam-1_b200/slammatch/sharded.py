"""Train-set sharding over the GPUs of one box (BASELINE configs 4 and 5).

The keyframe-descriptor database (config 5) or the visual vocabulary (config 4) is split into contiguous
row blocks, one per rank; the query batch is replicated.  Every rank runs the single-GPU kernel on its
block with ``train_index_base`` = first global row and emits packed top-2 keys
``(distance << 32) | global index``; ONE exchange of ``nq x 2`` keys per rank is the only communication step,
followed by a merge that keeps the two smallest keys -- unsigned key order is OpenCV's
``(distance, imgIdx, trainIdx)`` collection order, so the result is byte-identical for any shard count.
Frame-to-frame / local-map matching (configs 1-3) never shards.

Exchanges (``exchange=``):
  ``nvlink``  peer stores over NVLink / NVSwitch into torch symmetric memory + flags + merge kernel
              (slm_knn2_exchange: the tensor path's refine kernel is the producer).  Keys travel as 32-bit words
              ``distance << 16 | index`` when the whole train set has at most 65 536 rows (config 4's vocabulary).
              From 32 768 queries on, the ranks first exchange candidate-chunk keys and only the owner of a query's
              global best two chunks re-scores them exactly -- the re-scoring is shared, not repeated on every rank.
  ``nccl``    all-gather of the packed keys + slm_merge_top2 (the form north_star names; also the CPU / gloo form).
  ``a2a``     all-to-all of query slices + per-rank merge + all-gather of the merged results.
  ``auto``    nvlink when every rank can set it up (the ranks agree through an all-reduce), else nccl.

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


_MAX_NVLINK_WORLD = 16      # kSlmMaxWorld in csrc/slm_internal.cuh


class ShardedMatcher:
    """One rank's view of a sharded train set.

    ``local_keys(q) -> int64[nq, 2]`` and ``merge(gathered int64[W, nq, 2]) -> result`` default to the CUDA
    entry points (slm_knn2_keys / slm_merge_top2); tests inject CPU stand-ins to exercise the collective
    plumbing under gloo.  ``total_rows`` = number of train rows over all ranks (an all-reduce at the first query
    when omitted).
    """

    def __init__(self, train_shard, first_row: int, group=None, ratio=(7, 10), variant: Optional[str] = None,
                 local_keys: Optional[Callable] = None, merge: Optional[Callable] = None, exchange: str = "auto",
                 total_rows: Optional[int] = None):
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
        if exchange not in ("auto", "nccl", "nvlink", "a2a"):
            raise ValueError("exchange must be auto, nccl, nvlink or a2a")
        cuda_path = local_keys is None and merge is None and self.world > 1
        self.exchange = exchange if (exchange == "a2a" or cuda_path) else "nccl"
        self.total_rows = None if total_rows is None else int(total_rows)
        self._symm = None
        self._step = 0
        self.last_exchange = "none"     # what the last knn2() call actually used

    # -- CUDA implementations ---------------------------------------------------------------------
    def _ctx(self):
        from . import _lib
        return _lib.context(self.train.device.index or 0)

    def _cuda_local_keys(self, q, out=None):
        import torch
        from . import _lib
        ctx = self._ctx()
        nq, nt = q.shape[0], self.train.shape[0]
        keys = out if out is not None else torch.empty((nq, 2), dtype=torch.int64, device=q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        with ctx.using(self._variant):
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

    def _total_rows(self, device):
        """Train rows over all ranks (selects the exchange's key width; identical on every rank)."""
        if self.total_rows is None:
            import torch
            import torch.distributed as dist
            n = torch.tensor([int(self.train.shape[0])], dtype=torch.int64, device=device)
            if self.world > 1:
                dist.all_reduce(n, group=self.group)
            self.total_rows = int(n.item())
        return self.total_rows

    # -- NVLink exchange ----------------------------------------------------------------------------
    def _symm_setup(self, device, cap):
        """Peer-mapped key buffers + flags through torch symmetric memory (NVLink / NVSwitch P2P)."""
        import ctypes
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        # [2 phases][2 halves][world][cap][2] keys, uint32[2][2][world] flags (include/slammatch.h)
        keys = symm_mem.empty((4, self.world, cap, 2), dtype=torch.int64, device=device)
        flags = symm_mem.empty((4 * self.world,), dtype=torch.int32, device=device)
        flags.zero_()
        hk = symm_mem.rendezvous(keys, group)
        hf = symm_mem.rendezvous(flags, group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                       # every rank's flags are zero before anybody publishes
        arr = ctypes.c_uint64 * self.world
        self._symm = dict(keys=keys, flags=flags, hk=hk, hf=hf, cap=cap, rank=dist.get_rank(group),
                          key_ptrs=arr(*[int(x) for x in hk.buffer_ptrs]),
                          flag_ptrs=arr(*[int(x) for x in hf.buffer_ptrs]))
        self._step = 0                                  # fresh flags: the step counter restarts with them

    def _nvlink_ready(self, q) -> bool:
        if self.exchange in ("nccl", "a2a") or q.shape[0] == 0 or not getattr(q, "is_cuda", False):
            return False
        if self.world > _MAX_NVLINK_WORLD:
            if self.exchange == "nvlink":
                raise ValueError(f"the NVLink exchange supports at most {_MAX_NVLINK_WORLD} ranks")
            self.exchange = "nccl"
            return False
        nq = int(q.shape[0])
        if self._symm is None or self._symm["cap"] < nq:
            # (re)allocate for the larger batch -- nq is the same on every rank, so every rank gets here together.
            # The ranks then AGREE on the outcome: one that cannot map its peers would otherwise enter the all-gather
            # while the others wait on flags.
            import torch
            import torch.distributed as dist
            cap = max(2048, 1 << (nq - 1).bit_length())
            err = None
            try:
                if self._symm is not None:
                    torch.cuda.synchronize(q.device)    # nobody is still reading the old buffers
                    dist.barrier(group=self.group)
                self._symm_setup(q.device, cap)
            except Exception as e:                      # no peer mapping in this environment
                err = e
                self._symm = None
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=q.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._symm = None
                if self.exchange == "nvlink":
                    raise RuntimeError(f"NVLink exchange unavailable on at least one rank ({err!r})")
                import warnings
                warnings.warn(f"slammatch: NVLink exchange unavailable ({err!r}); every rank uses the NCCL all-gather")
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
        total = self._total_rows(dev)
        self._step += 1
        nt = self.train.shape[0]
        stream = torch.cuda.current_stream(dev).cuda_stream
        with ctx.using(self._variant):
            _lib.check(ctx.lib.slm_knn2_exchange(ctx.handle, q.data_ptr(), nq, self.train.data_ptr() if nt else None, nt,
                                                 self.first_row, s["cap"], total,
                                                 ctypes.cast(s["key_ptrs"], ctypes.c_void_p),
                                                 ctypes.cast(s["flag_ptrs"], ctypes.c_void_p), s["rank"], self.world,
                                                 self._step, int(num), int(den), idx.data_ptr(), dist_.data_ptr(),
                                                 acc.data_ptr(), stream))
        return idx, dist_, acc

    def check(self):
        """Raise if an earlier NVLink exchange on this rank lost a peer (call after synchronising the stream)."""
        from . import _lib
        ctx = self._ctx()
        _lib.check(ctx.lib.slm_exchange_status(ctx.handle))

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
        """Local top-2 keys -> exchange -> merge.  Every rank returns the full, identical result.

        NCCL form: query sets above ``query_batch`` rows are cut into batches so that the all-gather of batch b (16
        bytes per query and rank) overlaps the search of batch b+1."""
        import torch
        keys_fn, nq = self._local_keys, q.shape[0]
        if self.world == 1:
            keys = keys_fn(q)
            return self._merge(keys.reshape((1,) + tuple(keys.shape)))
        if self._nvlink_ready(q):
            self.last_exchange = "nvlink peer stores + flags (slm_knn2_exchange)"
            return self._knn2_nvlink(q)
        if self.exchange == "a2a":
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

    def knn2_host(self, q_host, train_host=None):
        """End-to-end host form of the sharded query: pinned or pageable numpy ``uint8[nq, 32]`` queries in, numpy results
        out.  ``train_host`` (this rank's rows, optional) re-uploads the shard first -- the form bench.py times as e2e."""
        import numpy as np
        import torch
        dev = self.train.device
        if train_host is not None:
            self.train.copy_(torch.from_numpy(train_host) if isinstance(train_host, np.ndarray) else train_host, non_blocking=True)
        qd = (torch.from_numpy(q_host) if isinstance(q_host, np.ndarray) else q_host).to(dev, non_blocking=True)
        idx, dist_, acc = self.knn2(qd)
        nq = idx.shape[0]
        # one packed read-back: [idx | dist | accept]
        packed = torch.empty((nq, 17), dtype=torch.uint8, device=dev)
        packed[:, 0:8] = idx.view(torch.uint8).view(nq, 8)
        packed[:, 8:16] = dist_.view(torch.uint8).view(nq, 8)
        packed[:, 16] = acc
        h = packed.cpu().numpy()
        return (h[:, 0:8].copy().view(np.int32).reshape(nq, 2), h[:, 8:16].copy().view(np.int32).reshape(nq, 2),
                h[:, 16].copy())


class QueryShardedMatcher:
    """The other decomposition of a multi-GPU search (note): the train set is REPLICATED and the
    QUERIES are split into contiguous slices, one per rank.

    It fits a small train set searched by very many queries -- BASELINE config 4: one million descriptors against a 2 MB
    vocabulary -- where the train-sharded form repeats the per-query work (candidate re-scoring, key exchange, merge) on
    every rank.  Here a rank runs the plain single-GPU search (slm_knn2_filter) on its slice, writing straight into its
    part of the result arrays, and the only communication is the in-place all-gather of the finished results (17 bytes per
    query): no candidate exchange, no merge, and the result is trivially identical to the single-GPU one.  It does not
    apply when the train set is what is too large for one GPU (config 5) -- that is what ShardedMatcher is for.

    ``knn2_slice(q_slice, out_idx, out_dist, out_acc)`` defaults to the CUDA entry point; tests inject a CPU stand-in to
    exercise the plumbing under gloo.
    """

    def __init__(self, train, group=None, ratio=(7, 10), cross_check: bool = False, variant: Optional[str] = None,
                 knn2_slice: Optional[Callable] = None):
        import torch.distributed as dist
        self.train = train
        self.group = group
        self.ratio = ratio
        self.cross_check = bool(cross_check)
        self._variant = variant
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._knn2_slice = knn2_slice or self._cuda_knn2_slice
        self._bufs = None
        self.last_exchange = "none"

    def _cuda_knn2_slice(self, q, idx, dist_, acc):
        import torch
        from . import _lib
        ctx = _lib.context(self.train.device.index or 0)
        nq, nt = q.shape[0], self.train.shape[0]
        if nq == 0:
            return
        num, den = self.ratio if self.ratio is not None else (0, 1)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        with ctx.using(self._variant):
            _lib.check(ctx.lib.slm_knn2_filter(ctx.handle, q.data_ptr(), nq, self.train.data_ptr() if nt else None, nt, 0,
                                               int(num), int(den), int(self.cross_check), idx.data_ptr(), dist_.data_ptr(),
                                               acc.data_ptr(), stream))

    def slice_bounds(self, nq: int):
        """(rows per rank, first, last_exclusive) of this rank's query slice: equal slices, the last ones ragged or empty."""
        per = -(-nq // self.world) if nq else 0
        lo = min(nq, self.rank * per)
        return per, lo, min(nq, lo + per)

    def knn2(self, q):
        """Every rank passes the SAME full query set and returns the full, identical result."""
        import torch
        import torch.distributed as dist
        nq, dev = q.shape[0], q.device
        per, lo, hi = self.slice_bounds(nq)
        pad = per * self.world
        b = self._bufs
        if b is None or b[0].shape[0] != pad or b[0].device != dev:
            b = self._bufs = (torch.empty((pad, 2), dtype=torch.int32, device=dev),
                              torch.empty((pad, 2), dtype=torch.int32, device=dev),
                              torch.empty((pad,), dtype=torch.uint8, device=dev))
        idx, dist_, acc = b
        if self.cross_check and self.world > 1:
            raise ValueError("cross-check needs every query on one rank: the reverse search ranks ALL queries of a train row")
        self._knn2_slice(q[lo:hi], idx[lo:hi], dist_[lo:hi], acc[lo:hi])
        if self.world > 1:
            # in place: this rank's slice already sits at its offset of the gathered arrays
            r0 = self.rank * per
            for x in (idx, dist_, acc):
                dist.all_gather_into_tensor(x, x[r0:r0 + per], group=self.group)
            self.last_exchange = "in-place all-gather of the finished results (no candidate exchange)"
        return idx[:nq], dist_[:nq], acc[:nq]

    def knn2_host(self, q_host, train_host=None):
        """Host form: this rank uploads ITS query slice only (and, optionally, the replicated train set again); the gathered
        result is read back in one packed copy."""
        import numpy as np
        import torch
        dev = self.train.device
        if train_host is not None:
            self.train.copy_(torch.from_numpy(train_host) if isinstance(train_host, np.ndarray) else train_host, non_blocking=True)
        qh = torch.from_numpy(q_host) if isinstance(q_host, np.ndarray) else q_host
        nq = qh.shape[0]
        per, lo, hi = self.slice_bounds(nq)
        qd = self._host_q
        if qd is None or qd.shape[0] != nq or qd.device != dev:
            qd = self._host_q = torch.empty((nq,) + tuple(qh.shape[1:]), dtype=qh.dtype, device=dev)
        qd[lo:hi].copy_(qh[lo:hi], non_blocking=True)
        idx, dist_, acc = self.knn2(qd)
        packed = torch.empty((nq, 17), dtype=torch.uint8, device=dev)
        packed[:, 0:8] = idx.view(torch.uint8).view(nq, 8)
        packed[:, 8:16] = dist_.view(torch.uint8).view(nq, 8)
        packed[:, 16] = acc
        h = packed.cpu().numpy()
        return (h[:, 0:8].copy().view(np.int32).reshape(nq, 2), h[:, 8:16].copy().view(np.int32).reshape(nq, 2),
                h[:, 16].copy())

    _host_q = None
