"""Single-GPU loopback of the sharded path's NVLink exchange (slm_knn2_exchange / slm_exchange_merge).

The kernels SCALE_rNN times on 2-8 GPUs -- the refine kernel that stores every query's keys into all peers' buffers and
publishes the flags, and exchange_wait_merge_kernel that acquires them and merges -- are driven here as W "ranks" on W
streams of ONE GPU: W contexts, plain CUDA tensors as the peer-mapped key buffers and flag arrays (every "peer
pointer" is just another tensor on the same device).  Each step runs all ranks' calls back to back, the ranks wait for
each other only through the flags, and every rank's result must equal the oracle on the full train set, bit for bit,
for several steps (both buffer halves are reused) and for both key widths.
"""
import ctypes
import os

import numpy as np
import pytest

import slammatch
from slammatch import _lib, synth
from slammatch.sharded import shard_bounds
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


class _Ranks:
    """W loopback ranks on one GPU."""

    def __init__(self, world, cap):
        import torch
        # few polling blocks per rank: the other ranks' search kernels need whole SMs of the same GPU to make progress;
        # the two-phase form from 2 ranks on (the default, 4, is a measured crossover, not a correctness limit)
        env = {"SLM_EXCHANGE_MAX_BLOCKS": "8", "SLM_EXCHANGE_TWO_PHASE_WORLD": "2"}
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            self.ctxs = [_lib.Context(0) for _ in range(world)]
        finally:
            for k, v in old.items():
                if v is None:
                    del os.environ[k]
                else:
                    os.environ[k] = v
        self.world, self.cap = world, cap
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        # [2 phases][2 halves][world][cap][2] keys and uint32[2][2][world] flags per rank (include/slammatch.h)
        self.keys = [torch.zeros((4, world, cap, 2), dtype=torch.int64, device="cuda") for _ in range(world)]
        self.flags = [torch.zeros((4 * world,), dtype=torch.int32, device="cuda") for _ in range(world)]
        arr = ctypes.c_uint64 * world
        self.key_ptrs = arr(*[k.data_ptr() for k in self.keys])
        self.flag_ptrs = arr(*[f.data_ptr() for f in self.flags])
        self.step = 0
        torch.cuda.synchronize()

    def close(self):
        for c in self.ctxs:
            c.close()

    def _warm(self, q_dev, shards, bounds, variant, fused):
        """Load every kernel the concurrent step will launch, one rank at a time.  CUDA loads kernels lazily at their first
        launch and a load may wait for running kernels to finish -- on ONE GPU that would deadlock against another
        'rank' whose merge kernel is polling for this rank's flags.  (No issue across real GPUs: one process per device.)"""
        import torch
        nq = q_dev.shape[0]
        keys = torch.zeros((4, 1, self.cap, 2), dtype=torch.int64, device="cuda")
        flags = torch.zeros((4,), dtype=torch.int32, device="cuda")
        kp, fp = (ctypes.c_uint64 * 1)(keys.data_ptr()), (ctypes.c_uint64 * 1)(flags.data_ptr())
        idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
        ctx = self.ctxs[0]
        for r in range(self.world):
            t, (a, b) = shards[r], bounds[r]
            with ctx.using(variant):
                if fused:
                    _lib.check(ctx.lib.slm_knn2_exchange(ctx.handle, q_dev.data_ptr(), nq, t.data_ptr() if b > a else None, b - a,
                                                         a, self.cap, 0, ctypes.cast(kp, ctypes.c_void_p),
                                                         ctypes.cast(fp, ctypes.c_void_p), 0, 1, r + 1, 7, 10, idx.data_ptr(),
                                                         None, None, None))
                else:
                    k = torch.empty((nq, 2), dtype=torch.int64, device="cuda")
                    _lib.check(ctx.lib.slm_knn2_keys(ctx.handle, q_dev.data_ptr(), nq, t.data_ptr() if b > a else None, b - a, a,
                                                     k.data_ptr(), None))
                    _lib.check(ctx.lib.slm_exchange_merge(ctx.handle, k.data_ptr(), nq, self.cap, 0, ctypes.cast(kp, ctypes.c_void_p),
                                                          ctypes.cast(fp, ctypes.c_void_p), 0, 1, r + 1, 7, 10, idx.data_ptr(),
                                                          None, None, None))
            torch.cuda.synchronize()

    def query(self, q_dev, shards, bounds, total_rows, ratio, variant, fused=True):
        """One sharded step on all ranks; returns every rank's (idx, dist, acc) as numpy."""
        import torch
        nq = q_dev.shape[0]
        self._warm(q_dev, shards, bounds, variant, fused)
        self.step += 1
        outs = []
        for r in range(self.world):
            ctx = self.ctxs[r]
            idx = torch.full((nq, 2), -9, dtype=torch.int32, device="cuda")
            dist = torch.full((nq, 2), -9, dtype=torch.int32, device="cuda")
            acc = torch.full((nq,), 9, dtype=torch.uint8, device="cuda")
            outs.append((idx, dist, acc))
            t, (a, b) = shards[r], bounds[r]
            s = self.streams[r]
            s.wait_stream(torch.cuda.current_stream())
            kp = ctypes.cast(self.key_ptrs, ctypes.c_void_p)
            fp = ctypes.cast(self.flag_ptrs, ctypes.c_void_p)
            with ctx.using(variant):
                if fused:
                    _lib.check(ctx.lib.slm_knn2_exchange(ctx.handle, q_dev.data_ptr(), nq, t.data_ptr() if b > a else None,
                                                         b - a, a, self.cap, total_rows, kp, fp, r, self.world, self.step,
                                                         ratio[0], ratio[1], idx.data_ptr(), dist.data_ptr(), acc.data_ptr(),
                                                         s.cuda_stream))
                else:
                    keys = torch.empty((nq, 2), dtype=torch.int64, device="cuda")
                    _lib.check(ctx.lib.slm_knn2_keys(ctx.handle, q_dev.data_ptr(), nq, t.data_ptr() if b > a else None, b - a,
                                                     a, keys.data_ptr(), s.cuda_stream))
                    _lib.check(ctx.lib.slm_exchange_merge(ctx.handle, keys.data_ptr(), nq, self.cap, total_rows, kp, fp, r,
                                                          self.world, self.step, ratio[0], ratio[1], idx.data_ptr(),
                                                          dist.data_ptr(), acc.data_ptr(), s.cuda_stream))
                    outs[-1] = outs[-1] + (keys,)
        torch.cuda.synchronize()
        for ctx in self.ctxs:
            _lib.check(ctx.lib.slm_exchange_status(ctx.handle))
        return [(o[0].cpu().numpy(), o[1].cpu().numpy(), o[2].cpu().numpy()) for o in outs]


def _case(nq, nt, seed):
    q, t = synth.planted(nq, nt, seed)
    return q, synth.with_duplicates(t, seed + 1, 0.2)        # exact duplicates across shard boundaries


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("variant", ["auto", "tensor", "popc"])
def test_loopback_exchange_equals_oracle_64bit_keys(world, variant):
    """Config 5 in miniature: few queries, long train set (global indices above 65 536 -> 64-bit keys)."""
    import torch
    nq, nt = 600, 150_000
    ranks = _Ranks(world, cap=1024)
    try:
        for step in range(5):
            q, t = _case(nq if step != 3 else 37, nt, 900 + 10 * step)      # a ragged batch in the middle
            bounds = shard_bounds(nt, world)
            shards = [torch.from_numpy(t[a:b]).cuda() for a, b in bounds]
            oi, od = orc.c_knn2(q, t)
            want = orc.c_ratio(od, 7, 10)
            res = ranks.query(torch.from_numpy(q).cuda(), shards, bounds, nt, (7, 10), variant)
            for r, (i, d, a) in enumerate(res):
                assert np.array_equal(i, oi) and np.array_equal(d, od), (world, variant, step, r)
                assert np.array_equal(a, want), (world, variant, step, r)
    finally:
        ranks.close()


@pytest.mark.parametrize("world,nq", [(2, 20_000), (8, 20_000), (2, 70_000), (4, 40_000)])
@pytest.mark.parametrize("variant", ["auto", "tensor"])
def test_loopback_exchange_config4_regime_32bit_keys(world, nq, variant):
    """Config 4 in miniature: many queries (above the old 8192 cap), a vocabulary of at most 65 536 words sharded by rows
    -> compact 32-bit keys (distance << 16 | word), merge kernel with a capped grid.  From 32 768 queries on the tensor path
    takes the two-phase form config 4 itself runs with: candidate-chunk keys are exchanged first and only the owner of a
    query's global best two chunks re-scores them (tc_chunk_keys_kernel / tc_refine_owned_kernel); 70 000 queries also
    reach the multi-group blocks (8 x 32 queries per block, one fence per block)."""
    import torch
    nt = 48_000 if nq <= 20_000 else 30_000
    ranks = _Ranks(world, cap=1 << (nq - 1).bit_length())
    try:
        for step in range(3):
            q, t = _case(nq, nt, 1900 + 10 * step)
            bounds = shard_bounds(nt, world)
            shards = [torch.from_numpy(t[a:b]).cuda() for a, b in bounds]
            oi, od = orc.c_knn2(q, t)
            res = ranks.query(torch.from_numpy(q).cuda(), shards, bounds, nt, (0, 1), variant)
            for r, (i, d, a) in enumerate(res):
                assert np.array_equal(i, oi) and np.array_equal(d, od), (world, variant, step, r)
                assert a.all()
    finally:
        ranks.close()


def test_loopback_exchange_merge_entry_point_and_key_widths():
    """slm_exchange_merge (local keys given; the producer is exchange_store_kernel): same result with compact and with
    64-bit keys, the per-rank keys equal the oracle's, an empty shard takes part like any other."""
    import torch
    world, nq, nt = 4, 3000, 50_000
    q, t = _case(nq, nt, 77)
    bounds = [(0, 20_000), (20_000, 20_000), (20_000, 35_008), (35_008, nt)]     # rank 1 holds no rows
    shards = [torch.from_numpy(t[a:b]).cuda() if b > a else torch.empty((0, 32), dtype=torch.uint8, device="cuda") for a, b in bounds]
    oi, od = orc.c_knn2(q, t)
    qd = torch.from_numpy(q).cuda()
    for total in (nt, 0):                    # 0 = unknown -> 64-bit keys
        ranks = _Ranks(world, cap=4096)
        try:
            for step in range(3):
                res = ranks.query(qd, shards, bounds, total, (7, 10), "auto", fused=False)
                for r, (i, d, a) in enumerate(res):
                    assert np.array_equal(i, oi) and np.array_equal(d, od), (total, step, r)
                    assert np.array_equal(a, orc.c_ratio(od, 7, 10))
        finally:
            ranks.close()


def test_lost_peer_is_reported_not_trapped():
    """A rank whose peers never publish must not hang or poison the CUDA context: the merge kernel gives up after its
    poll budget, leaves the outputs untouched and the next call reports SLM_ERR_TIMEOUT."""
    import torch
    os.environ["SLM_EXCHANGE_MAX_POLLS"] = "2000"
    try:
        ranks = _Ranks(2, cap=256)
    finally:
        del os.environ["SLM_EXCHANGE_MAX_POLLS"]
    try:
        q, t = _case(64, 5000, 5)
        qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
        ctx = ranks.ctxs[0]
        idx = torch.full((64, 2), -9, dtype=torch.int32, device="cuda")
        kp = ctypes.cast(ranks.key_ptrs, ctypes.c_void_p)
        fp = ctypes.cast(ranks.flag_ptrs, ctypes.c_void_p)
        _lib.check(ctx.lib.slm_knn2_exchange(ctx.handle, qd.data_ptr(), 64, td.data_ptr(), 2500, 0, 256, 5000, kp, fp, 0, 2, 1,
                                             7, 10, idx.data_ptr(), None, None, None))       # rank 1 never runs
        torch.cuda.synchronize()                                                              # returns: no hang, no trap
        assert (idx.cpu().numpy() == -9).all()
        with pytest.raises(slammatch.SlamMatchError) as e:
            _lib.check(ctx.lib.slm_exchange_status(ctx.handle))
        assert e.value.code == -5 and "rank 1" in str(e.value)
        _lib.check(ctx.lib.slm_exchange_status(ctx.handle))                                   # the report is cleared
        # the context is still usable
        i, d, a = slammatch.knn2(q, t)
        oi, od = orc.c_knn2(q, t)
        assert np.array_equal(i, oi) and np.array_equal(d, od)
    finally:
        ranks.close()
