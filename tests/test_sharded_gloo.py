"""The N>1 host path on CPU: two gloo ranks, each holding a contiguous block of the train set, exchange
packed top-2 keys with ONE all-gather and merge by unsigned key order.  The CUDA entry points are replaced
by the oracle's key functions (test infrastructure), so this checks the collective plumbing of
slammatch.sharded.ShardedMatcher: shard bounds, gather layout, merge order, shard-count invariance."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slammatch import synth
from slammatch.sharded import QueryShardedMatcher, ShardedMatcher, shard_bounds
from oracle import oracle as orc


def test_shard_bounds_cover_rows_contiguously():
    for n, w in ((10_000_000, 8), (65536, 8), (100, 3), (5, 8), (0, 2), (33, 1)):
        b = shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(x[0] <= x[1] for x in b)
        assert all(x[0] % 32 == 0 for x in b)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, t, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_bounds(t.shape[0], world)[rank]
        shard = t[a:b]

        def local_keys(qq):
            return torch.from_numpy(orc.np_knn2_keys(qq.numpy(), shard, train_index_base=a).view(np.int64))

        def merge(g):
            keys = orc.np_merge_top2(g.numpy().view(np.uint64))
            idx, dd = orc.keys_to_idx_dist(keys)
            return torch.from_numpy(idx), torch.from_numpy(dd), torch.from_numpy(orc.np_ratio(dd, 7, 10))

        sm = ShardedMatcher(torch.from_numpy(shard), a, local_keys=local_keys, merge=merge)
        assert sm.world == world
        idx, dd, acc = sm.knn2(torch.from_numpy(q))
        # batched form (config 4 path): the all-gather of batch b overlaps the search of batch b+1
        bi, bd, ba = sm.knn2(torch.from_numpy(q), query_batch=50)
        assert np.array_equal(np.asarray(bi), np.asarray(idx)) and np.array_equal(np.asarray(bd), np.asarray(dd))
        assert np.array_equal(np.asarray(ba), np.asarray(acc))
        # all-to-all of query slices + all-gather of merged results (large-query exchange), nq not divisible by world
        sa = ShardedMatcher(torch.from_numpy(shard), a, local_keys=local_keys, merge=merge, exchange="a2a")
        ai, ad, aa = sa.knn2(torch.from_numpy(q))
        assert "all-to-all" in sa.last_exchange
        assert np.array_equal(np.asarray(ai), np.asarray(idx)) and np.array_equal(np.asarray(ad), np.asarray(dd))
        assert np.array_equal(np.asarray(aa), np.asarray(acc))
        ai, ad, aa = sa.knn2(torch.from_numpy(q[:1]))          # fewer queries than ranks
        assert np.array_equal(np.asarray(ai), np.asarray(idx)[:1]) and np.array_equal(np.asarray(aa), np.asarray(acc)[:1])
        # query-sharded form: train set replicated, query slices, in-place all-gather of the finished results
        def knn2_slice(qq, o_idx, o_dist, o_acc):
            if qq.shape[0]:
                i, d = orc.np_knn2(qq.numpy(), t)
                o_idx.copy_(torch.from_numpy(i))
                o_dist.copy_(torch.from_numpy(d))
                o_acc.copy_(torch.from_numpy(orc.np_ratio(d, 7, 10)))

        qs = QueryShardedMatcher(torch.from_numpy(t), knn2_slice=knn2_slice)
        assert qs.world == world and qs.rank == rank
        for n_q in (q.shape[0], q.shape[0] - 1, 1):
            qi, qd, qa = qs.knn2(torch.from_numpy(q[:n_q]))
            assert np.array_equal(np.asarray(qi), np.asarray(idx)[:n_q]) and np.array_equal(np.asarray(qd), np.asarray(dd)[:n_q])
            assert np.array_equal(np.asarray(qa), np.asarray(acc)[:n_q])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx, dist=dd, acc=acc)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_matcher_all_gather_merge_under_gloo(tmp_path, world):
    q, t = synth.planted(120, 1500, 7)
    t = synth.with_duplicates(t, 8, 0.4)          # exact duplicates across the shard boundary
    mp.spawn(_worker, args=(world, _free_port(), q, t, str(tmp_path)), nprocs=world, join=True)
    oi, od = orc.c_knn2(q, t)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["idx"], oi) and np.array_equal(z["dist"], od), r
        assert np.array_equal(z["acc"], orc.c_ratio(od, 7, 10)), r
