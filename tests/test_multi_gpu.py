"""Sharded train set over >= 2 real GPUs with NCCL (skipped on a single-GPU box; the single-GPU loopback of the same
kernels is tests/test_exchange_loopback_gpu.py): every rank must return the oracle's result bit-exactly, i.e. the same
bytes as the 1-GPU path, for 2..N shards, every exchange (NCCL all-gather, NVLink peer stores, all-to-all), both key
widths, and query batches above the old 8192-query limit of the NVLink exchange."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (name, nq, nt, seed): config-5-like (few queries, long DB, 64-bit keys) and config-4-like (many queries, <= 65536 words,
# 32-bit keys on the NVLink exchange; two-phase refine from 4 ranks on)
CASES = [("c5like", 512, 300_000, 17), ("c4like", 40_000, 60_000, 19)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, exchange):
    import torch
    import torch.distributed as dist
    import slammatch
    from slammatch import synth
    from slammatch.sharded import QueryShardedMatcher, ShardedMatcher, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        used = {}
        for name, nq, nt, seed in CASES:
            q, t = synth.planted(nq, nt, seed)
            t = synth.with_duplicates(t, seed + 1, 0.2)
            a, b = shard_bounds(nt, world)[rank]
            sm = ShardedMatcher(torch.from_numpy(t[a:b]).cuda(), a, ratio=(7, 10), exchange=exchange)
            qd = torch.from_numpy(q).cuda()
            for _ in range(5):                    # several steps: both halves of the exchange buffers get reused
                idx, dd, acc = sm.knn2(qd)
            torch.cuda.synchronize()
            sm.check()
            used[name] = sm.last_exchange
            np.savez(os.path.join(out_dir, f"{name}_rank{rank}.npz"), idx=idx.cpu().numpy(), dist=dd.cpu().numpy(),
                     acc=acc.cpu().numpy())
            if name == "c4like" and exchange == "auto":
                # the other decomposition of the same search: train set replicated, queries sliced, results all-gathered;
                # a ragged query count (the last slices are shorter / empty)
                qs = QueryShardedMatcher(torch.from_numpy(t).cuda(), ratio=(7, 10))
                for n_q in (nq, 5, nq - 3):
                    for _ in range(2):
                        qi, qd_, qa = qs.knn2(qd[:n_q])
                    torch.cuda.synchronize()
                    assert torch.equal(qi, idx[:n_q]) and torch.equal(qd_, dd[:n_q]) and torch.equal(qa, acc[:n_q]), n_q
                assert "all-gather" in qs.last_exchange
        # the sharded keyframe DB (round-robin keyframes, rebased keys) on the same ranks
        frames = [synth.uniform(n, 4000 + i) for i, n in enumerate((700, 3, 1200, 0, 64, 900, 333))]
        db = slammatch.ShardedKeyframeDB(device=rank, capacity=256)
        for f in frames:
            db.add(f)
        qk = np.concatenate(frames)[::7][:300] ^ np.uint8(2)
        i, d, a = db.query(qk)
        np.savez(os.path.join(out_dir, f"kfdb_rank{rank}.npz"), idx=i, dist=d, acc=a)
        if rank == 0:
            open(os.path.join(out_dir, "exchange.txt"), "w").write(repr(used))
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_query_equals_oracle(tmp_path):
    import torch
    import torch.multiprocessing as mp
    from slammatch import synth
    from oracle import oracle as orc
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    want = {}
    for name, nq, nt, seed in CASES:
        q, t = synth.planted(nq, nt, seed)
        t = synth.with_duplicates(t, seed + 1, 0.2)
        want[name] = orc.c_knn2(q, t)
    frames = [synth.uniform(m, 4000 + i) for i, m in enumerate((700, 3, 1200, 0, 64, 900, 333))]
    flat = np.concatenate(frames)
    qk = flat[::7][:300] ^ np.uint8(2)
    want["kfdb"] = orc.c_knn2(qk, flat)
    # SLM_TEST_WORLDS=8 restricts the rank counts (the 2-rank pass is also what the single-GPU loopback tests cover)
    worlds = [int(x) for x in os.environ.get("SLM_TEST_WORLDS", "").split(",") if x] or sorted({2, n})
    for world in worlds:
        # SLM_TEST_EXCHANGES=auto restricts the exchanges (quick re-validation of one path)
        for exchange in [x for x in os.environ.get("SLM_TEST_EXCHANGES", "").split(",") if x] or ("nccl", "auto", "a2a"):
            mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), exchange), nprocs=world, join=True)
            print("world", world, "exchange requested", exchange, "used", open(tmp_path / "exchange.txt").read())
            for r in range(world):
                for name in list(x[0] for x in CASES) + ["kfdb"]:
                    z = np.load(tmp_path / f"{name}_rank{r}.npz")
                    oi, od = want[name]
                    assert np.array_equal(z["idx"], oi) and np.array_equal(z["dist"], od), (world, r, exchange, name)
                    assert np.array_equal(z["acc"], orc.c_ratio(od, 7, 10)), (world, r, exchange, name)
