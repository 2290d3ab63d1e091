"""Sharded train set over >= 2 real GPUs with NCCL (skipped on a single-GPU box): every rank must return the
oracle's result bit-exactly, i.e. the same bytes as the 1-GPU path, for 2..N shards."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, t, out_dir, exchange="auto"):
    import torch
    import torch.distributed as dist
    from slammatch.sharded import ShardedMatcher, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        a, b = shard_bounds(t.shape[0], world)[rank]
        sm = ShardedMatcher(torch.from_numpy(t[a:b]).cuda(), a, ratio=(7, 10), exchange=exchange)
        qd = torch.from_numpy(q).cuda()
        for _ in range(5):                    # several steps: both halves of the exchange buffers get reused
            idx, dd, acc = sm.knn2(qd)
        if rank == 0:
            open(os.path.join(out_dir, "exchange.txt"), "w").write(sm.last_exchange)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx.cpu().numpy(), dist=dd.cpu().numpy(),
                 acc=acc.cpu().numpy())
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
    q, t = synth.planted(512, 300_000, 17)
    t = synth.with_duplicates(t, 18, 0.2)
    oi, od = orc.c_knn2(q, t)
    for world in sorted({2, n}):
        for exchange in ("nccl", "auto", "a2a"):
            mp.spawn(_worker, args=(world, _free_port(), q, t, str(tmp_path), exchange), nprocs=world, join=True)
            print("world", world, "exchange requested", exchange, "used", open(tmp_path / "exchange.txt").read())
            for r in range(world):
                z = np.load(tmp_path / f"rank{r}.npz")
                assert np.array_equal(z["idx"], oi) and np.array_equal(z["dist"], od), (world, r, exchange)
                assert np.array_equal(z["acc"], orc.c_ratio(od, 7, 10)), (world, r, exchange)
