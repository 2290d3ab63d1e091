"""Chained batches (config 3): pairs that share a query frame are walked by one resident cluster (query tiles
expanded once, one candidate flush per pair).  Forced on for small batches through SLM_TC_CHAIN_MIN and checked
pair by pair against the oracle -- shuffled pair order, repeated pairs, both orientations, ragged frame sizes."""
import os

import numpy as np
import pytest

import slammatch
from slammatch import _lib, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _ctx(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return _lib.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def _batched(ctx, desc, pairs, ratio=(7, 10)):
    import torch
    frames, n = desc.shape[0], desc.shape[1]
    dd = torch.from_numpy(desc).cuda()
    P = pairs.shape[0]
    idx = torch.full((P, n, 2), -9, dtype=torch.int32, device="cuda")
    dist = torch.full((P, n, 2), -9, dtype=torch.int32, device="cuda")
    acc = torch.full((P, n), 9, dtype=torch.uint8, device="cuda")
    _lib.check(ctx.lib.slm_knn2_batched(ctx.handle, dd.data_ptr(), frames, n, pairs.ctypes.data, P, ratio[0], ratio[1],
                                        idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), None))
    torch.cuda.synchronize()
    return idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy()


# "tensor" = fp8 kernel (16-row chunks), "tensor4" = mxf4 kernel (240-row tiles, 20-row chunks): both have a chained instantiation
VARIANTS = ["tensor", "tensor4"]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [129, 300, 777, 1100, 2000])
@pytest.mark.parametrize("chain_min", [2, 3, 8])
def test_chained_batches_equal_oracle_pair_by_pair(n, chain_min, variant):
    if n == 2000 and chain_min != 3:
        pytest.skip("the config-3 frame size once per variant")
    ctx = _ctx(SLM_TC_CHAIN_MIN=chain_min, SLM_TC_MAX_CPG=1)      # one cluster per unit: the chained instantiation is taken
    ctx.set_variant(variant)
    frames = 7
    rng = np.random.default_rng(n + chain_min)
    base = synth.heavy_ties(n, 60 + n) if n == 129 else synth.uniform(n, 50 + n)
    desc = np.stack([base ^ np.packbits(rng.random((n, 256)) < 0.04 * (f + 1), axis=1, bitorder="little")
                     for f in range(frames)])
    pairs = [(i, j) for i in range(frames) for j in range(i + 1, frames)]
    pairs += [(5, 1), (5, 0), (3, 3), (0, 6), (0, 6)]            # reversed, self, repeated
    pairs = np.array(pairs, dtype=np.int32)[rng.permutation(len(pairs))]
    idx, dist, acc = _batched(ctx, desc, pairs)
    for p, (a, b) in enumerate(pairs):
        oi, od = orc.c_knn2(desc[a], desc[b])
        assert np.array_equal(idx[p], oi) and np.array_equal(dist[p], od), (n, chain_min, p, a, b)
        assert np.array_equal(acc[p], orc.c_ratio(od, 7, 10)), (n, chain_min, p)
    assert ctx.last_kernel() == ("knn2_tc4_kernel" if variant == "tensor4" else "knn2_tc2_kernel")
    ctx.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_chain_plan_is_invisible_in_the_results(variant):
    """Same batch with chaining off and on: byte-identical outputs (without a chain plan the batch runs on the fp8 kernel)."""
    frames, n = 9, 520
    rng = np.random.default_rng(3)
    desc = np.stack([synth.uniform(n, 900 + f) for f in range(frames)])
    desc[4, :200] = desc[2, 100:300]                      # shared rows between frames: exact duplicates
    pairs = np.array([(i, j) for i in range(frames) for j in range(frames) if i != j], dtype=np.int32)
    off, on = _ctx(SLM_TC_CHAIN=0), _ctx(SLM_TC_CHAIN_MIN=5)
    for c in (off, on):
        c.set_variant(variant)
    a, b = _batched(off, desc, pairs), _batched(on, desc, pairs)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[2].sum() > 0
    off.close()
    on.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [260, 1000])
def test_frame_refine_kernel_many_pairs(n, variant):
    """>= 148 pairs switch the refine step to one CTA per pair with the train frame in shared memory; with and
    without chaining, against the oracle and against the L2-fed refine kernel."""
    frames = 18
    rng = np.random.default_rng(n)
    base = synth.uniform(n, 70 + n)
    desc = np.stack([base ^ np.packbits(rng.random((n, 256)) < 0.03 * (1 + f % 4), axis=1, bitorder="little")
                     for f in range(frames)])
    desc[7, : n // 2] = desc[3, n // 4: n // 4 + n // 2]          # exact duplicates across frames
    pairs = np.array([(i, j) for i in range(frames) for j in range(i + 1, frames)], dtype=np.int32)   # 153 pairs
    ref = None
    for env in (dict(SLM_TC_CHAIN_MIN=4), dict(SLM_TC_CHAIN=0), dict(SLM_TC_NO_FRAME_REFINE=1, SLM_TC_CHAIN=0)):
        ctx = _ctx(**env)
        ctx.set_variant(variant)
        out = _batched(ctx, desc, pairs)
        ctx.close()
        if ref is None:
            ref = out
            for p in range(0, pairs.shape[0], 7):
                a, b = pairs[p]
                oi, od = orc.c_knn2(desc[a], desc[b])
                assert np.array_equal(out[0][p], oi) and np.array_equal(out[1][p], od), (n, p)
                assert np.array_equal(out[2][p], orc.c_ratio(od, 7, 10)), (n, p)
        else:
            for x, y in zip(ref, out):
                assert np.array_equal(x, y), env


def test_public_knn2_batched_host_device_and_out_buffers():
    """slammatch.knn2_batched (the public call of config 3): numpy in / numpy out, pinned CPU tensors in with caller-owned
    pinned result buffers, CUDA tensors in / out -- all equal to the oracle pair by pair; argument errors are ValueErrors."""
    import torch
    frames, n = 5, 700
    rng = np.random.default_rng(11)
    base = synth.uniform(n, 60)
    desc = np.stack([base ^ np.packbits(rng.random((n, 256)) < 0.04 * (f + 1), axis=1, bitorder="little") for f in range(frames)])
    pairs = np.array([(i, j) for i in range(frames) for j in range(frames) if i != j], dtype=np.int32)   # ordered pairs
    P = pairs.shape[0]
    want = [orc.c_knn2(desc[a], desc[b]) for a, b in pairs]

    def check(idx, dist, acc):
        for p in range(P):
            oi, od = want[p]
            assert np.array_equal(idx[p], oi) and np.array_equal(dist[p], od), p
            assert np.array_equal(acc[p], orc.c_ratio(od, 7, 10)), p

    check(*slammatch.knn2_batched(desc, pairs))
    out = (torch.empty((P, n, 2), dtype=torch.int32).pin_memory(), torch.empty((P, n, 2), dtype=torch.int32).pin_memory(),
           torch.empty((P, n), dtype=torch.uint8).pin_memory())
    for _ in range(2):
        res = slammatch.knn2_batched(torch.from_numpy(desc).pin_memory(), pairs, ratio=(7, 10), out=out)
        check(*res)
        assert res[0].ctypes.data == out[0].data_ptr()          # written in place
    out_np = (np.empty((P, n, 2), np.int32), np.empty((P, n, 2), np.int32), np.empty((P, n), np.uint8))
    check(*slammatch.knn2_batched(desc, pairs, out=out_np))
    check(*out_np)
    di, dd_, da = slammatch.knn2_batched(torch.from_numpy(desc).cuda(), pairs)
    assert di.is_cuda
    check(di.cpu().numpy(), dd_.cpu().numpy(), da.cpu().numpy())
    i0, d0, a0 = slammatch.knn2_batched(desc, np.zeros((0, 2), np.int32))
    assert i0.shape == (0, n, 2) and a0.shape == (0, n)
    with pytest.raises(ValueError):
        slammatch.knn2_batched(desc, np.array([[0, frames]], np.int32))
    with pytest.raises(ValueError):
        slammatch.knn2_batched(desc[:, :, :16], pairs)
    with pytest.raises(ValueError):
        slammatch.knn2_batched(desc, pairs, out=(out_np[0], out_np[1], np.empty((P, n + 1), np.uint8)))
