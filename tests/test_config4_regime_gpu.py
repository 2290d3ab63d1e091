"""BASELINE config 4's regime on ONE GPU: hundreds of thousands of descriptors against a 65 536-word binary vocabulary
(the re-specified ``BoW.hist -> kmeans.predict`` of bag_of_words.py:23-26), every variant, bit-exact against the C
oracle.  This is the shape class the small-query tests never reach: >= 148 cluster units of the tensor kernel (several
waves of clusters), tc_refine_kernel<8> over every query, one candidate slot per query, 16-bit word indices.

The oracle (13 G comparisons) is computed once per session; each variant is compared with the same arrays."""
import numpy as np
import pytest

import slammatch
from slammatch import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

NQ, NT = 200_000, 65_536


@pytest.fixture(scope="module")
def c4_case():
    # words: uniform with exact duplicates (lowest word id must win); descriptors: half planted near a word
    q, t = synth.planted(NQ, NT, 4001)
    t = synth.with_duplicates(t, 4002, 0.05)
    q[-7:] = t[-7:]                                   # exact hits on the last words (tail tile of the vocabulary)
    oi, od = orc.c_knn2(q, t)
    return q, t, oi, od


@pytest.mark.parametrize("variant", ["tensor", "tensor4", "auto", "popc", "bmma"])
def test_config4_regime_equals_oracle(c4_case, variant):
    import torch
    q, t, oi, od = c4_case
    ctx = slammatch.context(0)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    try:
        i, d, acc = slammatch.knn2(qd, td, ratio=(7, 10), variant=variant)
    except slammatch.SlamMatchError as e:
        if e.code == -4:
            pytest.skip(f"variant {variant} not built")
        raise
    finally:
        ctx.set_variant("auto")
    torch.cuda.synchronize()
    i, d, acc = i.cpu().numpy(), d.cpu().numpy(), acc.cpu().numpy()
    bad = np.nonzero((i != oi).any(axis=1) | (d != od).any(axis=1))[0]
    assert bad.size == 0, (variant, bad[:8], i[bad[:4]], oi[bad[:4]], d[bad[:4]], od[bad[:4]])
    assert np.array_equal(acc, orc.c_ratio(od, 7, 10))
    assert 0.2 * NQ < acc.sum() < 0.7 * NQ                # planted data: the ratio verdicts are not vacuous
    if variant in ("tensor", "tensor4", "auto"):
        assert ctx.last_kernel().startswith("knn2_tc"), ctx.last_kernel()


def test_config4_regime_word_ids_through_bow(c4_case):
    """The consumer of config 4: word id = idx[:, 0] (slammatch.bow), here through the host path in two batches."""
    q, t, oi, od = c4_case
    i, d, _ = slammatch.knn2(q[:100_000], t, ratio=None)
    assert np.array_equal(i[:, 0], oi[:100_000, 0]) and np.array_equal(d, od[:100_000])


def test_config4_regime_sharded_vocabulary_on_one_gpu(c4_case):
    """Vocabulary row-sharded 8 ways (8192 words per shard, indices < 65 536): per-shard keys -> slm_merge_top2 gives
    the unsharded bytes.  (The multi-GPU form of the same path is tests/test_multi_gpu.py.)"""
    import torch
    q, t, oi, od = c4_case
    nq = 120_000
    ctx = slammatch.context(0)
    qd = torch.from_numpy(q[:nq]).cuda()
    shards = 8
    keys = torch.empty((shards, nq, 2), dtype=torch.int64, device="cuda")
    tds = []
    for s in range(shards):
        a, b = s * NT // shards, (s + 1) * NT // shards
        td = torch.from_numpy(t[a:b]).cuda()
        tds.append(td)
        slammatch._lib.check(ctx.lib.slm_knn2_keys(ctx.handle, qd.data_ptr(), nq, td.data_ptr(), b - a, a,
                                                   keys[s].data_ptr(), None))
    idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    dist = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    acc = torch.empty((nq,), dtype=torch.uint8, device="cuda")
    slammatch._lib.check(ctx.lib.slm_merge_top2(ctx.handle, keys.data_ptr(), shards, nq, 7, 10, idx.data_ptr(),
                                                dist.data_ptr(), acc.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), oi[:nq]) and np.array_equal(dist.cpu().numpy(), od[:nq])
    assert np.array_equal(acc.cpu().numpy(), orc.c_ratio(od[:nq], 7, 10))
