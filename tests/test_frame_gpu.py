"""Single-launch frame-to-frame kernel (csrc/knn2_frame.cu): forced on for every shape through its own ctx
(SLM_FRAME_MAX_CLK is read at slm_create) and compared bit for bit with the oracle -- kNN-2, ratio verdicts,
cross-check, packed keys with an index base, multi-tile train slices, repeated launches (ticket reset)."""
import os

import numpy as np
import pytest

import slammatch
from slammatch import _lib, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame_ctx():
    old = os.environ.get("SLM_FRAME_MAX_CLK")
    os.environ["SLM_FRAME_MAX_CLK"] = str(1 << 60)
    try:
        ctx = _lib.Context(0)
    finally:
        if old is None:
            del os.environ["SLM_FRAME_MAX_CLK"]
        else:
            os.environ["SLM_FRAME_MAX_CLK"] = old
    yield ctx
    ctx.close()


def _run(ctx, q, t, ratio, cross, base=0):
    import torch
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    nq, nt = q.shape[0], t.shape[0]
    idx = torch.full((nq, 2), -9, dtype=torch.int32, device="cuda")
    dist = torch.full((nq, 2), -9, dtype=torch.int32, device="cuda")
    acc = torch.full((nq,), 9, dtype=torch.uint8, device="cuda")
    num, den = ratio if ratio else (0, 1)
    _lib.check(ctx.lib.slm_knn2_filter(ctx.handle, qd.data_ptr(), nq, td.data_ptr(), nt, base, num, den, int(cross),
                                       idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), None))
    torch.cuda.synchronize()
    return idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy()


SHAPES = [(1000, 1000), (2000, 2000), (1, 1), (9, 1), (1, 9), (33, 5), (129, 4097), (500, 9000), (64, 20000),
          (31, 513), (257, 255), (4100, 300)]


@pytest.mark.parametrize("nq,nt", SHAPES)
@pytest.mark.parametrize("cross", [False, True])
def test_frame_kernel_equals_oracle(frame_ctx, nq, nt, cross):
    if (nq + nt) % 3 == 0:
        q, t = synth.heavy_ties(nq, nq + 11), synth.heavy_ties(nt, nt + 12)
    else:
        q, t = synth.planted(nq, nt, 7000 + nq + nt)
        t = synth.with_duplicates(t, nq, 0.3)
    oi, od = orc.c_knn2(q, t)
    for ratio in ((7, 10), None):
        i, d, a = _run(frame_ctx, q, t, ratio, cross)
        assert frame_ctx.last_kernel() == "knn2_frame_kernel"
        assert np.array_equal(i, oi) and np.array_equal(d, od), (nq, nt, cross)
        want = orc.c_ratio(od, *ratio) if ratio else (oi[:, 0] >= 0).astype(np.uint8)
        if cross:
            want = want & orc.c_cross_check(q, t, oi)
        assert np.array_equal(a, want), (nq, nt, cross, ratio)


def test_frame_kernel_keys_with_index_base(frame_ctx):
    import torch
    q, t = synth.planted(300, 777, 99)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    keys = torch.zeros((300, 2), dtype=torch.int64, device="cuda")
    _lib.check(frame_ctx.lib.slm_knn2_keys(frame_ctx.handle, qd.data_ptr(), 300, td.data_ptr(), 777, 123456,
                                           keys.data_ptr(), None))
    torch.cuda.synchronize()
    assert frame_ctx.last_kernel() == "knn2_frame_kernel"
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), orc.np_knn2_keys(q, t, 123456))
    i, d, a = _run(frame_ctx, q, t, (3, 4), False, base=123456)
    oi, od = orc.c_knn2(q, t, train_index_base=123456)
    assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, orc.c_ratio(od, 3, 4))


def test_frame_kernel_repeated_cross_check_launches(frame_ctx):
    """The last-CTA ticket counter must be back at zero after every launch."""
    q, t = synth.planted(700, 650, 31)
    oi, od = orc.c_knn2(q, t)
    want = orc.c_ratio(od, 7, 10) & orc.c_cross_check(q, t, oi)
    for _ in range(5):
        i, d, a = _run(frame_ctx, q, t, (7, 10), True)
        assert np.array_equal(i, oi) and np.array_equal(a, want)


def test_auto_policy_uses_frame_kernel_for_the_reference_shape_only():
    """BASELINE config 1 (1000 x 1000) runs in one launch; the loop-closure shape stays on the tensor pipe."""
    ctx = slammatch.context(0)
    ctx.set_variant("auto")
    q, t = synth.planted(1000, 1000, 1)
    before = ctx.launch_count()
    slammatch.knn2(q, t, ratio=(3, 4))
    assert ctx.last_kernel() == "knn2_frame_kernel" and ctx.launch_count() - before == 1
    q, t = synth.planted(2000, 200000, 2)
    slammatch.knn2(q, t)
    assert ctx.last_kernel() in ("knn2_tc4_kernel", "knn2_tc2_kernel")      # mxf4 (default) or fp8 (SLM_TC_FP4=0)


@pytest.mark.parametrize("variant", ["auto", "tensor", "popc"])
@pytest.mark.parametrize("nq,nt", [(2000, 20000), (300, 301), (1000, 40000), (5000, 9000), (64, 3000)])
def test_reduced_reverse_search_cross_check(variant, nq, nt):
    """Cross-check with nq < nt searches only the train rows that are some query's best match (nq x nq) -- through
    the frame kernel for frame-sized nq under AUTO, through gather + search + finalize otherwise; duplicated train
    rows and heavy ties make the lowest-index rules on both sides matter."""
    import torch
    ctx = slammatch.context(0)
    if nq % 2:
        q, t = synth.heavy_ties(nq, nq + 5), synth.heavy_ties(nt, nt + 6)
    else:
        q, t = synth.planted(nq, nt, nq + nt)
        t = synth.with_duplicates(t, 3, 0.3)
        q[1::7] = q[0]                      # duplicated queries: only the lowest query index can be mutual
    oi, od = orc.c_knn2(q, t)
    for ratio in ((7, 10), None):
        want = (orc.c_ratio(od, *ratio) if ratio else np.ones(nq, np.uint8)) & orc.c_cross_check(q, t, oi)
        try:
            i, d, a = slammatch.knn2(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), ratio=ratio,
                                     cross_check=True, variant=variant, train_index_base=1000)
        finally:
            ctx.set_variant("auto")
        assert np.array_equal(i.cpu().numpy(), oi + 1000) and np.array_equal(d.cpu().numpy(), od)
        assert np.array_equal(a.cpu().numpy(), want), (variant, nq, nt, ratio)


@pytest.mark.skipif(not os.environ.get("SLM_RUN_EXPERIMENTAL"), reason="experimental planner: opt-in (SLM_RUN_EXPERIMENTAL=1)")
@pytest.mark.parametrize("nq,nt", [(2000, 20000), (300, 5000), (1000, 100000), (4000, 3000), (520, 70001)])
def test_experimental_mt_planner_parity(nq, nt):
    """SLM_TC_PLAN_MT=1 lets the tensor kernel's planner also choose the query tiles per CTA;
    off by default until calibrated, so this test only runs on request."""
    import torch
    os.environ["SLM_TC_PLAN_MT"] = "1"
    try:
        ctx = _lib.Context(0)
    finally:
        del os.environ["SLM_TC_PLAN_MT"]
    ctx.set_variant("tensor")
    q, t = synth.planted(nq, nt, nq + nt)
    t = synth.with_duplicates(t, 5, 0.3)
    i, d, a = _run(ctx, q, t, (7, 10), False)
    oi, od = orc.c_knn2(q, t)
    assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, orc.c_ratio(od, 7, 10))
    ctx.close()
