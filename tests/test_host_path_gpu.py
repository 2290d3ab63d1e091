"""slm_knn2_host with PAGEABLE inputs (what the drop-in caller hands over: numpy arrays from ORB, orb.py:23-24): long train
sets and long query sets are staged through the pinned ring by host threads (csrc/host_stager.h) -- same bytes out as with
the driver's own staging (SLM_HOST_STAGE_THREADS=0), as with pinned inputs, and as the oracle."""
import os

import numpy as np
import pytest

from slammatch import _lib, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _ctx(threads):
    old = os.environ.get("SLM_HOST_STAGE_THREADS")
    os.environ["SLM_HOST_STAGE_THREADS"] = str(threads)
    try:
        return _lib.Context(0)
    finally:
        if old is None:
            del os.environ["SLM_HOST_STAGE_THREADS"]
        else:
            os.environ["SLM_HOST_STAGE_THREADS"] = old


def _host_call(ctx, q, t, ratio=(7, 10)):
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -7, np.int32)
    dist = np.full((nq, 2), -7, np.int32)
    acc = np.full(nq, 7, np.uint8)
    _lib.check(ctx.lib.slm_knn2_host(ctx.handle, q.ctypes.data, nq, t.ctypes.data, nt, ratio[0], ratio[1], 0,
                                     idx.ctypes.data, dist.ctypes.data, acc.ctypes.data))
    return idx, dist, acc


def test_pageable_long_train_set_through_the_staging_ring():
    """5.x chunks of 2^20 rows: more chunks than ring slots, a ragged last chunk."""
    import torch
    nq, nt = 64, 5 * (1 << 20) + 12345
    q, t = synth.planted(nq, nt, 4242)
    oi, od = orc.c_knn2(q, t)
    want = orc.c_ratio(od, 7, 10)
    for threads in (0, 3, 8):
        ctx = _ctx(threads)
        try:
            for _ in range(2):                                   # the ring and its events are reused from call to call
                i, d, a = _host_call(ctx, q, t)
                assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, want), threads
        finally:
            ctx.close()
    # pinned inputs take the direct path
    ctx = _ctx(4)
    try:
        tp = torch.from_numpy(t).pin_memory()
        i, d, a = _host_call(ctx, q, tp.numpy())
        assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, want)
    finally:
        ctx.close()


def test_pageable_long_query_set_and_mid_size_train_set():
    """Config-4-like: 600 000 pageable queries (19 MB, staged in 8 MB pieces) against a small vocabulary; and a 2.5 M-row
    pageable train set below the chunked-search threshold of three chunks."""
    ctx = _ctx(4)
    try:
        q, t = synth.planted(600_000, 300, 77)
        oi, od = orc.c_knn2(q, t)
        i, d, a = _host_call(ctx, q, t, ratio=(3, 4))
        assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, orc.c_ratio(od, 3, 4))
        q, t = synth.planted(40, 2 * (1 << 20) - 5, 78)
        oi, od = orc.c_knn2(q, t)
        i, d, a = _host_call(ctx, q, t)
        assert np.array_equal(i, oi) and np.array_equal(d, od) and np.array_equal(a, orc.c_ratio(od, 7, 10))
    finally:
        ctx.close()
