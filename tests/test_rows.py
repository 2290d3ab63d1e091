"""Host-side result rows: slammatch._rows (C) must build exactly what the plain cv2.DMatch constructor builds --
row lengths min(k, neighbours), crossCheck-style empty rows, (imgIdx, local trainIdx) for multi-image collections.
CPU only (no search involved)."""
import numpy as np
import pytest

from slammatch import matcher


def _plain(idx, dist, keep, k, offsets):
    saved, matcher._fast_rows = matcher._fast_rows, False
    try:
        return matcher.Matcher._rows(idx, dist, keep, k, offsets)
    finally:
        matcher._fast_rows = saved


def _as_lists(rows):
    return [[(type(m).__name__, m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in r] for r in rows]


def test_c_rows_equal_python_rows():
    pytest.importorskip("cv2")
    matcher._fast_rows = None
    matcher._init_fast_rows()
    assert matcher._fast_rows, "slammatch._rows failed its layout self-test"
    rng = np.random.default_rng(0)
    for nq in (0, 1, 7, 1000):
        idx = rng.integers(0, 900, (nq, 2)).astype(np.int32)
        dist = rng.integers(0, 257, (nq, 2)).astype(np.int32)
        if nq > 5:
            idx[3, 1] = -1
            dist[3, 1] = -1
            idx[5] = -1
            dist[5] = -1
        keep = (rng.random(nq) < 0.5).astype(np.uint8)
        offsets = np.array([0, 100, 101, 500, 900])
        for k in (1, 2):
            for kp in (None, keep):
                for off in (None, offsets):
                    fast = matcher.Matcher._rows(idx, dist, kp, k, off)
                    assert isinstance(fast, tuple) and all(isinstance(r, tuple) for r in fast)
                    assert _as_lists(fast) == _as_lists(_plain(idx, dist, kp, k, off)), (nq, k, kp is not None, off is not None)


def test_reference_loop_on_c_rows_stops_at_short_row():
    """`for m, n in matches` raises ValueError at the first row with fewer than two neighbours (tracking.py:25-30)."""
    pytest.importorskip("cv2")
    idx = np.array([[1, 2], [3, 4], [5, -1], [6, 7]], dtype=np.int32)
    dist = np.array([[10, 50], [40, 41], [0, -1], [1, 99]], dtype=np.int32)
    rows = matcher.Matcher._rows(idx, dist, None, 2, None)
    good = []
    with pytest.raises(ValueError):
        for m, n in rows:
            if m.distance < 0.7 * n.distance:
                good.append(m)
    assert [(g.queryIdx, g.trainIdx) for g in good] == [(0, 1)]
