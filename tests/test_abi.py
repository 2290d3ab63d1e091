"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/slammatch.h
declares, fails loudly without a GPU, and the Python mirror validates arguments like OpenCV does.
No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest

import slammatch
from slammatch import _lib, matcher
from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "slammatch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(slammatch.LIB_PATH)
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/slammatch.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding and header disagree on the symbol list"


def test_version_and_error_string():
    lib = slammatch.load()
    assert lib.slm_version() == 200
    assert isinstance(lib.slm_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu():
    lib = slammatch.load()
    assert lib.slm_create(0, None) == -1
    assert b"NULL" in lib.slm_last_error()
    assert lib.slm_set_variant(None, 1) == -1
    assert lib.slm_knn2_host(None, None, 1, None, 1, 7, 10, 0, None, None, None) == -1
    assert lib.slm_knn2_masked(None, None, 1, None, 1, 0, None, 1, 7, 10, None, None, None, None) == -1


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    q = np.zeros((4, 32), np.uint8)
    with pytest.raises(slammatch.SlamMatchError) as e:
        slammatch.knn2(q, q)
    assert "no CPU path" in str(e.value)


def test_matcher_validates_inputs_like_opencv():
    m = slammatch.Matcher(indexParams=dict(algorithm=6, table_number=6, key_size=12, multi_probe_level=1),
                          searchParams=dict(checks=50))
    good = np.zeros((3, 32), np.uint8)
    with pytest.raises(ValueError):
        m.knnMatch(good.astype(np.float32), good, k=2)       # dtype mismatch (cv2.error in OpenCV)
    with pytest.raises(ValueError):
        m.knnMatch(good, np.zeros((3, 16), np.uint8), k=2)   # column mismatch
    with pytest.raises(ValueError):
        m.knnMatch(good, good, k=3)
    with pytest.raises(ValueError):
        m.knnMatch(good, good, k=2, mask=np.ones((3, 4), np.uint8))   # OpenCV: matchers.cpp:639 asserts the mask's shape
    with pytest.raises(ValueError):
        m.knnMatch(good, good, k=2, mask=np.ones((3, 3), np.int32))   # ... and CV_8UC1
    with pytest.raises(ValueError):
        slammatch.Matcher(crossCheck=True).knnMatch(good, good, k=1, mask=np.ones((3, 3), np.uint8))  # mask.empty() asserted
    with pytest.raises(ValueError):
        slammatch.knn2(good, good, cross_check=True, mask=np.ones((3, 3), np.uint8))
    with pytest.raises(ValueError):
        slammatch.Matcher(crossCheck=True).knnMatch(good, good, k=2)  # OpenCV asserts K == 1
    with pytest.raises(ValueError):
        slammatch.knn2(good, good, ratio=(0, 10))


def test_rows_layout_matches_opencv_short_rows():
    idx = np.array([[5, 2], [7, -1], [-1, -1]], np.int32)
    dist = np.array([[1, 9], [3, -1], [-1, -1]], np.int32)
    rows = matcher.Matcher._rows(idx, dist, None, 2, None)
    assert [len(r) for r in rows] == [2, 1, 0]
    assert (rows[0][0].queryIdx, rows[0][0].trainIdx, rows[0][0].distance) == (0, 5, 1.0)
    assert (rows[0][1].trainIdx, rows[1][0].queryIdx) == (2, 1)
    # multi-image collection: global index -> (imgIdx, local trainIdx)
    offsets = np.array([0, 4, 6, 10])
    rows = matcher.Matcher._rows(idx, dist, None, 2, offsets)
    assert (rows[0][0].imgIdx, rows[0][0].trainIdx) == (1, 1)
    assert (rows[0][1].imgIdx, rows[0][1].trainIdx) == (0, 2)
    assert (rows[1][0].imgIdx, rows[1][0].trainIdx) == (2, 1)


def test_install_rebinds_the_seam_and_uninstall_restores():
    cv2 = pytest.importorskip("cv2")
    orig = cv2.FlannBasedMatcher
    slammatch.install(cv2)
    try:
        assert cv2.FlannBasedMatcher is slammatch.Matcher
        m = cv2.FlannBasedMatcher(indexParams=dict(algorithm=6), searchParams=dict(checks=50))
        assert isinstance(m, slammatch.Matcher)
    finally:
        slammatch.uninstall(cv2)
    assert cv2.FlannBasedMatcher is orig


def test_synthetic_generators_are_deterministic():
    from slammatch import synth
    a, b = synth.planted(50, 60, 3)
    a2, b2 = synth.planted(50, 60, 3)
    assert np.array_equal(a, a2) and np.array_equal(b, b2)
    assert synth.heavy_ties(10, 1).shape == (10, 32) and synth.uniform(0, 1).shape == (0, 32)
