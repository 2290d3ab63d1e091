#!/usr/bin/env python
"""Generate tests/golden/*.npz -- run HERE (CPU container), where /root/reference and cv2 exist.

The reference ships no golden vectors, so these are produced from

  (1) cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2): the exhaustive member of the OpenCV
      matcher family the reference calls at tracking.py:22 / keypoint.py:44 / Point3D.py:40;
  (2) cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t) for the cross-check definition, and
      cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2, mask=M) for the optional mask argument;
  (3) the reference's OWN functions, imported unmodified from /root/reference:
      tracking.get_matches (tracking.py:12-34) and Point3D.find_2D_and_3D_correspondenses
      (Point3D.py:33-54), with ``cv2.FlannBasedMatcher`` rebound to the exhaustive matcher
      (the seam verified in SURVEY.md section 3.2) so that the approximate LSH index does not make
      the vectors non-deterministic.

/root/reference does not exist on the GPU box; the tests only read the committed .npz files.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "slam-1_b200"))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
from slammatch import synth  # noqa: E402

REFERENCE = "/root/reference"


def rows_to_arrays(rows, nq):
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    for i, row in enumerate(rows):
        for c, m in enumerate(row[:2]):
            idx[i, c] = m.trainIdx
            dist[i, c] = int(round(m.distance))
    return idx, dist


def bf_knn2(q, t):
    nq = q.shape[0]
    if nq == 0 or t.shape[0] == 0:
        return np.full((nq, 2), -1, np.int32), np.full((nq, 2), -1, np.int32)
    return rows_to_arrays(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2), nq)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# (name, generator, nq, nt, seed) -- sizes from SURVEY.md section 4, none a multiple of a tile
KNN_CASES = [
    ("uniform", 0, 5, 1), ("uniform", 1, 0, 2), ("uniform", 1, 1, 3), ("uniform", 1, 2, 4),
    ("uniform", 7, 3, 5), ("uniform", 257, 1000, 6), ("uniform", 1000, 1000, 7),
    ("uniform", 1000, 4099, 8), ("uniform", 2000, 20000, 9), ("uniform", 3, 70001, 10),
    ("planted", 1000, 1000, 11), ("planted", 2000, 20000, 12), ("planted", 257, 4099, 13),
    ("ties", 257, 1000, 14), ("ties", 1000, 4099, 15), ("ties", 64, 2, 16),
    ("dups", 500, 3000, 17), ("dups", 1000, 1000, 18),
]


def make_inputs(kind, nq, nt, seed):
    if kind == "uniform":
        return synth.uniform(nq, seed), synth.uniform(nt, seed + 100000)
    if kind == "planted":
        return synth.planted(nq, nt, seed)
    if kind == "ties":
        return synth.heavy_ties(nq, seed), synth.heavy_ties(nt, seed + 100000)
    if kind == "dups":
        q, t = synth.planted(nq, nt, seed)
        return q, synth.with_duplicates(t, seed + 1, 0.3)
    raise ValueError(kind)


def gen_knn2():
    out = {}
    names = []
    for kind, nq, nt, seed in KNN_CASES:
        q, t = make_inputs(kind, nq, nt, seed)
        idx, dist = bf_knn2(q, t)
        name = f"{kind}_{nq}x{nt}_s{seed}"
        names.append(name)
        out[name + "/idx"] = idx
        out[name + "/dist"] = dist
        out[name + "/sha_q"] = np.array(sha(q))
        out[name + "/sha_t"] = np.array(sha(t))
        # small inputs are stored verbatim so the fixture does not depend on the generator
        if q.nbytes + t.nbytes <= 48 * 1024:
            out[name + "/q"] = q
            out[name + "/t"] = t
        # cross-check pairs from OpenCV's own crossCheck=True matcher
        if nq > 0 and nt > 0:
            m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
            pairs = np.array(sorted((x.queryIdx, x.trainIdx, int(round(x.distance))) for x in m),
                             dtype=np.int32).reshape(-1, 3)
        else:
            pairs = np.zeros((0, 3), np.int32)
        out[name + "/cross"] = pairs
    out["names"] = np.array(names)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "knn2_bfmatcher.npz"), **out)
    print("knn2_bfmatcher.npz:", len(names), "cases")


# (descriptor kind, nq, nt, seed, mask kind) -- knnMatch(q, t, k=2, mask=M), SURVEY.md section 8(b)/(c)(vii)
MASK_CASES = [
    ("uniform", 7, 3, 31, "half"), ("uniform", 5, 1, 32, "half"), ("uniform", 64, 2, 33, "sparse"),
    ("uniform", 257, 1000, 34, "half"), ("planted", 1000, 1000, 35, "sparse"), ("ties", 257, 1000, 36, "half"),
    ("ties", 130, 4099, 37, "band"), ("dups", 500, 3000, 38, "band"), ("uniform", 33, 4099, 39, "sparse"),
    ("uniform", 100, 700, 40, "ones"), ("uniform", 100, 700, 41, "zeros"), ("planted", 2000, 20000, 42, "sparse"),
    ("ties", 1000, 20000, 43, "half"),
]


def gen_knn2_masked():
    out, names = {}, []
    for kind, nq, nt, seed, mkind in MASK_CASES:
        q, t = make_inputs(kind, nq, nt, seed)
        mask = synth.match_mask(nq, nt, seed + 200000, mkind)
        idx, dist = rows_to_arrays(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2, mask=mask), nq)
        name = f"{kind}_{nq}x{nt}_s{seed}_{mkind}"
        names.append(name)
        out[name + "/idx"] = idx
        out[name + "/dist"] = dist
        out[name + "/sha_q"] = np.array(sha(q))
        out[name + "/sha_t"] = np.array(sha(t))
        out[name + "/sha_mask"] = np.array(sha(mask))
    out["names"] = np.array(names)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "knn2_masked.npz"), **out)
    print("knn2_masked.npz:", len(names), "cases")


def gen_ratio_table():
    """Exhaustive 257x257 table of the reference's float comparison m.distance < r*n.distance
    (tracking.py:27) evaluated the way the reference evaluates it: DMatch.distance is a float32
    holding an integer, the literal 0.7 / 0.75 is a Python double."""
    d = np.arange(257)
    tab = {}
    for name, r in (("r07", 0.7), ("r075", 0.75)):
        t = np.zeros((257, 257), dtype=np.uint8)
        for d1 in d:
            m = cv2.DMatch(0, 0, 0, float(d1)).distance
            for d2 in d:
                n = cv2.DMatch(0, 0, 0, float(d2)).distance
                t[d1, d2] = 1 if m < r * n else 0
        tab[name] = t
    np.savez_compressed(os.path.join(HERE, "ratio_table.npz"), **tab)
    print("ratio_table.npz")


def gen_reference_functions():
    """Run the reference's own helpers with the exhaustive matcher bound at the verified seam."""
    sys.path.insert(0, REFERENCE)
    import tracking  # /root/reference/tracking.py
    import Point3D   # /root/reference/Point3D.py

    flann_ctor = cv2.FlannBasedMatcher
    cv2.FlannBasedMatcher = lambda indexParams=None, searchParams=None: cv2.BFMatcher(cv2.NORM_HAMMING)
    out = {}
    try:
        cases = [("planted", 800, 900, 21), ("planted", 1000, 1000, 22), ("ties", 120, 300, 23),
                 ("planted", 50, 1, 24), ("uniform", 300, 400, 25)]
        names = []
        for kind, nq, nt, seed in cases:
            q, t = make_inputs(kind, nq, nt, seed)
            rng = np.random.default_rng(seed + 7)
            p1 = rng.uniform(0, 1226, size=(nq, 2)).astype(np.float32)
            p2 = rng.uniform(0, 370, size=(nt, 2)).astype(np.float32)
            kp1 = [cv2.KeyPoint(float(x), float(y), 31.0) for x, y in p1]
            kp2 = [cv2.KeyPoint(float(x), float(y), 31.0) for x, y in p2]
            q1, q2 = tracking.get_matches(kp1, q, kp2, t)
            pts3d = rng.normal(0, 400, size=(nq, 3))
            r_q2, r_Q1, r_q1 = Point3D.find_2D_and_3D_correspondenses(q, p1, kp2, t, pts3d, 500)
            name = f"{kind}_{nq}x{nt}_s{seed}"
            names.append(name)
            out[name + "/q"], out[name + "/t"] = q, t
            out[name + "/p1"], out[name + "/p2"], out[name + "/pts3d"] = p1, p2, pts3d
            out[name + "/gm_q1"] = np.asarray(q1, dtype=np.float32)
            out[name + "/gm_q2"] = np.asarray(q2, dtype=np.float32)
            out[name + "/p3_q2"] = np.asarray(r_q2, dtype=np.float64)
            out[name + "/p3_Q1"] = np.asarray(r_Q1, dtype=np.float64)
            out[name + "/p3_q1"] = np.asarray(r_q1, dtype=np.float64)
        out["names"] = np.array(names)
    finally:
        cv2.FlannBasedMatcher = flann_ctor
    np.savez_compressed(os.path.join(HERE, "reference_functions.npz"), **out)
    print("reference_functions.npz:", len(out["names"]), "cases")


def gen_reference_stereo():
    """keypoint.track_keypoints_left_to_right_new (keypoint.py:35-78) run UNMODIFIED up to its fundamental-matrix step:
    cv2.findFundamentalMat is intercepted, the four gathered arrays of keypoint.py:53-57 (pts_left, pts_right, des_left,
    des_right) are read from the reference function's own frame, and the call is aborted there (what follows is RANSAC-like
    sampling and cv2.imshow, which the headless OpenCV build does not have)."""
    sys.path.insert(0, REFERENCE)
    import keypoint  # /root/reference/keypoint.py

    class _Captured(Exception):
        pass

    grabbed = {}

    def fake_fundamental(pts_left, pts_right, method=None, *a, **k):
        loc = sys._getframe(1).f_locals
        for name in ("pts_left", "pts_right", "des_left", "des_right"):
            grabbed[name] = np.array(loc[name])
        raise _Captured()

    flann_ctor, fmat = cv2.FlannBasedMatcher, cv2.findFundamentalMat
    cv2.FlannBasedMatcher = lambda indexParams=None, searchParams=None: cv2.BFMatcher(cv2.NORM_HAMMING)
    cv2.findFundamentalMat = fake_fundamental
    out, names = {}, []
    try:
        for kind, nq, nt, seed in [("planted", 700, 800, 31), ("planted", 1000, 1000, 32), ("ties", 90, 200, 33),
                                   ("planted", 40, 1, 34)]:
            q, t = make_inputs(kind, nq, nt, seed)
            rng = np.random.default_rng(seed + 7)
            p1 = rng.uniform(0, 1226, size=(nq, 2)).astype(np.float32)
            p2 = rng.uniform(0, 370, size=(nt, 2)).astype(np.float32)
            kp1 = [cv2.KeyPoint(float(x), float(y), 31.0) for x, y in p1]
            kp2 = [cv2.KeyPoint(float(x), float(y), 31.0) for x, y in p2]
            grabbed.clear()
            try:
                keypoint.track_keypoints_left_to_right_new(kp1, q, kp2, t, None, None)
            except _Captured:
                pass
            name = f"{kind}_{nq}x{nt}_s{seed}"
            names.append(name)
            out[name + "/q"], out[name + "/t"], out[name + "/p1"], out[name + "/p2"] = q, t, p1, p2
            out[name + "/pts_left"] = np.asarray(grabbed["pts_left"], dtype=np.float64).reshape(-1, 2)
            out[name + "/pts_right"] = np.asarray(grabbed["pts_right"], dtype=np.float64).reshape(-1, 2)
            out[name + "/des_left"] = np.asarray(grabbed["des_left"], dtype=np.uint8).reshape(-1, 32)
            out[name + "/des_right"] = np.asarray(grabbed["des_right"], dtype=np.uint8).reshape(-1, 32)
        out["names"] = np.array(names)
    finally:
        cv2.FlannBasedMatcher, cv2.findFundamentalMat = flann_ctor, fmat
    np.savez_compressed(os.path.join(HERE, "reference_stereo.npz"), **out)
    print("reference_stereo.npz:", len(names), "cases", [out[n + "/pts_left"].shape[0] for n in names])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stereo":
        gen_reference_stereo()
    elif len(sys.argv) > 1 and sys.argv[1] == "masked":
        gen_knn2_masked()
    else:
        gen_knn2()
        gen_knn2_masked()
        gen_ratio_table()
        gen_reference_functions()
        gen_reference_stereo()
