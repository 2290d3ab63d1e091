"""oracle/oracle.py -- CPU restatements of the reference's matching hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module, and only as the checker
or the CPU baseline.  The product (slam-1_b200/) never imports it and has no CPU path.

Three independent statements of the same contract:

  * ``np_*``  -- dependency-free numpy (``np.bitwise_count`` on uint64 views + integer keys).
  * ``c_*``   -- plain C (oracle/hamming_knn2.c), OpenMP over query rows.
  * ``cv_*``  -- ``cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)``: the exhaustive member of
                the OpenCV ``DescriptorMatcher`` family the reference calls
                (tracking.py:17,22  keypoint.py:43,44  Point3D.py:39,40 build the approximate,
                non-deterministic ``FlannBasedMatcher(LSH)``).  OpenCV is an
                un-vendored, un-pinned third-party wheel (requirements.txt:4); probed with
                opencv-python-headless 4.13.0.92.

Pinning: the reference ships no tests or golden vectors.  tests/golden/ holds
vectors generated HERE by tests/golden/make_golden.py from cv2.BFMatcher and from the reference's
own ``tracking.get_matches`` / ``Point3D.find_2D_and_3D_correspondenses`` (imported from
/root/reference with the matcher constructor swapped for the exhaustive one).  All three statements
are checked against those vectors in tests/test_oracle.py.

Conventions: descriptors are ``uint8[n, 32]`` C-contiguous (orb.py:23-24).  Results are
``idx int32[nq, 2]``, ``dist int32[nq, 2]`` ordered by (distance, train index); -1 marks a missing
neighbour (OpenCV returns rows of length min(2, nt)).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NONE_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)
CV_MAX_TRAIN_ROWS = 262143  # OpenCV matchers.cpp:860 packs imgIdx<<18|trainIdx into an int32


# ----------------------------------------------------------------------------------------------
# numpy restatement
# ----------------------------------------------------------------------------------------------
def _as_desc(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8 or a.ndim != 2 or a.shape[1] != 32:
        raise ValueError("descriptors must be uint8[n, 32]")
    return a


def np_distance_matrix(q, t) -> np.ndarray:
    """All-pairs Hamming distance, int32[nq, nt] (small inputs only)."""
    q64 = _as_desc(q).view(np.uint64)
    t64 = _as_desc(t).view(np.uint64)
    return np.bitwise_count(q64[:, None, :] ^ t64[None, :, :]).sum(axis=-1).astype(np.int32)


def np_knn2_keys(q, t, train_index_base: int = 0, chunk: int = 1 << 22) -> np.ndarray:
    """Packed top-2 keys uint64[nq, 2]: (distance << 32) | global train index, NONE_KEY if missing.

    Restates knnMatch(k=2) (tracking.py:22) in exact form: unsigned order of the key is the
    (distance, trainIdx) order OpenCV's brute-force matcher returns.
    """
    q = _as_desc(q)
    t = _as_desc(t)
    nq, nt = q.shape[0], t.shape[0]
    out = np.full((nq, 2), NONE_KEY, dtype=np.uint64)
    if nq == 0 or nt == 0:
        return out
    q64 = q.view(np.uint64)
    t64 = t.view(np.uint64)
    rows = max(1, chunk // max(nt, 1))
    col = (np.arange(nt, dtype=np.uint64) + np.uint64(train_index_base))[None, :]
    for s in range(0, nq, rows):
        qs = q64[s:s + rows]
        d = np.zeros((qs.shape[0], nt), dtype=np.uint64)
        for w in range(4):
            d += np.bitwise_count(qs[:, w:w + 1] ^ t64[None, :, w]).astype(np.uint64)
        key = (d << np.uint64(32)) | col
        if nt == 1:
            out[s:s + rows, 0] = key[:, 0]
        else:
            part = np.partition(key, 1, axis=1)[:, :2]
            part.sort(axis=1)
            out[s:s + rows] = part
    return out


def keys_to_idx_dist(keys: np.ndarray):
    keys = np.asarray(keys, dtype=np.uint64)
    none = keys == NONE_KEY
    idx = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64).astype(np.int32)
    dist = (keys >> np.uint64(32)).astype(np.int64).astype(np.int32)
    idx[none] = -1
    dist[none] = -1
    return idx, dist


def np_knn2(q, t, train_index_base: int = 0):
    return keys_to_idx_dist(np_knn2_keys(q, t, train_index_base))


def np_knn2_masked(q, t, mask, train_index_base: int = 0):
    """knnMatch(q, t, k=2, mask=mask) restated: pair (i, j) takes part iff mask[i, j] != 0 (OpenCV's
    DescriptorMatcher mask); a query with fewer than two allowed rows gets -1 in
    the missing columns (OpenCV returns a short row).  Order (distance, trainIdx) as in np_knn2_keys."""
    q = _as_desc(q)
    t = _as_desc(t)
    nq, nt = q.shape[0], t.shape[0]
    keys = np.full((nq, 2), NONE_KEY, dtype=np.uint64)
    if nq and nt:
        mask = np.asarray(mask)
        if mask.shape != (nq, nt) or mask.dtype != np.uint8:
            raise ValueError("mask must be uint8[nq, nt]")
        q64, t64 = q.view(np.uint64), t.view(np.uint64)
        col = np.arange(nt, dtype=np.uint64) + np.uint64(train_index_base)
        rows = max(1, (1 << 22) // nt)
        for s in range(0, nq, rows):
            qs = q64[s:s + rows]
            d = np.zeros((qs.shape[0], nt), dtype=np.uint64)
            for w in range(4):
                d += np.bitwise_count(qs[:, w:w + 1] ^ t64[None, :, w]).astype(np.uint64)
            key = (d << np.uint64(32)) | col[None, :]
            key[mask[s:s + rows] == 0] = NONE_KEY
            if nt == 1:
                keys[s:s + rows, 0] = key[:, 0]
            else:
                part = np.partition(key, 1, axis=1)[:, :2]
                part.sort(axis=1)
                keys[s:s + rows] = part
    return keys_to_idx_dist(keys)


def np_ratio(dist, num: int, den: int) -> np.ndarray:
    """Integer form of ``m.distance < ratio * n.distance`` (tracking.py:27): den*d1 < num*d2."""
    dist = np.asarray(dist, dtype=np.int64)
    ok = (dist[:, 0] >= 0) & (dist[:, 1] >= 0)
    return (ok & (den * dist[:, 0] < num * dist[:, 1])).astype(np.uint8)


def np_reverse_best(q, t) -> np.ndarray:
    """For every train row the closest query row, lowest query index on ties; int32[nt]."""
    q = _as_desc(q)
    t = _as_desc(t)
    if q.shape[0] == 0:
        return np.full(t.shape[0], -1, dtype=np.int32)
    idx, _ = np_knn2(t, q)
    return idx[:, 0].copy()


def np_cross_check(q, t, idx) -> np.ndarray:
    """Mutual-best test on the best neighbour."""
    idx = np.asarray(idx)
    nq = idx.shape[0]
    acc = np.zeros(nq, dtype=np.uint8)
    if nq == 0 or _as_desc(t).shape[0] == 0:
        return acc
    rb = np_reverse_best(q, t)
    j = idx[:, 0]
    ok = j >= 0
    acc[ok] = (rb[j[ok]] == np.arange(nq, dtype=np.int32)[ok]).astype(np.uint8)
    return acc


def np_merge_top2(keys) -> np.ndarray:
    """keys uint64[n_shards, nq, 2] -> uint64[nq, 2]; plain unsigned order (distance, global idx)."""
    keys = np.asarray(keys, dtype=np.uint64)
    s, nq, _ = keys.shape
    flat = np.transpose(keys, (1, 0, 2)).reshape(nq, 2 * s)
    flat = np.sort(flat, axis=1)
    return np.ascontiguousarray(flat[:, :2])


# ----------------------------------------------------------------------------------------------
# The three reference helpers, restated on top of a kNN-2 result
# ----------------------------------------------------------------------------------------------
def good_rows(idx, dist, num: int = 7, den: int = 10) -> np.ndarray:
    """Query rows the reference's ``for m, n in matches`` loop keeps (tracking.py:24-30).

    The loop unpacks two neighbours per row; the first row with fewer than two raises
    ValueError, which the reference swallows -- silently truncating ``good`` there.
    """
    idx = np.asarray(idx)
    dist = np.asarray(dist)
    short = np.nonzero((idx[:, 1] < 0) | (idx[:, 0] < 0))[0]
    stop = int(short[0]) if short.size else idx.shape[0]
    acc = np_ratio(dist[:stop], num, den).astype(bool)
    return np.nonzero(acc)[0]


def get_matches_restated(kp1_pts, des1, kp2_pts, des2, knn=np_knn2):
    """tracking.py:12-34 with keypoints given as float arrays of .pt coordinates."""
    idx, dist = knn(des1, des2)
    rows = good_rows(idx, dist)
    q1 = np.float32(np.asarray(kp1_pts)[rows]).reshape(-1, 2) if rows.size else np.float32([])
    q2 = np.float32(np.asarray(kp2_pts)[idx[rows, 0]]).reshape(-1, 2) if rows.size else np.float32([])
    return q1, q2


def find_2d_3d_restated(des_i, kp_i, kp_i1_pts, des_i1, pts3d, max_distance=1000, knn=np_knn2):
    """Point3D.py:33-54: ratio 0.7 and |X|,|Y|,|Z| < max_Distance on the query's 3-D point."""
    idx, dist = knn(des_i, des_i1)
    rows = good_rows(idx, dist)
    pts3d = np.asarray(pts3d)
    if rows.size:
        keep = (np.abs(pts3d[rows, 0]) < max_distance) & (np.abs(pts3d[rows, 1]) < max_distance) & \
               (np.abs(pts3d[rows, 2]) < max_distance)
        rows = rows[keep]
    Q1 = np.asarray([pts3d[r] for r in rows])
    q1 = np.asarray([np.asarray(kp_i)[r] for r in rows])
    q2 = np.asarray([np.asarray(kp_i1_pts)[idx[r, 0]] for r in rows])
    return q2, Q1, q1


def stereo_matches_restated(kp_left_pts, des_left, kp_right_pts, des_right, knn=np_knn2):
    """keypoint.py:35-57 up to the fundamental-matrix step: ratio 0.7 and the four gathers (left / right keypoint
    coordinates, left / right descriptors of the good matches)."""
    idx, dist = knn(des_left, des_right)
    rows = good_rows(idx, dist)
    tr = idx[rows, 0]
    return (np.asarray(kp_left_pts)[rows].reshape(-1, 2), np.asarray(kp_right_pts)[tr].reshape(-1, 2),
            np.asarray(des_left)[rows].reshape(-1, 32), np.asarray(des_right)[tr].reshape(-1, 32))


# ----------------------------------------------------------------------------------------------
# C restatement (oracle/hamming_knn2.c)
# ----------------------------------------------------------------------------------------------
_C_LIB = None


def build_c(force: bool = False) -> str:
    """Compile oracle/hamming_knn2.c with oracle/Makefile; returns the .so path."""
    so = os.path.join(_HERE, "liboracle_knn2.so")
    src = os.path.join(_HERE, "hamming_knn2.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_knn2.so"])
    return so


def c_lib():
    global _C_LIB
    if _C_LIB is None:
        so = os.path.join(_HERE, "liboracle_knn2.so")
        if not os.path.exists(so):
            so = build_c()
        lib = ctypes.CDLL(so)
        u8p, i32p, u64p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int32),
                           ctypes.POINTER(ctypes.c_uint64))
        lib.orc_num_threads.restype = ctypes.c_int
        lib.orc_knn2.argtypes = [u8p, ctypes.c_int64, u8p, ctypes.c_int64, ctypes.c_int64, i32p, i32p]
        lib.orc_ratio.argtypes = [i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, u8p]
        lib.orc_reverse_best.argtypes = [u8p, ctypes.c_int64, u8p, ctypes.c_int64, i32p]
        lib.orc_cross_check.argtypes = [u8p, ctypes.c_int64, u8p, ctypes.c_int64, i32p, u8p]
        lib.orc_merge_top2.argtypes = [u64p, ctypes.c_int32, ctypes.c_int64, i32p, i32p]
        for f in (lib.orc_knn2, lib.orc_ratio, lib.orc_reverse_best, lib.orc_cross_check, lib.orc_merge_top2):
            f.restype = None
        _C_LIB = lib
    return _C_LIB


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def c_knn2(q, t, train_index_base: int = 0):
    q = _as_desc(q)
    t = _as_desc(t)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.empty((nq, 2), dtype=np.int32)
    dist = np.empty((nq, 2), dtype=np.int32)
    c_lib().orc_knn2(_p(q, ctypes.c_uint8), nq, _p(t, ctypes.c_uint8), nt, train_index_base,
                     _p(idx, ctypes.c_int32), _p(dist, ctypes.c_int32))
    return idx, dist


def c_ratio(dist, num: int, den: int) -> np.ndarray:
    dist = np.ascontiguousarray(dist, dtype=np.int32)
    acc = np.empty(dist.shape[0], dtype=np.uint8)
    c_lib().orc_ratio(_p(dist, ctypes.c_int32), dist.shape[0], num, den, _p(acc, ctypes.c_uint8))
    return acc


def c_cross_check(q, t, idx) -> np.ndarray:
    q = _as_desc(q)
    t = _as_desc(t)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    acc = np.empty(q.shape[0], dtype=np.uint8)
    c_lib().orc_cross_check(_p(q, ctypes.c_uint8), q.shape[0], _p(t, ctypes.c_uint8), t.shape[0],
                            _p(idx, ctypes.c_int32), _p(acc, ctypes.c_uint8))
    return acc


def c_merge_top2(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    s, nq, _ = keys.shape
    idx = np.empty((nq, 2), dtype=np.int32)
    dist = np.empty((nq, 2), dtype=np.int32)
    c_lib().orc_merge_top2(_p(keys, ctypes.c_uint64), s, nq, _p(idx, ctypes.c_int32), _p(dist, ctypes.c_int32))
    return idx, dist


def c_num_threads() -> int:
    return int(c_lib().orc_num_threads())


# ----------------------------------------------------------------------------------------------
# OpenCV exhaustive matcher (the reference's dependency, same API family)
# ----------------------------------------------------------------------------------------------
def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def dmatch_rows_to_arrays(rows, nq: int):
    """tuple-of-tuples of cv2.DMatch -> (idx, dist) with -1 padding."""
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    for i, row in enumerate(rows):
        for c, m in enumerate(row[:2]):
            idx[i, c] = m.trainIdx
            dist[i, c] = int(round(m.distance))
    return idx, dist


def cv_knn2(q, t, train_index_base: int = 0):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2); train sets >= 2^18 rows are chunked and merged
    by (distance, global index) -- OpenCV's own (distance, imgIdx, trainIdx) collection order."""
    import cv2
    q = _as_desc(q)
    t = _as_desc(t)
    nq, nt = q.shape[0], t.shape[0]
    if nq == 0 or nt == 0:
        return np.full((nq, 2), -1, np.int32), np.full((nq, 2), -1, np.int32)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    if nt <= CV_MAX_TRAIN_ROWS:
        idx, dist = dmatch_rows_to_arrays(bf.knnMatch(q, t, k=2), nq)
        idx[idx >= 0] += train_index_base
        return idx, dist
    parts = []
    for s in range(0, nt, CV_MAX_TRAIN_ROWS):
        i, d = dmatch_rows_to_arrays(bf.knnMatch(q, t[s:s + CV_MAX_TRAIN_ROWS], k=2), nq)
        k = (d.astype(np.uint64) << np.uint64(32)) | (i.astype(np.int64) + s + train_index_base).astype(np.uint64)
        k[i < 0] = NONE_KEY
        parts.append(k)
    return keys_to_idx_dist(np_merge_top2(np.stack(parts)))


def cv_knn2_masked(q, t, mask):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2, mask=mask) as (idx, dist) with -1 padding."""
    import cv2
    q = _as_desc(q)
    t = _as_desc(t)
    nq, nt = q.shape[0], t.shape[0]
    if nq == 0 or nt == 0:
        return np.full((nq, 2), -1, np.int32), np.full((nq, 2), -1, np.int32)
    if nt > CV_MAX_TRAIN_ROWS:
        raise ValueError("cv_knn2_masked: train set too long for one OpenCV call")
    rows = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2, mask=np.ascontiguousarray(mask, dtype=np.uint8))
    return dmatch_rows_to_arrays(rows, nq)


def cv_cross_check_pairs(q, t):
    """BFMatcher(NORM_HAMMING, crossCheck=True).match(q, t) -> sorted list of (queryIdx, trainIdx, dist)."""
    import cv2
    m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(_as_desc(q), _as_desc(t))
    return sorted((x.queryIdx, x.trainIdx, int(round(x.distance))) for x in m)


# ----------------------------------------------------------------------------------------------
# Bag-of-words follow-on (bag_of_words.py:23-42), restated with the reference's own numpy expressions
# ----------------------------------------------------------------------------------------------
def np_bow_hist(labels, n_clusters: int) -> np.ndarray:
    """bag_of_words.py:25: np.histogram(labels, bins=k, range=(0, k - 1))."""
    hist, _ = np.histogram(np.asarray(labels), bins=n_clusters, range=(0, n_clusters - 1))
    return hist


def np_chi2(x, y) -> float:
    """bag_of_words.py:30-31 (and :47-48)."""
    return np.sum(2 * (x - y) ** 2 / (np.maximum(1, x + y)))


def np_predict_previous(h, db, img_index: int, threshold: int):
    """bag_of_words.py:33-42 with the histogram already computed."""
    if img_index < threshold:
        return -1, -1
    dist = []
    for i in range(0, (img_index + 1 - threshold)):
        dist.append(np_chi2(h, db[i]))
    return np.argmin(dist), np.min(dist)


# ----------------------------------------------------------------------------------------------
# Binary vocabulary training (bag_of_words.py:14,20 KMeans.fit, re-specified as k-majority in Hamming
# space: SURVEY.md section 8(f) rank 3).  No reference arithmetic exists for this (the reference clusters
# float vectors with sklearn and cannot be constructed on a current sklearn): the rule
# below IS the specification the CUDA path (csrc/vocab.cu) is held to -- parity unpinned for this row.
# ----------------------------------------------------------------------------------------------
def np_vocab_update(desc, words, vocab):
    """One update step: every word becomes the bitwise majority of its members; an exact tie keeps the old
    bit, a word without members keeps its centroid.  Returns (new vocab uint8[k,32], members int32[k], changed)."""
    desc, vocab = _as_desc(desc), _as_desc(vocab).copy()
    words = np.asarray(words)
    k = vocab.shape[0]
    counts = np.zeros(k, dtype=np.int32)
    changed = 0
    for w in range(k):
        rows = desc[words == w]
        m = rows.shape[0]
        counts[w] = m
        if m == 0:
            continue
        ones = np.unpackbits(rows, axis=1, bitorder="little").astype(np.int64).sum(axis=0)
        old = np.unpackbits(vocab[w], bitorder="little")
        new = np.where(2 * ones > m, 1, np.where(2 * ones == m, old, 0)).astype(np.uint8)
        packed = np.packbits(new, bitorder="little")
        if not np.array_equal(packed, vocab[w]):
            changed += 1
            vocab[w] = packed
    return vocab, counts, changed


def vocab_init(desc, k: int, seed: int):
    """Initial vocabulary: k distinct rows of the pool, chosen by a seeded numpy Generator (shared by the CUDA
    path's host side and by this restatement, so that both start from the same words)."""
    desc = _as_desc(desc)
    if not 1 <= k <= desc.shape[0]:
        raise ValueError("need 1 <= k <= number of descriptors")
    rows = np.random.default_rng(seed).choice(desc.shape[0], size=k, replace=False)
    return desc[np.sort(rows)].copy()


def np_train_vocabulary(desc, k: int, iters: int = 10, seed: int = 0, knn=None):
    """Lloyd iterations in Hamming space: assign (nearest word, lowest index on ties) + np_vocab_update, until no
    centroid changes or `iters` is reached.  Returns (vocab, iterations run)."""
    knn = knn or np_knn2
    vocab = vocab_init(desc, k, seed)
    it = 0
    for it in range(1, iters + 1):
        words = knn(desc, vocab)[0][:, 0]
        vocab, _, changed = np_vocab_update(desc, words, vocab)
        if changed == 0:
            break
    return vocab, it
