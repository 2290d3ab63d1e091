"""Host-side mirror of the reference's matcher interface (the Python side of the drop-in boundary).

Reference call sites (relative to /root/reference):
    flann = cv2.FlannBasedMatcher(indexParams=..., searchParams=...)   tracking.py:14-17  keypoint.py:40-43  Point3D.py:35-39
    matches = flann.knnMatch(des1, des2, k=2)                          tracking.py:22     keypoint.py:44     Point3D.py:40
    for m, n in matches: if m.distance < 0.7 * n.distance: ...         tracking.py:24-30  keypoint.py:45-51  Point3D.py:41-49

``Matcher`` keeps those names, argument meanings, return layout (tuple of rows of ``DMatch`` with
``queryIdx / trainIdx / imgIdx / distance``; rows have min(k, nt) entries so the reference's swallowed
``ValueError`` on short rows behaves identically) and error behaviour (dtype / column checks raise, empty
inputs give empty rows).  The arithmetic runs in libslammatch.so on a B200; results equal the exhaustive
``cv2.BFMatcher(cv2.NORM_HAMMING)`` bit for bit (SURVEY.md D1 explains why the approximate LSH matcher
itself cannot be an oracle).

``knn2`` is the array fast path (numpy in / numpy out through ``slm_knn2_host``, or torch CUDA tensors
in / out through ``slm_knn2_filter`` with no copies).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

try:  # real cv2.DMatch objects when OpenCV is importable (it is wherever the reference runs)
    import cv2 as _cv2
    _DMatch = _cv2.DMatch
except Exception:  # pragma: no cover - cv2 is present in the supported images
    _cv2 = None

    class _DMatch:  # minimal stand-in with the four fields the reference reads
        __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

        def __init__(self, queryIdx=-1, trainIdx=-1, imgIdx=-1, distance=float("inf")):
            self.queryIdx, self.trainIdx, self.imgIdx, self.distance = queryIdx, trainIdx, imgIdx, float(distance)

DMatch = _DMatch
_fast_rows = None        # (module, payload offset) once verified; False = unavailable


def _init_fast_rows():
    """slammatch._rows builds the DMatch rows in C (what OpenCV's own binding does in C++; ~1 us per object from
    Python otherwise).  It writes the cv::DMatch payload of freshly allocated objects in place, so the layout is
    verified here against a normally constructed object and a small end-to-end sample before first use."""
    global _fast_rows
    _fast_rows = False
    if _cv2 is None:
        return
    try:
        try:
            from . import _rows
        except ImportError:
            from . import build as _build
            _build.build_rows()
            from . import _rows
        import struct
        off = DMatch.__basicsize__ - 16
        probe = DMatch(7, 8, 9, 2.5)
        raw = bytes((ctypes.c_ubyte * 16).from_address(id(probe) + off))
        if struct.unpack("<iiif", raw) != (7, 8, 9, 2.5):
            return
        idx = np.array([[3, -1], [5, 6]], dtype=np.int32)
        dist = np.array([[10, -1], [0, 256]], dtype=np.int32)
        rows = _rows.make_rows(DMatch, off, idx.ctypes.data, dist.ctypes.data, 0, 2, 2, 0, 0)
        got = [[(type(m) is DMatch, m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in r] for r in rows]
        if got == [[(True, 0, 3, 0, 10.0)], [(True, 1, 5, 0, 0.0), (True, 1, 6, 0, 256.0)]]:
            _fast_rows = (_rows, off)
    except Exception:
        _fast_rows = False


REFERENCE_RATIO = (7, 10)   # the literal 0.7 at tracking.py:27, keypoint.py:48, Point3D.py:44


def _as_desc(a, name):
    """uint8[n, 32] C-contiguous, as orb.py:23-24 produces; OpenCV asserts on dtype / column mismatch."""
    if a is None:
        raise ValueError(f"{name} is None")
    a = np.asarray(a)
    if a.size == 0:
        return np.zeros((0, 32), dtype=np.uint8)
    if a.dtype != np.uint8:
        raise ValueError(f"{name}: dtype must be uint8 (ORB descriptors), got {a.dtype}")
    if a.ndim != 2 or a.shape[1] != 32:
        raise ValueError(f"{name}: shape must be [n, 32] (256-bit descriptors), got {a.shape}")
    return np.ascontiguousarray(a)


def _ratio_args(ratio):
    if ratio is None:
        return 0, 1
    num, den = int(ratio[0]), int(ratio[1])
    if num <= 0 or den <= 0:
        raise ValueError("ratio must be a pair of positive integers (num, den)")
    return num, den


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def knn2(q, t, ratio=REFERENCE_RATIO, cross_check: bool = False, device: int | None = None,
         train_index_base: int = 0, variant: str | None = None, mask=None):
    """Exact Hamming 2-NN of every row of ``q`` in ``t`` + Lowe ratio (+ cross-check).

    Returns ``idx int32[nq, 2]``, ``dist int32[nq, 2]``, ``accept uint8[nq]`` -- numpy arrays for numpy
    inputs, CUDA tensors for CUDA-tensor inputs.  ``ratio=(num, den)`` is the integer form of the
    reference's ``m.distance < r * n.distance``; ``None`` disables it.  ``mask`` (optional, ``uint8[nq, nt]``,
    OpenCV's ``knnMatch(..., mask=)``): only pairs with a non-zero entry take part; missing neighbours are -1.
    """
    num, den = _ratio_args(ratio)
    if mask is not None:
        if cross_check:
            # OpenCV: batch_distance.cpp:303 asserts mask.empty() when crossCheck is set
            raise ValueError("mask= cannot be combined with cross-check (as in OpenCV)")
        return _knn2_masked(q, t, mask, num, den, 0 if device is None else device, train_index_base)
    if _is_torch_cuda(q) or _is_torch_cuda(t):
        return _knn2_device(q, t, num, den, cross_check, train_index_base, variant)
    q = _as_desc(q, "queryDescriptors")
    t = _as_desc(t, "trainDescriptors")
    ctx = _lib.context(0 if device is None else device)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    acc = np.zeros(nq, dtype=np.uint8)
    if nq:
        with ctx.using(variant):
            _lib.check(ctx.lib.slm_knn2_host(ctx.handle, q.ctypes.data, nq, t.ctypes.data if nt else None, nt,
                                             num, den, int(bool(cross_check)), idx.ctypes.data, dist.ctypes.data,
                                             acc.ctypes.data))
        if train_index_base:
            idx[idx >= 0] += train_index_base
    return idx, dist, acc


def _dev_desc(x, name):
    import torch
    if not _is_torch_cuda(x):
        raise ValueError(f"{name}: mixing host and device inputs is not supported")
    if x.dtype == torch.uint8 and x.dim() == 2 and x.shape[1] == 32:
        pass
    elif x.dtype == torch.int32 and x.dim() == 2 and x.shape[1] == 8:
        pass  # packed uint32x8 view of the same bytes
    else:
        raise ValueError(f"{name}: expected uint8[n,32] or int32[n,8] CUDA tensor, got {x.dtype} {tuple(x.shape)}")
    return x.contiguous()


def _knn2_device(q, t, num, den, cross_check, base, variant):
    import torch
    q = _dev_desc(q, "queryDescriptors")
    t = _dev_desc(t, "trainDescriptors")
    dev = q.device
    ctx = _lib.context(dev.index or 0)
    nq, nt = q.shape[0], t.shape[0]
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    acc = torch.empty((nq,), dtype=torch.uint8, device=dev)
    if nq:
        stream = torch.cuda.current_stream(dev).cuda_stream
        with ctx.using(variant):
            _lib.check(ctx.lib.slm_knn2_filter(ctx.handle, q.data_ptr(), nq, t.data_ptr() if nt else None, nt, int(base),
                                               num, den, int(bool(cross_check)), idx.data_ptr(), dist.data_ptr(),
                                               acc.data_ptr(), stream))
    return idx, dist, acc


def _knn2_masked(q, t, mask, num, den, device, base):
    """``knnMatch(q, t, k=2, mask=mask)`` through ``slm_knn2_masked`` (one integer-pipe kernel; the mask, one byte per
    pair, is the largest operand).  Host arrays are staged through torch; the result comes back in the inputs' kind."""
    import torch
    on_device = _is_torch_cuda(q) or _is_torch_cuda(t) or _is_torch_cuda(mask)
    if on_device:
        qd, td = _dev_desc(q, "queryDescriptors"), _dev_desc(t, "trainDescriptors")
        if not _is_torch_cuda(mask):
            raise ValueError("mask: mixing host and device inputs is not supported")
        dev = qd.device
    else:
        qh, th = _as_desc(q, "queryDescriptors"), _as_desc(t, "trainDescriptors")
        mask = np.asarray(mask)
    nq = (qd if on_device else qh).shape[0]
    nt = (td if on_device else th).shape[0]
    # OpenCV (matchers.cpp:639): masks[i].type() == CV_8UC1 && rows == queryDescriptorsCount && cols == train rows
    if tuple(mask.shape) != (nq, nt) or mask.dtype != (torch.uint8 if on_device else np.uint8):
        raise ValueError(f"mask: expected uint8[{nq}, {nt}], got {mask.dtype} {tuple(mask.shape)}")
    if not on_device:
        dev = torch.device("cuda", device)
        if not torch.cuda.is_available():
            raise _lib.SlamMatchError(-2, "CUDA device required: there is no CPU path")
        qd, td = torch.from_numpy(qh).to(dev), torch.from_numpy(th).to(dev)
        mask = torch.from_numpy(np.ascontiguousarray(mask)).to(dev)
    md = mask.contiguous()
    ctx = _lib.context(dev.index or 0)
    idx = torch.full((nq, 2), -1, dtype=torch.int32, device=dev)
    dist = torch.full((nq, 2), -1, dtype=torch.int32, device=dev)
    acc = torch.zeros((nq,), dtype=torch.uint8, device=dev)
    if nq:
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(ctx.lib.slm_knn2_masked(ctx.handle, qd.data_ptr(), nq, td.data_ptr() if nt else None, nt, int(base),
                                           md.data_ptr() if nt else None, nt, num, den, idx.data_ptr(), dist.data_ptr(),
                                           acc.data_ptr(), stream))
    if on_device:
        return idx, dist, acc
    return idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy()


def knn2_batched(desc, pairs, ratio=REFERENCE_RATIO, out=None, device: int | None = None, variant: str | None = None):
    """Many frame-to-frame problems in one call (BASELINE config 3: the local-mapping batch -- every keyframe pair of a window
    matched with the arithmetic of tracking.py:22-30): ``desc`` is ``uint8[n_frames, n, 32]``, ``pairs`` is
    ``int32[n_pairs, 2]`` of (query frame, train frame).  Returns ``idx int32[n_pairs, n, 2]`` (train rows local to the train
    frame), ``dist int32[n_pairs, n, 2]``, ``accept uint8[n_pairs, n]``.

    CUDA tensors in -> CUDA tensors out.  Host arrays (numpy, or CPU tensors -- pinned ones are copied asynchronously) in ->
    numpy arrays out; ``out=(idx, dist, accept)`` names preallocated host buffers (numpy arrays or CPU tensors, ideally
    pinned and reused from call to call: a fresh pageable 68 MB result costs more than the search of config 3 itself).
    """
    import torch
    num, den = _ratio_args(ratio)
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32))
    if pairs.ndim != 2 or pairs.shape[1] != 2:
        raise ValueError(f"pairs: expected int32[n_pairs, 2], got {pairs.shape}")
    on_device = _is_torch_cuda(desc)
    d = desc if on_device or isinstance(desc, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(desc))
    if d.dim() != 3 or not ((d.dtype == torch.uint8 and d.shape[2] == 32) or (d.dtype == torch.int32 and d.shape[2] == 8)):
        raise ValueError(f"desc: expected uint8[n_frames, n, 32], got {d.dtype} {tuple(d.shape)}")
    n_frames, n = int(d.shape[0]), int(d.shape[1])
    if pairs.size and (pairs.min() < 0 or pairs.max() >= n_frames):
        raise ValueError("pairs: frame index out of range")
    dev = d.device if on_device else torch.device("cuda", 0 if device is None else device)
    if not on_device and not torch.cuda.is_available():
        raise _lib.SlamMatchError(-2, "CUDA device required: there is no CPU path")
    dd = d.contiguous() if on_device else d.contiguous().to(dev, non_blocking=True)
    ctx = _lib.context(dev.index or 0)
    P = pairs.shape[0]
    idx = torch.empty((P, n, 2), dtype=torch.int32, device=dev)
    dist = torch.empty((P, n, 2), dtype=torch.int32, device=dev)
    acc = torch.empty((P, n), dtype=torch.uint8, device=dev)
    if P and n:
        stream = torch.cuda.current_stream(dev).cuda_stream
        with ctx.using(variant):
            _lib.check(ctx.lib.slm_knn2_batched(ctx.handle, dd.data_ptr(), n_frames, n, pairs.ctypes.data, P, num, den,
                                                idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), stream))
    if on_device:
        return idx, dist, acc
    if out is None:
        return idx.cpu().numpy(), dist.cpu().numpy(), acc.cpu().numpy()
    res = []
    for o, src in zip(out, (idx, dist, acc)):
        ot = o if isinstance(o, torch.Tensor) else torch.from_numpy(o)
        if tuple(ot.shape) != tuple(src.shape) or ot.dtype != src.dtype or not ot.is_contiguous():
            raise ValueError(f"out: expected contiguous {src.dtype} {tuple(src.shape)}, got {ot.dtype} {tuple(ot.shape)}")
        ot.copy_(src, non_blocking=True)
        res.append(ot.numpy())
    torch.cuda.current_stream(dev).synchronize()
    return tuple(res)


class Matcher:
    """Drop-in for the object built by ``cv2.FlannBasedMatcher(indexParams=..., searchParams=...)``.

    The LSH parameters are accepted and ignored (the search is exhaustive and exact).  Only the
    ``DescriptorMatcher`` members on or next to the reference's path are provided: ``knnMatch``,
    ``match``, ``add`` / ``clear`` / ``getTrainDescriptors`` / ``empty``.
    """

    def __init__(self, indexParams=None, searchParams=None, *, crossCheck: bool = False, device: int = 0,
                 variant: str | None = None):
        self.indexParams, self.searchParams = indexParams, searchParams
        self.crossCheck = bool(crossCheck)
        self.device = device
        self.variant = variant
        self._train: list[np.ndarray] = []
        self._db = None     # device-resident copy of the collection, uploaded once in add()

    # -- train collection (OpenCV's add([...]) + knnMatch(q, k) form) --------------------------------
    def add(self, descriptors):
        """``matcher.add([des_a, des_b, ...])``: the collection is uploaded to the GPU once, here, so that
        later ``knnMatch(q, k)`` calls only move the query (see slammatch.keyframe_db.KeyframeDB)."""
        from .keyframe_db import KeyframeDB
        for d in descriptors:
            d = _as_desc(d, "descriptors")
            self._train.append(d)
            if self._db is None:
                self._db = KeyframeDB(self.device)
            self._db.add(d)

    def clear(self):
        self._train = []
        self._db = None

    def empty(self) -> bool:
        return not self._train

    def getTrainDescriptors(self):
        return list(self._train)

    def train(self):  # nothing to build: the search is exhaustive
        return None

    def _collection(self):
        sizes = np.array([d.shape[0] for d in self._train], dtype=np.int64)
        flat = np.concatenate(self._train, axis=0) if self._train else np.zeros((0, 32), np.uint8)
        return flat, np.concatenate([[0], np.cumsum(sizes)])

    # -- the hot path ----------------------------------------------------------------------------
    def knnMatch(self, queryDescriptors, trainDescriptors=None, k: int = 2, mask=None, compactResult: bool = False,
                 masks=None):
        """``matcher.knnMatch(des1, des2, k=2)`` (tracking.py:22): tuple[nq] of tuple[min(k, nt)] of DMatch.

        ``mask`` (``uint8[nq, nt]``; with a collection: ``masks=[uint8[nq, n_i], ...]``, one per added image) restricts the
        search to the pairs with a non-zero entry, as in OpenCV; rows then hold as many neighbours as were allowed (up to k)
        and ``compactResult=True`` drops the empty ones."""
        if isinstance(trainDescriptors, (int, np.integer)) and k == 2:  # knnMatch(q, k) positional form
            trainDescriptors, k = None, int(trainDescriptors)
        if k not in (1, 2):
            raise ValueError("only k=1 and k=2 are supported (the reference uses k=2)")
        if self.crossCheck and k != 1:
            # OpenCV: batch_distance.cpp:303 asserts K == 1 when crossCheck is set
            raise ValueError("crossCheck=True requires k == 1 (as in OpenCV); use knn2(cross_check=True) for kNN-2 + cross-check")
        if mask is not None or masks:
            return self._knn_match_masked(queryDescriptors, trainDescriptors, k, mask, masks, compactResult)
        q = _as_desc(queryDescriptors, "queryDescriptors")
        if trainDescriptors is None and self._db is not None and q.shape[0] > 0:
            # resident collection: only the query crosses PCIe
            idx, dist, acc = self._db.query(q, ratio=None, cross_check=self.crossCheck)
            return self._rows(idx, dist, acc if self.crossCheck else None, k, np.asarray(self._db._offsets))
        if trainDescriptors is None:
            t, offsets = self._collection()
        else:
            t, offsets = _as_desc(trainDescriptors, "trainDescriptors"), None
        idx, dist, acc = knn2(q, t, ratio=None, cross_check=self.crossCheck, device=self.device, variant=self.variant)
        return self._rows(idx, dist, acc if self.crossCheck else None, k, offsets)

    def _knn_match_masked(self, queryDescriptors, trainDescriptors, k, mask, masks, compact):
        if self.crossCheck:
            # OpenCV: batch_distance.cpp:303 asserts mask.empty() when crossCheck is set
            raise ValueError("mask= cannot be combined with crossCheck=True (as in OpenCV)")
        q = _as_desc(queryDescriptors, "queryDescriptors")
        if trainDescriptors is None:
            t, offsets = self._collection()
            if mask is not None and masks is None:
                masks = [mask]
            if len(masks) != len(self._train):
                raise ValueError(f"masks: expected one mask per added image ({len(self._train)}), got {len(masks)}")
            for m, d in zip(masks, self._train):
                if np.asarray(m).shape != (q.shape[0], d.shape[0]):
                    raise ValueError(f"masks: expected uint8[{q.shape[0]}, {d.shape[0]}], got {np.asarray(m).shape}")
            mask = np.concatenate([np.asarray(m) for m in masks], axis=1) if masks else np.zeros((q.shape[0], 0), np.uint8)
        else:
            t, offsets = _as_desc(trainDescriptors, "trainDescriptors"), None
        idx, dist, _ = knn2(q, t, ratio=None, device=self.device, mask=mask)
        rows = self._rows(idx, dist, None, k, offsets)
        return tuple(r for r in rows if r) if compact else rows

    def match(self, queryDescriptors, trainDescriptors=None, mask=None):
        """Best neighbour per query as a flat list (``DescriptorMatcher.match``); honours crossCheck."""
        rows = self.knnMatch(queryDescriptors, trainDescriptors, k=1, mask=mask)
        return [r[0] for r in rows if r]

    @staticmethod
    def _rows(idx, dist, keep, k, offsets):
        if _fast_rows is None:
            _init_fast_rows()
        if _fast_rows:
            mod, off = _fast_rows
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            dist = np.ascontiguousarray(dist, dtype=np.int32)
            keep_a = np.ascontiguousarray(keep, dtype=np.uint8) if keep is not None else None
            img = loc = None
            if offsets is not None:
                offsets = np.asarray(offsets)
                img = (np.searchsorted(offsets, np.maximum(idx, 0), side="right") - 1).astype(np.int32)
                loc = np.ascontiguousarray(idx - offsets[img], dtype=np.int32)
            return mod.make_rows(DMatch, off, idx.ctypes.data, dist.ctypes.data,
                                 keep_a.ctypes.data if keep_a is not None else 0, idx.shape[0], int(k),
                                 img.ctypes.data if img is not None else 0, loc.ctypes.data if loc is not None else 0)
        out = []
        idx_l, dist_l = idx.tolist(), dist.tolist()
        keep_l = keep.tolist() if keep is not None else None
        if offsets is not None:
            img_all = np.searchsorted(offsets, np.maximum(idx, 0), side="right") - 1
            loc_all = (idx - offsets[img_all]).tolist()
            img_l = img_all.tolist()
        for i, (ii, dd) in enumerate(zip(idx_l, dist_l)):
            row = []
            if keep_l is None or keep_l[i]:
                for c in range(k):
                    if ii[c] < 0:
                        break
                    if offsets is None:
                        row.append(DMatch(i, ii[c], 0, float(dd[c])))
                    else:
                        row.append(DMatch(i, loc_all[i][c], img_l[i][c], float(dd[c])))
            out.append(tuple(row))
        return tuple(out)


# ---- the verified seam: rebind cv2.FlannBasedMatcher -------------------------
_saved = {}


def install(cv2_module=None, also_bfmatcher: bool = False):
    """Make the UNMODIFIED reference use this matcher: ``cv2.FlannBasedMatcher = slammatch.Matcher``.

    tracking.get_matches, keypoint.track_keypoints_left_to_right_new and
    Point3D.find_2D_and_3D_correspondenses all construct the matcher through that name on every call.
    """
    mod = cv2_module or _cv2
    if mod is None:
        raise RuntimeError("cv2 is not importable; pass the module the reference uses")
    if "FlannBasedMatcher" not in _saved:
        _saved["FlannBasedMatcher"] = mod.FlannBasedMatcher
    mod.FlannBasedMatcher = Matcher
    if also_bfmatcher:
        if "BFMatcher" not in _saved:
            _saved["BFMatcher"] = mod.BFMatcher

        def _bf(normType=None, crossCheck=False):
            return Matcher(crossCheck=crossCheck)
        mod.BFMatcher = _bf
    return mod


def uninstall(cv2_module=None):
    mod = cv2_module or _cv2
    for name, obj in list(_saved.items()):
        setattr(mod, name, obj)
        del _saved[name]


# ---- array-level mirrors of the three reference helpers (no DMatch objects, no Python loops) ---------
def good_matches(des1, des2, ratio=REFERENCE_RATIO, device: int = 0):
    """The ``good`` list of tracking.py:24-30 as arrays: (queryIdx, trainIdx, distance) of accepted rows,
    truncated at the first row with fewer than two neighbours exactly like the swallowed ValueError."""
    idx, dist, acc = knn2(des1, des2, ratio=ratio, device=device)
    short = np.nonzero(idx[:, 1] < 0)[0]
    stop = int(short[0]) if short.size else idx.shape[0]
    rows = np.nonzero(acc[:stop])[0]
    return rows, idx[rows, 0], dist[rows, 0]


def get_matches(kp1_pts, des1, kp2_pts, des2, device: int = 0):
    """tracking.get_matches (tracking.py:12-34) with keypoint coordinates given as float arrays."""
    rows, tidx, _ = good_matches(des1, des2, device=device)
    q1 = np.float32(np.asarray(kp1_pts)[rows])
    q2 = np.float32(np.asarray(kp2_pts)[tidx])
    return q1, q2


def _compact_and_gather(idx, dist, acc, gathers):
    """Device-side tail shared by the three reference helpers: ordered compaction of the accepted rows (with the
    truncation at the first short row, tracking.py:25-30) + one slm_gather_rows per ``(column, tensor)`` in ``gathers``
    (column 0 gathers by queryIdx, 1 by trainIdx).  One 4-byte read-back (the match count) is the only host sync."""
    import torch
    dev, nq = idx.device, idx.shape[0]
    ctx = _lib.context(dev.index or 0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    matches = torch.empty((max(nq, 1), 3), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(ctx.lib.slm_compact_matches(ctx.handle, idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), nq, 1,
                                           matches.data_ptr(), count.data_ptr(), stream))
    out = []
    for col, src in gathers:
        src = src.contiguous()
        if src.dim() != 2 or not src.is_cuda:
            raise ValueError("gather sources must be 2-D CUDA tensors")
        row_bytes = src.shape[1] * src.element_size()
        if row_bytes % 4:
            raise ValueError("gather rows must be a multiple of 4 bytes")
        o = torch.empty((max(nq, 1), src.shape[1]), dtype=src.dtype, device=dev)
        _lib.check(ctx.lib.slm_gather_rows(ctx.handle, src.data_ptr(), row_bytes, matches.data_ptr(), count.data_ptr(), nq,
                                           col, o.data_ptr(), stream))
        out.append(o)
    m = int(count.item())
    return [o[:m] for o in out]


def get_matches_device(kp1_xy, des1, kp2_xy, des2, ratio=REFERENCE_RATIO):
    """tracking.get_matches (tracking.py:12-34) with everything device-resident: descriptors ``uint8[n,32]`` and
    keypoint coordinates ``float32[n,2]`` are CUDA tensors; the search, the ratio loop (with its truncation at the
    first short row), the compaction of ``good`` and both gathers run on the GPU.  Returns ``(q1, q2)`` CUDA
    tensors ``float32[M,2]``."""
    import torch
    for xy in (kp1_xy, kp2_xy):
        if xy.dtype != torch.float32 or xy.dim() != 2:
            raise ValueError("keypoint coordinates must be float32[n, k] CUDA tensors")
    idx, dist, acc = knn2(des1, des2, ratio=ratio)
    q1, q2 = _compact_and_gather(idx, dist, acc, [(0, kp1_xy), (1, kp2_xy)])
    return q1, q2


def find_2d_3d_device(des_i, kp_i_xy, kp_i1_xy, des_i1, pts3d, max_distance=1000, ratio=REFERENCE_RATIO):
    """Point3D.find_2D_and_3D_correspondenses (Point3D.py:33-54) device-resident: kNN-2 + ratio 0.7, then a match is kept
    only if the query's triangulated point has |X|, |Y|, |Z| < max_Distance (Point3D.py:45-46; slm_filter_points3d), then
    the three gathers of Point3D.py:50-52.  ``pts3d`` is ``float64[nq,3]``.  Returns ``(q2, Q1, q1)`` like the reference:
    train keypoints of the kept matches, their 3-D points, their query keypoints."""
    import torch
    if pts3d.dtype != torch.float64 or pts3d.dim() != 2 or pts3d.shape[1] != 3:
        raise ValueError("pts3d must be a float64[nq, 3] CUDA tensor")
    idx, dist, acc = knn2(des_i, des_i1, ratio=ratio)
    dev, nq = idx.device, idx.shape[0]
    if pts3d.shape[0] != nq:
        raise ValueError("one 3-D point per query descriptor")
    if nq:
        ctx = _lib.context(dev.index or 0)
        p = pts3d.contiguous()
        _lib.check(ctx.lib.slm_filter_points3d(ctx.handle, p.data_ptr(), nq, float(max_distance), acc.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
    q2, Q1, q1 = _compact_and_gather(idx, dist, acc, [(1, kp_i1_xy), (0, pts3d), (0, kp_i_xy)])
    return q2, Q1, q1


def stereo_matches_device(kp_left_xy, des_left, kp_right_xy, des_right, ratio=REFERENCE_RATIO):
    """keypoint.track_keypoints_left_to_right_new (keypoint.py:35-57) up to its fundamental-matrix step, device-resident:
    kNN-2 + ratio 0.7 and the four gathers -- left / right keypoint coordinates and left / right DESCRIPTORS of the good
    matches (keypoint.py:53-57).  Returns ``(pts_left, pts_right, des_left_good, des_right_good)`` CUDA tensors."""
    idx, dist, acc = knn2(des_left, des_right, ratio=ratio)
    dl = _dev_desc(des_left, "descriptors_left")
    dr = _dev_desc(des_right, "descriptors_right")
    return tuple(_compact_and_gather(idx, dist, acc, [(0, kp_left_xy), (1, kp_right_xy), (0, dl), (1, dr)]))
