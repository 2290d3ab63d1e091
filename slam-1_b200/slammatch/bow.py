"""Bag-of-words place recognition on the GPU (BASELINE config 4's consumer).

Mirror of the reference's ``BoW`` class (bag_of_words.py:10-53) with the same method names and return
values, re-specified as SURVEY.md D6 requires: the reference assigns words with ``sklearn KMeans.predict``
(Euclidean, float centroids -- and cannot even be constructed on a current sklearn: ``n_jobs``), here a word
is the Hamming-nearest row of a BINARY vocabulary ``uint8[k, 32]``, found by the same kNN-2 kernel
(word = best index).  ``hist`` / the chi-square scan / ``(argmin, min)`` follow the reference line by line
and are bit-exact with its numpy arithmetic (tests/test_bow.py).

ORB extraction stays outside (it is image-domain OpenCV code): methods take descriptors, not images.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .matcher import knn2


def vocab_init(descriptors, k: int, seed: int = 0) -> np.ndarray:
    """Initial vocabulary: k distinct rows of the descriptor pool (seeded numpy Generator, rows kept in pool order)."""
    d = np.ascontiguousarray(descriptors)
    if not 1 <= k <= d.shape[0]:
        raise ValueError("need 1 <= k <= number of descriptors")
    rows = np.random.default_rng(seed).choice(d.shape[0], size=k, replace=False)
    return d[np.sort(rows)].copy()


def train_vocabulary(descriptors, k: int, iters: int = 10, seed: int = 0, device: int = 0, init=None):
    """``KMeans(n_clusters=k).fit(dpool)`` of bag_of_words.py:14,20, re-specified for binary descriptors
    (SURVEY.md D6 / section 8(f) rank 3): Lloyd iterations in Hamming space (k-majority).

    Assignment is the kNN kernel itself (nearest word, lowest index on ties -- for the config-4 shape, 1M
    descriptors x 64k words, that is the tensor-pipe kernel); the update is ``slm_vocab_update`` (bitwise majority
    per word, ties keep the old bit, empty words keep their centroid).  Everything stays on the GPU; the only host
    read per iteration is the 4-byte "centroids changed" count.  Returns ``(vocabulary uint8[k, 32], iterations)``.
    """
    import torch
    dev = torch.device("cuda", device)
    if hasattr(descriptors, "is_cuda"):
        d = descriptors.to(dev).contiguous()
        pool = None
    else:
        pool = np.ascontiguousarray(descriptors)
        if pool.dtype != np.uint8 or pool.ndim != 2 or pool.shape[1] != 32:
            raise ValueError("descriptors must be uint8[n, 32]")
        d = torch.from_numpy(pool).to(dev)
    if init is None:
        init = vocab_init(pool if pool is not None else d.cpu().numpy(), k, seed)
    init = np.ascontiguousarray(init)
    if init.shape != (k, 32) or init.dtype != np.uint8:
        raise ValueError("init must be uint8[k, 32]")
    vocab = torch.from_numpy(init).to(dev)
    ctx = _lib.context(device)
    changed = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    it = 0
    for it in range(1, iters + 1):
        idx, _, _ = knn2(d, vocab, ratio=None)
        _lib.check(ctx.lib.slm_vocab_update(ctx.handle, d.data_ptr(), d.shape[0], idx.data_ptr(), 2, vocab.data_ptr(),
                                            k, None, changed.data_ptr(), stream))
        if int(changed.item()) == 0:
            break
    return vocab.cpu().numpy(), it


class BoW:
    @classmethod
    def fit(cls, descriptor_list, n_clusters: int = 50, iters: int = 10, seed: int = 0, device: int = 0):
        """The reference's ``BoW(n_clusters).train(imgs)`` (bag_of_words.py:16-21) from per-image descriptors:
        pool them, train the vocabulary, then store one word histogram per image."""
        pool = np.concatenate([np.ascontiguousarray(d) for d in descriptor_list])
        vocab, _ = train_vocabulary(pool, n_clusters, iters=iters, seed=seed, device=device)
        bow = cls(vocab, device=device, capacity=max(len(descriptor_list), 8))
        bow.train(descriptor_list)
        return bow

    def __init__(self, vocabulary, device: int = 0, capacity: int = 1024):
        import torch
        vocabulary = np.ascontiguousarray(vocabulary)
        if vocabulary.dtype != np.uint8 or vocabulary.ndim != 2 or vocabulary.shape[1] != 32:
            raise ValueError("vocabulary must be uint8[k, 32] (binary visual words)")
        self.n_clusters = int(vocabulary.shape[0])
        self.device = torch.device("cuda", device)
        self._vocab = torch.from_numpy(vocabulary).to(self.device)
        self._db = torch.empty((capacity, self.n_clusters), dtype=torch.int32, device=self.device)
        self._n = 0
        self._ctx = _lib.context(device)

    # -- bag_of_words.py:23-26 ------------------------------------------------------------------------
    def _hist_device(self, descriptors):
        import torch
        d = descriptors if hasattr(descriptors, "is_cuda") else torch.from_numpy(np.ascontiguousarray(descriptors))
        d = d.to(self.device)
        idx, _, _ = knn2(d, self._vocab, ratio=None)            # word = nearest vocabulary row
        hist = torch.empty(self.n_clusters, dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._ctx.lib.slm_bow_hist(self._ctx.handle, idx.data_ptr(), idx.shape[0], 2, self.n_clusters,
                                              hist.data_ptr(), stream))
        return hist

    def hist(self, descriptors) -> np.ndarray:
        """Word histogram of one image's descriptors (int64[k], as np.histogram returns)."""
        return self._hist_device(descriptors).cpu().numpy().astype(np.int64)

    # -- bag_of_words.py:16-21 (the vocabulary is given; the db is one histogram per image) ------------------
    def add(self, descriptors) -> int:
        import torch
        if self._n == self._db.shape[0]:
            grown = torch.empty((2 * self._db.shape[0], self.n_clusters), dtype=torch.int32, device=self.device)
            grown[: self._n] = self._db[: self._n]
            self._db = grown
        self._db[self._n] = self._hist_device(descriptors)
        self._n += 1
        return self._n - 1

    def train(self, descriptor_list):
        self._n = 0
        for d in descriptor_list:
            self.add(d)

    @property
    def db(self) -> np.ndarray:
        return self._db[: self._n].cpu().numpy().astype(np.int64)

    # -- bag_of_words.py:29-53 ------------------------------------------------------------------------
    MAX_SCAN_WORDS = 1 << 19   # slm_chi2_scan's limit (include/slammatch.h); config 4's 65 536-word vocabulary is well inside

    def _scan(self, hist_dev, n_db: int):
        import torch
        if self.n_clusters > self.MAX_SCAN_WORDS:
            raise ValueError(f"predict / predict_previous support vocabularies of at most {self.MAX_SCAN_WORDS} words "
                             f"(this one has {self.n_clusters}); hist() and word assignment have no such limit")
        dist = torch.empty(n_db, dtype=torch.float64, device=self.device)
        best_i = torch.empty(1, dtype=torch.int32, device=self.device)
        best_v = torch.empty(1, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._ctx.lib.slm_chi2_scan(self._ctx.handle, hist_dev.data_ptr(), self._db.data_ptr(), n_db,
                                               self.n_clusters, dist.data_ptr(), best_i.data_ptr(), best_v.data_ptr(),
                                               stream))
        return int(best_i.item()), float(best_v.item()), dist

    def predict_previous(self, descriptors, img_index: int, threshold: int):
        """(argmin, min) of the chi-square distance to db[0 : img_index + 1 - threshold]; (-1, -1) if too early."""
        if img_index < threshold:
            return -1, -1
        n = min(img_index + 1 - threshold, self._n)
        if n <= 0:
            raise ValueError("attempt to get argmin of an empty sequence")     # what np.argmin([]) raises
        i, v, _ = self._scan(self._hist_device(descriptors), n)
        return i, v

    def predict(self, descriptors):
        if self._n == 0:
            raise ValueError("attempt to get argmin of an empty sequence")
        i, v, _ = self._scan(self._hist_device(descriptors), self._n)
        return i, v
