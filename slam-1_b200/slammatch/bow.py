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


class BoW:
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
    def _scan(self, hist_dev, n_db: int):
        import torch
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
