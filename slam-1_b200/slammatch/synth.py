"""Synthetic ORB-descriptor generators.

The reference produces descriptors with cv2.ORB (orb.py:4-38 -> ``uint8[N, 32]`` C-contiguous,
orb.py:23-24).  Neither KITTI images nor a network are available, so tests and bench.py use the
seeded generators below; every generator is a pure function of its arguments.
"""
from __future__ import annotations

import numpy as np


def uniform(n: int, seed: int) -> np.ndarray:
    """Uniform random bytes: distances ~ Binomial(256, 1/2) (mean 128, sigma 8 => frequent ties)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def planted(nq: int, nt: int, seed: int, frac: float = 0.5, flip: float = 0.06):
    """Half of the queries are a distinct train row with ~6 % of the bits flipped, the rest uniform,
    so the Lowe ratio test accepts a known, non-trivial fraction (uniform data accepts ~0 rows)."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    n_pl = min(int(nq * frac), nt)
    if n_pl > 0:
        rows = rng.choice(nq, size=n_pl, replace=False)
        src = rng.choice(nt, size=n_pl, replace=False)
        noise_bits = rng.random((n_pl, 256)) < flip
        noise = np.packbits(noise_bits, axis=1, bitorder="little")
        q[rows] = t[src] ^ noise
    return q, t


def heavy_ties(n: int, seed: int, live_bits: int = 3) -> np.ndarray:
    """Descriptors with only a few live bits => distances in 0..2*live_bits, massive ties;
    pins the lowest-index tie-break for best and second best."""
    rng = np.random.default_rng(seed)
    bits = np.zeros((n, 256), dtype=bool)
    pos = rng.integers(0, 16, size=(n, live_bits))  # confined to 16 positions => many exact repeats
    on = rng.random((n, live_bits)) < 0.7
    for k in range(live_bits):
        bits[np.arange(n), pos[:, k]] |= on[:, k]
    return np.packbits(bits, axis=1, bitorder="little")


def with_duplicates(t: np.ndarray, seed: int, frac: float = 0.2) -> np.ndarray:
    """Overwrite a fraction of rows with copies of other rows (exact duplicates within / across shards)."""
    rng = np.random.default_rng(seed)
    t = t.copy()
    n = t.shape[0]
    k = int(n * frac)
    if n >= 2 and k > 0:
        dst = rng.choice(n, size=k, replace=False)
        src = rng.integers(0, n, size=k)
        t[dst] = t[src]
    return t


def match_mask(nq: int, nt: int, seed: int, kind: str = "half") -> np.ndarray:
    """Masks for ``knnMatch(..., mask=)``: uint8[nq, nt], pair (i, j) allowed iff the entry is non-zero.

    ``half``: every pair allowed with probability 1/2, values drawn from {1, 7, 255} (any non-zero byte counts);
    ``sparse``: ~3 allowed rows per query, and every fifth query has none / exactly one / exactly two allowed rows in turn
    (the short-row cases); ``band``: only |i * nt / nq - j| <= 8 (a spatial search window); ``ones`` / ``zeros``.
    """
    rng = np.random.default_rng(seed)
    if kind == "ones":
        return np.ones((nq, nt), np.uint8)
    if kind == "zeros":
        return np.zeros((nq, nt), np.uint8)
    if kind == "half":
        m = rng.random((nq, nt)) < 0.5
        return (m * rng.choice(np.array([1, 7, 255], np.uint8), size=(nq, nt))).astype(np.uint8)
    if kind == "sparse":
        m = (rng.random((nq, nt)) < min(1.0, 3.0 / max(nt, 1))).astype(np.uint8)
        for i in range(0, nq, 5):
            m[i] = 0
            n_allowed = (i // 5) % 3
            if n_allowed and nt:
                m[i, rng.choice(nt, size=min(n_allowed, nt), replace=False)] = 1
        return m
    if kind == "band":
        centre = (np.arange(nq, dtype=np.int64) * max(nt, 1)) // max(nq, 1)
        return (np.abs(centre[:, None] - np.arange(nt, dtype=np.int64)[None, :]) <= 8).astype(np.uint8)
    raise ValueError(kind)
