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
