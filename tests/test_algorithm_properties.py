"""CPU checks (numpy, oracle only) of the two algorithmic shortcuts the CUDA path relies on -- so that a GPU
parity failure can be told apart from a flaw in the idea:

1. Candidate chunks (csrc/knn2_tc.cu): the exact top-2 train rows by (distance, index) always lie inside the
   best two W-column chunks ordered by (chunk minimum distance asc, chunk index asc), for any chunk width W
   (32 for single problems, 16 in chained batches).
2. Reduced reverse search (csrc/api.cu, cross-check): query i is mutual iff the nearest QUERY to train row
   best(i) -- lowest query index on ties -- is i; searching only the rows best(i) gives the same verdicts as
   the full swapped search, including under duplicated train rows and duplicated queries.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from slammatch import synth


@pytest.mark.parametrize("width", [16, 32])
@pytest.mark.parametrize("kind", ["uniform", "ties", "dups"])
def test_top2_rows_lie_in_the_best_two_chunks(width, kind):
    rng = np.random.default_rng(width + len(kind))
    for nq, nt in ((40, 1000), (17, 257), (8, 33), (5, 16), (3, 1)):
        if kind == "ties":
            q, t = synth.heavy_ties(nq, 1 + nt), synth.heavy_ties(nt, 2 + nt)
        else:
            q, t = synth.planted(nq, nt, 3 + nt)
            if kind == "dups":
                t = synth.with_duplicates(t, 4, 0.5)
        d = orc.np_distance_matrix(q, t)                                   # [nq, nt]
        n_chunks = -(-nt // width)
        pad = np.full((nq, n_chunks * width), 10_000, dtype=np.int64)
        pad[:, :nt] = d
        cmin = pad.reshape(nq, n_chunks, width).min(axis=2)                # chunk minimum = what the epilogue's max dot encodes
        order = np.lexsort((np.broadcast_to(np.arange(n_chunks), cmin.shape), cmin), axis=1)   # (min asc, chunk asc)
        best2 = order[:, :2]
        idx, _ = orc.np_knn2(q, t)
        for i in range(nq):
            for c in range(min(2, nt)):
                assert idx[i, c] // width in best2[i, :min(2, n_chunks)], (kind, width, nq, nt, i, c)


def _chunk_bounds(nt, geom, base=0):
    """Chunk boundaries of a row block that starts at global row `base`: uniform widths (the kernels use 120 / 40 / 32 / 16), or
    non-uniform widths per 240-row tile (two chunks 128 + 112, or six chunks 40 40 48 40 40 32 -- the geometry of a measured
    kernel variant; the property does not depend on the widths)."""
    widths = {"wide": [128, 112], "narrow": [40, 40, 48, 40, 40, 32]}.get(geom) or [int(geom)]
    out, r, k = [], 0, 0
    while r < nt:
        w = widths[k % len(widths)]
        out.append((base + r, base + min(r + w, nt)))
        r += w
        k += 1
    return out


@pytest.mark.parametrize("geom", ["120", "40", "wide", "narrow"])
@pytest.mark.parametrize("shards", [1, 3, 8])
def test_top2_rows_lie_in_the_global_best_two_chunks_of_any_sharding(geom, shards):
    """Any chunk geometry, and the two-phase sharded refine (csrc/knn2_tc.cu
    tc_chunk_keys_kernel / tc_refine_owned_kernel): every shard chunks ITS OWN row block from its first row, every rank
    contributes its best two chunks per query, and the exact top-2 rows lie in the best two chunks of the union by
    (chunk minimum distance asc, first global row asc) -- so only the owners of those two chunks need to re-score."""
    for nq, nt, seed in ((30, 2000, 1), (12, 700, 2), (9, 241, 3), (4, 50, 4)):
        q, t = synth.planted(nq, nt, seed)
        t = synth.with_duplicates(t, seed + 9, 0.4)
        d = orc.np_distance_matrix(q, t)
        idx, _ = orc.np_knn2(q, t)
        cuts = np.linspace(0, nt, shards + 1).astype(int)
        for i in range(nq):
            contributed = []
            for a, b in zip(cuts[:-1], cuts[1:]):
                chunks = [(d[i, lo:hi].min(), lo, hi) for lo, hi in _chunk_bounds(b - a, geom, a) if hi > lo]
                contributed += sorted(chunks)[:2]                          # the rank's best two by (min distance, first row)
            best2 = sorted(contributed)[:2]
            for c in range(min(2, nt)):
                assert any(lo <= idx[i, c] < hi for _, lo, hi in best2), (geom, shards, nq, nt, i, c)


def test_reduced_reverse_search_gives_the_full_cross_check():
    for nq, nt, seed, kind in ((200, 900, 1, "planted"), (64, 64, 2, "ties"), (300, 40, 3, "planted"),
                               (50, 500, 4, "dupq"), (1, 7, 5, "planted"), (9, 1, 6, "planted")):
        if kind == "ties":
            q, t = synth.heavy_ties(nq, seed), synth.heavy_ties(nt, seed + 50)
        else:
            q, t = synth.planted(nq, nt, seed)
            t = synth.with_duplicates(t, seed, 0.4)
            if kind == "dupq":
                q[1::5] = q[0]
        idx, _ = orc.np_knn2(q, t)
        full = orc.np_cross_check(q, t, idx)                                # full swapped search (oracle definition)
        best_rows = t[np.maximum(idx[:, 0], 0)]
        rev, _ = orc.np_knn2(best_rows, q)                                  # best(i) x all queries
        reduced = ((idx[:, 0] >= 0) & (rev[:, 0] == np.arange(nq))).astype(np.uint8)
        assert np.array_equal(reduced, full), (nq, nt, kind)
        assert np.array_equal(full, orc.c_cross_check(q, t, idx))
