"""pytest configuration: registers the ``gpu`` marker and puts the package on sys.path.

``-m "not gpu"`` = oracle vs golden vectors, host logic, C-ABI symbol checks (runs on the CPU
container); ``-m gpu`` = parity tests proper, through the C-ABI on a real B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "slam-1_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_inputs(kind, nq, nt, seed):
    """Regenerate a golden case's inputs from its seed (same code as tests/golden/make_golden.py)."""
    from slammatch import synth
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


def load_knn2_golden():
    """Yield (name, q, t, idx, dist, cross_pairs) for every committed BFMatcher vector."""
    import hashlib
    z = np.load(os.path.join(GOLDEN, "knn2_bfmatcher.npz"))
    out = []
    for name in z["names"]:
        name = str(name)
        kind, shape, seed = name.split("_")
        nq, nt = (int(x) for x in shape.split("x"))
        seed = int(seed[1:])
        if name + "/q" in z.files:
            q, t = z[name + "/q"], z[name + "/t"]
        else:
            q, t = golden_inputs(kind, nq, nt, seed)
        # the inputs must be byte-identical to what the vector was generated from
        assert hashlib.sha256(np.ascontiguousarray(q).tobytes()).hexdigest() == str(z[name + "/sha_q"]), name
        assert hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest() == str(z[name + "/sha_t"]), name
        out.append((name, q, t, z[name + "/idx"], z[name + "/dist"], z[name + "/cross"]))
    return out


@pytest.fixture(scope="session")
def knn2_golden():
    return load_knn2_golden()


def load_masked_golden():
    """Yield (name, q, t, mask, idx, dist) for every committed knnMatch(..., mask=) vector of cv2.BFMatcher
    (tests/golden/knn2_masked.npz); inputs and masks are regenerated from their seeds and checked by hash."""
    import hashlib
    from slammatch import synth
    z = np.load(os.path.join(GOLDEN, "knn2_masked.npz"))
    out = []
    for name in z["names"]:
        name = str(name)
        kind, shape, seed, mkind = name.split("_")
        nq, nt = (int(x) for x in shape.split("x"))
        seed = int(seed[1:])
        q, t = golden_inputs(kind, nq, nt, seed)
        mask = synth.match_mask(nq, nt, seed + 200000, mkind)
        for a, key in ((q, "sha_q"), (t, "sha_t"), (mask, "sha_mask")):
            assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == str(z[f"{name}/{key}"]), (name, key)
        out.append((name, q, t, mask, z[name + "/idx"], z[name + "/dist"]))
    return out


@pytest.fixture(scope="session")
def masked_golden():
    return load_masked_golden()
