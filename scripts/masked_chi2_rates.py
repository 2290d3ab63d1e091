"""Rates of the two round-2 side kernels on one GPU: the masked search (knn2_masked_kernel; Gcmp/s over ALL pairs and over the
allowed ones, against the in-process POPC ceiling) and the wide chi-square scan (chi2_scan_wide_kernel; GB/s of stored
histograms against the measured HBM bandwidth).  usage: masked_chi2_rates.py"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

import slammatch
from slammatch import _lib, synth

ctx = _lib.context(0)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
except Exception:
    pass
popc_tcmp, lanes = ctx.probe_popc_peak(3)
print(f"in-process POPC ceiling: {popc_tcmp * 1000:.0f} Gcmp/s ({lanes:.1f} POPC lanes/clk/SM)")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


MASKED = () if "--chi2-only" in sys.argv else ((2000, 20000, "half"), (2000, 20000, "band"), (2000, 20000, "ones"), (1000, 1000, "half"),
                                                (2000, 200000, "half"))
for nq, nt, kind in MASKED:
    q, t = synth.planted(nq, nt, 5)
    m = synth.match_mask(nq, nt, 6, kind)
    qd, td, md = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(m).cuda()
    ms = timed(lambda: slammatch.knn2(qd, td, ratio=(7, 10), mask=md))
    ms_plain = timed(lambda: slammatch.knn2(qd, td, ratio=(7, 10)))
    allowed = float((m != 0).mean())
    print(f"masked {nq} x {nt} mask={kind:6s} allowed {allowed:.3f}: {ms * 1e3:8.1f} us  {nq * nt / ms / 1e6:8.1f} Gcmp/s over all pairs "
          f"({nq * nt / ms / 1e6 / (popc_tcmp * 1000):.2f} of the POPC ceiling; mask stream {nq * nt / ms / 1e6:.0f} GB/s); "
          f"unmasked call {ms_plain * 1e3:.1f} us")

hbm = float(peaks.get("hbm_gbs") or 0.0)       # measured copy bandwidth (read + write bytes)
rng = np.random.default_rng(3)


def chi2_case(k, n_db, kind):
    """dense: every count uniform in 0..3 (about 70 % of the terms need the float64 division -- the worst case for the scan);
    sparse: every stored histogram is the word histogram of 2000 descriptors (what BoW.hist produces for a 64k-word
    vocabulary: >= 97 % of the words are empty in both histograms and their terms are +0.0 without a division)."""
    if kind == "dense":
        db = torch.from_numpy(rng.integers(0, 4, (n_db, k)).astype(np.int32)).cuda()
    else:
        g = torch.Generator(device="cuda"); g.manual_seed(4)
        words = torch.randint(0, k, (n_db, 2000), device="cuda", generator=g)
        db = torch.zeros((n_db, k), dtype=torch.int32, device="cuda")
        db.scatter_add_(1, words, torch.ones_like(words, dtype=torch.int32))
    h = db[7].clone()
    dist = torch.empty(n_db, dtype=torch.float64, device="cuda")
    bi = torch.empty(1, dtype=torch.int32, device="cuda")
    bv = torch.empty(1, dtype=torch.float64, device="cuda")
    ms = timed(lambda: _lib.check(ctx.lib.slm_chi2_scan(ctx.handle, h.data_ptr(), db.data_ptr(), n_db, k, dist.data_ptr(),
                                                        bi.data_ptr(), bv.data_ptr(), None)), n=10)
    # spot check against the reference's numpy expression (bag_of_words.py:30-31), bit for bit
    hn = h.cpu().numpy().astype(np.int64)
    for i in (0, 7, n_db - 1):
        y = db[i].cpu().numpy().astype(np.int64)
        want = np.sum(2 * (hn - y) ** 2 / np.maximum(1, hn + y))
        assert float(dist[i].item()) == float(want), (k, n_db, kind, i)
    gbs = n_db * k * 4 / ms / 1e6
    print(f"chi2 scan {n_db} stored histograms x {k} words, {kind:6s}: {ms * 1e3:8.1f} us  {gbs:8.1f} GB/s"
          + (f" = {gbs / hbm:.2f} of the measured HBM copy bandwidth ({hbm:.0f} GB/s)" if hbm else "") + f"  argmin {int(bi.item())}")


for k, n_db, kind in ((50, 100000, "dense"), (1024, 20000, "dense"), (4096, 8000, "dense"), (65536, 2000, "dense"),
                      (65536, 2000, "sparse"), (65536, 8000, "sparse"), (4096, 20000, "sparse")):
    chi2_case(k, n_db, kind)
