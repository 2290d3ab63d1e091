"""One-shot check of the chi-square scan after a kernel change, sized for a few seconds of GPU time: the chi-square tests of
tests/test_bow.py called directly (same shapes, same seeds as the suite), then four rates.  Every line is flushed to
gpurun_out/chi2_quick.txt as soon as it is known.  usage: chi2_quick_check.py [--dry]"""
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "slam-1_b200")]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "chi2_quick.txt"), "w")
T0 = time.perf_counter()


def say(msg):
    line = f"[{time.perf_counter() - T0:6.2f} s] {msg}"
    print(line, flush=True)
    OUT.write(line + "\n")
    OUT.flush()
    os.fsync(OUT.fileno())


import numpy as np  # noqa: E402
import torch  # noqa: E402
from slammatch import _lib  # noqa: E402
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("test_bow_direct", os.path.join(ROOT, "tests", "test_bow.py"))
tb = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tb)

say("imports done")
if "--dry" in sys.argv:
    sys.exit(0)
ctx = _lib.context(0)
for name in ("test_chi2_scan_first_minimum_wins_and_sum_order_is_numpys", "test_chi2_term_count_ranges_zero_one_and_beyond_16_bits",
             "test_chi2_scan_wide_vocabularies_are_bit_exact_too"):
    getattr(tb, name)()
    say(f"PASS {name}")


def rate(k, n_db, kind):
    if kind == "dense":
        db = torch.randint(0, 4, (n_db, k), device="cuda", dtype=torch.int32)
    else:
        words = torch.randint(0, k, (n_db, 2000), device="cuda")
        db = torch.zeros((n_db, k), dtype=torch.int32, device="cuda")
        db.scatter_add_(1, words, torch.ones_like(words, dtype=torch.int32))
    h = db[7].clone()
    dist = torch.empty(n_db, dtype=torch.float64, device="cuda")
    bi = torch.empty(1, dtype=torch.int32, device="cuda")
    bv = torch.empty(1, dtype=torch.float64, device="cuda")

    def fn():
        _lib.check(ctx.lib.slm_chi2_scan(ctx.handle, h.data_ptr(), db.data_ptr(), n_db, k, dist.data_ptr(), bi.data_ptr(),
                                         bv.data_ptr(), None))
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    hn = h.cpu().numpy().astype(np.int64)
    for i in (0, 7, n_db - 1):
        y = db[i].cpu().numpy().astype(np.int64)
        assert float(dist[i].item()) == float(np.sum(2 * (hn - y) ** 2 / np.maximum(1, hn + y))), (k, n_db, kind, i)
    say(f"chi2 scan {n_db} x {k} words, {kind:6s}: {ms * 1e3:8.1f} us  {n_db * k * 4 / ms / 1e6:8.1f} GB/s  (6544 GB/s = measured HBM copy)")


for k, n_db, kind in ((65536, 2000, "sparse"), (65536, 8000, "sparse"), (65536, 2000, "dense"), (50, 100000, "dense"),
                      (4096, 8000, "dense"), (1024, 20000, "dense")):
    rate(k, n_db, kind)
for name in ("test_bow_predict_with_a_64k_word_vocabulary", "test_bow_hist_and_scan_match_reference_arithmetic"):
    getattr(tb, name)()
    say(f"PASS {name}")
say("ALL DONE")
