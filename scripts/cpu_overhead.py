"""CPU-side cost per call of the public device path (tiny problem, GPU work negligible)."""
import sys, time
sys.path.insert(0, "slam-1_b200"); sys.path.insert(0, ".")
import torch, numpy as np, slammatch
from slammatch import synth, _lib
q = torch.from_numpy(synth.uniform(64, 1)).cuda(); t = torch.from_numpy(synth.uniform(2048, 2)).cuda()
ctx = _lib.context(0)
def bench(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("knn2 device path (64x2048)      %.1f us/call" % bench(lambda: slammatch.knn2(q, t)))
ctx.profile(True)
print("  with profiling marks           %.1f us/call" % bench(lambda: slammatch.knn2(q, t), 500)); ctx.profile_read(); ctx.profile(False)
idx = torch.empty((64, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx); acc = torch.empty(64, dtype=torch.uint8, device="cuda")
def raw():
    _lib.check(ctx.lib.slm_knn2_filter(ctx.handle, q.data_ptr(), 64, t.data_ptr(), 2048, 0, 7, 10, 0, idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), None))
print("raw ctypes slm_knn2_filter       %.1f us/call" % bench(raw))
qh, th = synth.uniform(64, 1), synth.uniform(2048, 2)
print("host path knn2(numpy)            %.1f us/call" % bench(lambda: slammatch.knn2(qh, th)))
m = slammatch.Matcher()
print("Matcher.knnMatch (DMatch rows)   %.1f us/call" % bench(lambda: m.knnMatch(qh, th, k=2), 500))
q2 = torch.from_numpy(synth.uniform(1000, 3)).cuda(); t2 = torch.from_numpy(synth.uniform(1000, 4)).cuda()
print("knn2 device path (1000x1000)     %.1f us/call" % bench(lambda: slammatch.knn2(q2, t2)))
# ---- the reference's own shape (c1: 1000 x 1000, host arrays in / out) broken down -------------------------------
qh, th = synth.planted(1000, 1000, 5)
print("c1 host path knn2(numpy)         %.1f us/call" % bench(lambda: slammatch.knn2(qh, th, ratio=(3, 4))))
qp, tp = torch.from_numpy(qh).pin_memory().numpy(), torch.from_numpy(th).pin_memory().numpy()
print("c1 host path, pinned inputs      %.1f us/call" % bench(lambda: slammatch.knn2(qp, tp, ratio=(3, 4))))
i_h = np.empty((1000, 2), np.int32); d_h = np.empty((1000, 2), np.int32); a_h = np.empty(1000, np.uint8)
ptrs = (qh.ctypes.data, th.ctypes.data, i_h.ctypes.data, d_h.ctypes.data, a_h.ctypes.data)
def raw_host():
    _lib.check(ctx.lib.slm_knn2_host(ctx.handle, ptrs[0], 1000, ptrs[1], 1000, 3, 4, 0, ptrs[2], ptrs[3], ptrs[4]))
print("c1 raw ctypes slm_knn2_host      %.1f us/call" % bench(raw_host))
m = slammatch.Matcher()
print("c1 Matcher.knnMatch (DMatch)     %.1f us/call" % bench(lambda: m.knnMatch(qh, th, k=2), 300))
try:
    import cv2
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    t0 = time.perf_counter()
    for _ in range(20): bf.knnMatch(qh, th, k=2)
    print("c1 cv2.BFMatcher.knnMatch (CPU)  %.1f us/call" % ((time.perf_counter() - t0) / 20 * 1e6))
    fl = cv2.FlannBasedMatcher(indexParams=dict(algorithm=6, table_number=6, key_size=12, multi_probe_level=1), searchParams=dict(checks=50))
    t0 = time.perf_counter()
    for _ in range(20): fl.knnMatch(qh, th, k=2)
    print("c1 cv2 FLANN-LSH knnMatch (CPU, the reference as shipped; approximate)  %.1f us/call" % ((time.perf_counter() - t0) / 20 * 1e6))
except Exception as e:
    print("cv2 unavailable:", e)
