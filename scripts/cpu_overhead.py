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
