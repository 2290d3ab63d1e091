"""Kernel time vs train size at nq=2000 (1 GPU): the intercept is the fixed per-launch cost of the tcgen05 kernel."""
import sys, time
sys.path.insert(0, "slam-1_b200"); sys.path.insert(0, ".")
import torch, numpy as np, slammatch
from slammatch import synth, _lib
ctx = _lib.context(0)
q = torch.from_numpy(synth.uniform(2000, 1)).cuda()
T = torch.from_numpy(synth.uniform(10_000_000, 2)).cuda()
for nt in (39_168, 156_250, 625_000, 1_250_000, 2_500_000, 5_000_000, 10_000_000):
    t = T[:nt]
    for _ in range(5): slammatch.knn2(q, t)
    torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n): slammatch.knn2(q, t)
    e1.record(); torch.cuda.synchronize()
    ms, k = ctx.profile_read(); ctx.profile(False)
    tiles = (nt + 255) // 256
    print(f"nt {nt:9d} tiles {tiles:6d}  kernel {ms/k*1e3:8.1f} us  step {e0.elapsed_time(e1)/n*1e3:8.1f} us   ideal MMA {tiles*16*1024/148/1.965e3:8.1f} us at 1965 MHz")
