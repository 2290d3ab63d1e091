#!/usr/bin/env python
"""Fixed cost of one sharded step: kernel time (library-internal CUDA events) and whole device-resident call time of
slammatch.knn2 for 2000 queries against train sets from one tile per cluster to a 1.25M-row shard (config 5 at 8 ranks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-1_b200"))
import numpy as np, torch
import slammatch
from slammatch import synth, _lib

ctx = _lib.context(0)
q = torch.from_numpy(synth.uniform(2000, 1)).cuda()
variant = sys.argv[1] if len(sys.argv) > 1 else "tensor4"
for nt in (17_760, 88_800, 177_600, 355_200, 710_400, 1_250_000, 2_500_000):
    t = torch.from_numpy(synth.uniform(nt, 2)).cuda()
    for _ in range(5):
        slammatch.knn2(q, t, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        slammatch.knn2(q, t, variant=variant)
    e1.record(); torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_read()          # kernel time in a separate pass: the event records serialise the launches
    for _ in range(n):
        slammatch.knn2(q, t, variant=variant)
    torch.cuda.synchronize()
    km, kn = ctx.profile_read(); ctx.profile(False)
    tiles = -(-nt // 240)
    print(f"{variant} nt={nt:8d} tiles/cluster={tiles/74:6.1f}  kernel {km/kn*1e3:8.1f} us  call {e0.elapsed_time(e1)/n*1e3:8.1f} us  "
          f"ideal MMA {tiles/74*8*480/1.965e3:8.1f} us")
