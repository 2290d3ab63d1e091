#!/bin/bash
# Round 2 (1 GPU): where a job's cycles go in the mxf4 kernel (diagnostic build, SLM_TC4_TIMING=1), on c5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SLM_TC4_TIMING=1 timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/tc4_timing.txt
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'slam-1_b200')
import torch, slammatch
from slammatch import synth
q = torch.from_numpy(synth.uniform(2000, 1)).cuda()
for nt in (1_250_000, 10_000_000):
    t = torch.from_numpy(synth.uniform(nt, 2)).cuda()
    for _ in range(2):
        slammatch.knn2(q, t, variant="tensor4")
    torch.cuda.synchronize()
    print("nt", nt, flush=True)
PY
