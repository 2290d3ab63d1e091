#!/bin/bash
# Round 2 (1 GPU): one ncu capture of the split-job mxf4 kernel (v6) to see why it is slow.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
C5="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs none --e2e-steps 1 --workload c5"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc4_kernel -s 3 -c 1 -o gpurun_out/prof_tc4_v6_c5 $C5 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu full exit $?"
