#!/bin/bash
# First GPU call of round 2 (1 GPU): the fp4 probe prepared at the end of round 1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python slam-1_b200/csrc/microbench/build.py
timeout 120 slam-1_b200/csrc/microbench/bin/fp4_probe > gpurun_out/fp4_probe.txt 2>&1; echo "fp4 probe exit $?" >> gpurun_out/fp4_probe.txt
timeout 120 slam-1_b200/csrc/microbench/bin/tc_probe > gpurun_out/tc_probe.txt 2>&1
cat gpurun_out/fp4_probe.txt; tail -4 gpurun_out/tc_probe.txt
