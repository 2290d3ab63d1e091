#!/bin/bash
# First GPU call of round 2 (1 GPU): the fp4 probe prepared at the end of round 1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python slam-1_b200/csrc/microbench/build.py
timeout 120 slam-1_b200/csrc/microbench/bin/fp4_probe > gpurun_out/fp4_probe.txt 2>&1; echo "fp4 probe exit $?" >> gpurun_out/fp4_probe.txt
timeout 120 slam-1_b200/csrc/microbench/bin/tc_probe > gpurun_out/tc_probe.txt 2>&1
cat gpurun_out/fp4_probe.txt; tail -4 gpurun_out/tc_probe.txt
# experimental tensor-kernel planner (query tiles per CTA): parity, then c2 with and without it
SLM_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_frame_gpu.py -q -m gpu -k experimental 2>&1 | tail -2
for v in 0 1; do
  if [ $v -eq 1 ]; then export SLM_TC_PLAN_MT=1; else unset SLM_TC_PLAN_MT; fi
  timeout 200 python bench.py --workload c2 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('plan_mt=$v c2', round(d['value'],1), round(d['ms_per_step']*1e3,1), 'us; kernel', round(d['roofline']['kernel_ms']*1e3,1), 'us')"
done
unset SLM_TC_PLAN_MT
