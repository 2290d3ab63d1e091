#!/bin/bash
# Round 2, call 1 (1 GPU): mxf4 probes (single CTA and the product's CTA-pair shape), fp8 probe for the same box,
# experimental planner A/B on c2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=slam-1_b200/csrc/microbench/bin
for a in "" "128 1024 256" "128 2048 256" "1024 128 256"; do
  echo "== fp4_probe $a" >> gpurun_out/fp4_probe.txt
  timeout 120 $B/fp4_probe $a >> gpurun_out/fp4_probe.txt 2>&1; echo "exit $?" >> gpurun_out/fp4_probe.txt
done
for a in "240 480" "224 480" "128 480" "256 256"; do
  echo "== fp4_probe2 $a" >> gpurun_out/fp4_probe2.txt
  timeout 120 $B/fp4_probe2 $a >> gpurun_out/fp4_probe2.txt 2>&1; echo "exit $?" >> gpurun_out/fp4_probe2.txt
done
timeout 120 $B/tc_probe > gpurun_out/tc_probe.txt 2>&1
cat gpurun_out/fp4_probe.txt gpurun_out/fp4_probe2.txt; tail -6 gpurun_out/tc_probe.txt
SLM_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_frame_gpu.py -q -m gpu -k experimental 2>&1 | tail -2
for v in 0 1; do
  if [ $v -eq 1 ]; then export SLM_TC_PLAN_MT=1; else unset SLM_TC_PLAN_MT; fi
  timeout 200 python bench.py --workload c2 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('plan_mt=$v c2', round(d['value'],1), round(d['ms_per_step']*1e3,1), 'us; kernel', round(d['roofline']['kernel_ms']*1e3,1), 'us')"
done
unset SLM_TC_PLAN_MT
