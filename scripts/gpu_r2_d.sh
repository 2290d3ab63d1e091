#!/bin/bash
# Round 2, call 2 (1 GPU): first run of the mxf4 kernel + the new exchange: targeted tests first, then the whole suite, then quick benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "tensor4 or auto" > gpurun_out/pytest_tc4.txt 2>&1; echo "pytest tc4 exit $?"; tail -15 gpurun_out/pytest_tc4.txt
timeout 600 python -m pytest tests/test_exchange_loopback_gpu.py -q -m gpu -x > gpurun_out/pytest_loopback.txt 2>&1; echo "pytest loopback exit $?"; tail -15 gpurun_out/pytest_loopback.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?"; tail -25 gpurun_out/pytest_gpu.txt
for v in tensor tensor4; do
  for wl in c5 c4 c2; do
    timeout 300 python bench.py --workload $wl --variant $v --no-cpu --e2e-steps 1 --steps 10 2>gpurun_out/bench_${wl}_${v}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl $v', round(d['value'],1), 'Gcmp/s', round(d['ms_per_step'],4), 'ms; kernel', round(d['roofline']['kernel_ms'],4), 'ms', d['roofline']['kernel'])" || tail -5 gpurun_out/bench_${wl}_${v}.err
  done
done
