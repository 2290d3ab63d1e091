#!/bin/bash
# Round 2 (1 GPU): mxf4 kernel v7 (all three pieces requested at once, accumulator handed back before any reduction): parity, timing, benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_config4_regime_gpu.py -q -m gpu -k "tensor4 or auto" > gpurun_out/pytest_sel.txt 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_sel.txt
bash scripts/gpu_r2_n.sh 2>&1 | grep "tc4 timing" | tail -3
timeout 300 python scripts/tc4_fixed_cost.py tensor4 2>&1 | tee gpurun_out/tc4_fixed_cost.txt | tail -3
for wl in c5 c4; do
    timeout 300 python bench.py --workload $wl --variant tensor4 --no-cpu --configs none --e2e-steps 1 --steps 10 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl tensor4', round(d['value'],1), 'Gcmp/s', round(d['ms_per_step'],4), 'ms; kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), d['parity_check']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/err.txt
done
