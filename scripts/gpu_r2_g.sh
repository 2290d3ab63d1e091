#!/bin/bash
# Round 2, call 7 (1 GPU): tc4 epilogue v5 (pieces 0+1 requested together, lean barrier wait, uniform TMEM addresses).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_config4_regime_gpu.py -q -m gpu > gpurun_out/pytest_sel.txt 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_sel.txt
for wl in c5 c4; do
    timeout 300 python bench.py --workload $wl --variant tensor4 --no-cpu --configs none --e2e-steps 1 --steps 10 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$wl tensor4', round(d['value'],1), 'Gcmp/s', round(d['ms_per_step'],4), 'ms; kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), d['parity_check']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/err.txt
done
C5="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs none --e2e-steps 1 --workload c5"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc4_kernel -s 3 -c 1 -o gpurun_out/prof_tc4_v5_c5 $C5 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu full exit $?"
