#!/bin/bash
# Round 2 (1 GPU): whole GPU suite after the two-phase exchange / refine unrolling / early TMA prologue; fixed cost of one sharded
# step; default bench line (all configs).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?"; tail -6 gpurun_out/pytest_gpu.txt
timeout 300 python scripts/tc4_fixed_cost.py tensor4 2>&1 | tee gpurun_out/tc4_fixed_cost.txt
timeout 300 python scripts/tc4_fixed_cost.py tensor 2>&1 | tail -3 | tee -a gpurun_out/tc4_fixed_cost.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('c5', d['config']['variant'], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'min', round(d['ms_min'],4), 'kernel', d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_check']['ok'], 'cpu', d['cpu_baseline']['value'])
for k,v in d['configs'].items(): print(k, v['variant'], round(v['value'],1), 'ms', round(v['ms_per_step'],4), v['kernel'], round(v['kernel_ms'],4), 'frac', v['roofline_frac'] and round(v['roofline_frac'],3), 'e2e', round(v['e2e']['value'],1), 'parity', v['parity_check'] and v['parity_check']['ok'])
PY
