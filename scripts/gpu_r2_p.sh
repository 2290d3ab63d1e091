#!/bin/bash
# Round 2 (1 GPU): chained batches on the mxf4 kernel (config 3): parity, then c3 with the fp8 and the fp4 kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_batched_chain_gpu.py tests/test_parity_gpu.py -q -m gpu -x > gpurun_out/pytest_sel.txt 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_sel.txt
for v in tensor tensor4 auto; do
    timeout 300 python bench.py --workload c3 --variant $v --no-cpu --configs none --e2e-steps 1 --steps 10 > gpurun_out/bench_c3_$v.json 2>gpurun_out/err.txt; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_c3_$v.json') if l.startswith('{')][-1]); print('c3 $v', round(d['value'],1), 'Gcmp/s', round(d['ms_per_step'],4), 'ms; kernel', round(d['roofline']['kernel_ms'],4), d['roofline']['kernel'], 'frac', round(d['roofline']['frac'],3), d['parity_check'], d['gpu_launches'])" || tail -5 gpurun_out/err.txt
done
C3="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs none --e2e-steps 1 --workload c3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3_tc4.csv $C3 > gpurun_out/ncu_list_c3.log 2>&1; echo "ncu list exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c3_tc4.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[1:][-6:]: print(r[k][:70], r[v])
PY
