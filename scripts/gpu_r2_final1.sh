#!/bin/bash
# Round 2 closing evidence on 1 GPU: whole GPU suite, smoke, reference arm, default bench (all configs), fixed cost, launch lists and a
# full ncu capture of the final mxf4 kernel (each ncu pass only after the same command exited 0 without ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" | tee -a gpurun_out/status.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" | tee -a gpurun_out/status.txt; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" | tee -a gpurun_out/status.txt; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('c5', d['config']['variant'], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'min', round(d['ms_min'],4), 'kernel', d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'peak', round(d['roofline']['peak']), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('pageable'), d['e2e'].get('resident_db',{}).get('value'), 'parity', d['parity_check']['ok'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
for k,v in d['configs'].items(): print(k, v['variant'], round(v['value'],1), 'ms', round(v['ms_per_step'],4), v['kernel'], round(v['kernel_ms'],4), 'frac', v['roofline_frac'] and round(v['roofline_frac'],3), 'e2e', round(v['e2e']['value'],1), 'parity', v['parity_check'] and v['parity_check']['ok'])
r=json.loads([l for l in open('gpurun_out/bench_reference.json') if l.startswith('{')][-1]); print('reference', r['value'], r['cpu_baseline']['cores'], r['config']['sample'])
PY
for v in tensor popc bmma; do
  timeout 300 python bench.py --workload c5 --variant $v --no-cpu --configs none --e2e-steps 1 --steps 5 > gpurun_out/bench_c5_${v}.json 2>gpurun_out/err.txt; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_c5_${v}.json') if l.startswith('{')][-1]); print('c5 $v', round(d['value'],1), 'Gcmp/s', round(d['ms_per_step'],4), 'ms; kernel', round(d['roofline']['kernel_ms'],4), d['roofline']['kernel'], 'frac', d['roofline']['frac'], d['parity_check']['ok'])" || tail -3 gpurun_out/err.txt
done
timeout 300 python scripts/tc4_fixed_cost.py tensor4 > gpurun_out/tc4_fixed_cost.txt 2>&1; tail -7 gpurun_out/tc4_fixed_cost.txt
timeout 300 python scripts/cpu_overhead.py > gpurun_out/cpu_overhead.txt 2>&1; echo "cpu overhead exit $?" | tee -a gpurun_out/status.txt
C5="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs none --e2e-steps 1 --workload c5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c5_final.csv $C5 > gpurun_out/ncu_list_c5.log 2>&1; echo "ncu list c5 exit $?" | tee -a gpurun_out/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc4_kernel -s 3 -c 1 -o gpurun_out/prof_tc4_final_c5 $C5 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu full exit $?" | tee -a gpurun_out/status.txt
