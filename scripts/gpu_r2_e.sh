#!/bin/bash
# Round 2, call 3 (1 GPU): loopback tests after the done_counter fix, the new bench.py (all configs in one line), fp8 vs mxf4 on c5 / c4,
# launch list + full ncu capture of the mxf4 kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_exchange_loopback_gpu.py tests/test_config4_regime_gpu.py -q -m gpu > gpurun_out/pytest_loopback.txt 2>&1; echo "pytest loopback+c4 exit $?"; tail -8 gpurun_out/pytest_loopback.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?"; tail -3 gpurun_out/bench_default.err
python scripts/summ.py gpurun_out/bench_default.json 2>/dev/null || python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('c5', d['config']['variant'], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'min', round(d['ms_min'],4), 'kernel', d['roofline']['kernel'], round(d['roofline']['kernel_ms'],4), 'frac', d['roofline']['frac'], 'peak', d['roofline']['peak'], 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_check'])
for k,v in d['configs'].items(): print(k, v['variant'], round(v['value'],1), 'ms', round(v['ms_per_step'],4), v['kernel'], round(v['kernel_ms'],4), 'frac', v['roofline_frac'], 'e2e', round(v['e2e']['value'],1), 'parity', v['parity_check'] and v['parity_check']['ok'])
PY
for v in tensor tensor4; do
  for wl in c5 c4; do
    timeout 300 python bench.py --workload $wl --variant $v --no-cpu --configs none --e2e-steps 1 --steps 10 > gpurun_out/bench_${wl}_${v}.json 2>gpurun_out/bench_${wl}_${v}.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_${wl}_${v}.json') if l.startswith('{')][-1]); print('$wl $v', round(d['value'],1), 'Gcmp/s', round(d['ms_per_step'],4), 'ms; kernel', round(d['roofline']['kernel_ms'],4), 'ms', d['roofline']['kernel'], 'frac', d['roofline']['frac'], 'peak', d['roofline']['peak'], d['clocks'])" || tail -5 gpurun_out/bench_${wl}_${v}.err
  done
done
SLM_TC4_CHUNK=40 timeout 300 python bench.py --workload c5 --variant tensor4 --no-cpu --configs none --e2e-steps 1 --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c5 tensor4 chunk40', round(d['value'],1), round(d['roofline']['kernel_ms'],4))"
SLM_TC4_CHUNK=120 timeout 300 python bench.py --workload c4 --variant tensor4 --no-cpu --configs none --e2e-steps 1 --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c4 tensor4 chunk120', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
C5="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs none --e2e-steps 1 --workload c5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c5_tc4.csv $C5 > gpurun_out/ncu_list_c5.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc4_kernel -s 3 -c 1 -o gpurun_out/prof_tc4_c5 $C5 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu full exit $?"
