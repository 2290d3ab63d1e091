#!/bin/bash
# Round 2 (1 GPU): tc4 v6 (jobs issued as two half-N MMA chains): parity, fixed cost, c5 / c4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_config4_regime_gpu.py tests/test_exchange_loopback_gpu.py tests/test_frame_gpu.py -q -m gpu > gpurun_out/pytest_sel.txt 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_sel.txt
timeout 300 python scripts/tc4_fixed_cost.py tensor4 2>&1 | tee gpurun_out/tc4_fixed_cost.txt
run() {  # n, tag, extra args
  local n=$1; shift; local tag=$1; shift
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --configs c4 "$@" > gpurun_out/scale_${tag}_n1.json 2> gpurun_out/scale_${tag}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/scale_${tag}_n${n}.json 2> gpurun_out/scale_${tag}_n${n}.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scale_${tag}_n${n}.json') if l.startswith('{')][-1])
    c4=d['configs'].get('c4',{})
    print('${tag} n=${n} c5', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'min', round(d['ms_min'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches'], 'parity', d['parity_check']['ok'],
          '| c4', round(c4.get('value',0),1), 'ms', round(c4.get('ms_per_step',0),4), 'kernel', round(c4.get('kernel_ms',0),4), 'launches', c4.get('gpu_launches'), 'parity', (c4.get('parity_check') or {}).get('ok'))
except Exception as e:
    print('${tag} n=${n} FAILED', e); print(open('gpurun_out/scale_${tag}_n${n}.err').read()[-1500:])
PY
}
run 1 auto
