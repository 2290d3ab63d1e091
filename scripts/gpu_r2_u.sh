#!/bin/bash
# Round 2 closing run on 8 GPUs (v3: after the refine rework and the tiled two-phase kernels): multi-rank parity tests at 8
# ranks, default bench at 8 / 4 / 2 ranks (c5 headline + c4 train-sharded + c4q query-sharded), one-phase A/B of c4 at 8.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
SLM_TEST_WORLDS=$NG timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -s --timeout 500 > gpurun_out/pytest_multi_n${NG}_v3.txt 2>&1; echo "pytest multi exit $?"; grep -E "world|passed|failed|Error" gpurun_out/pytest_multi_n${NG}_v3.txt | tail -8
run() {  # n, tag, extra args
  local n=$1; shift; local tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/scale_${tag}_n${n}_v3.json 2> gpurun_out/scale_${tag}_n${n}_v3.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scale_${tag}_n${n}_v3.json') if l.startswith('{')][-1])
    print('${tag} n=${n}', d['config']['workload'][:3], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'min', round(d['ms_min'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_check']['ok'], 'launches', d['gpu_launches'])
    for k, c in d['configs'].items():
        print('    ', k, round(c['value'],1), 'ms', round(c['ms_per_step'],4), 'kernel', round(c.get('kernel_ms') or 0,4), 'e2e', round(c['e2e']['value'],1), 'parity', (c.get('parity_check') or {}).get('ok'), 'launches', c['gpu_launches'], c['parallelism'][-70:])
except Exception as e:
    print('${tag} n=${n} FAILED', e); print(open('gpurun_out/scale_${tag}_n${n}_v3.err').read()[-1500:])
PY
}
run $NG auto
SLM_EXCHANGE_TWO_PHASE_MIN=0 run $NG onephase --workload c4 --configs none
[ $NG -ge 8 ] && run 4 auto
[ $NG -ge 4 ] && run 2 auto
exit 0
