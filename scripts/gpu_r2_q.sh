#!/bin/bash
# Round 2 closing run on 8 GPUs: multi-rank parity tests, default bench at 8 / 4 / 2 ranks, and at 8 ranks the A/Bs of the
# exchange (NCCL all-gather; one-phase refine forced with SLM_EXCHANGE_TWO_PHASE_MIN=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -s > gpurun_out/pytest_multi_n${NG}_v2.txt 2>&1; echo "pytest multi exit $?"; grep -E "world|passed|failed|Error" gpurun_out/pytest_multi_n${NG}_v2.txt | tail -12
run() {  # n, tag, extra args
  local n=$1; shift; local tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/scale_${tag}_n${n}_v2.json 2> gpurun_out/scale_${tag}_n${n}_v2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scale_${tag}_n${n}_v2.json') if l.startswith('{')][-1])
    c4=d['configs'].get('c4',{})
    print('${tag} n=${n} c5', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'min', round(d['ms_min'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_check']['ok'], d['config']['parallelism'][-60:],
          '| c4', round(c4.get('value',0),1), 'ms', round(c4.get('ms_per_step',0),4), 'kernel', round(c4.get('kernel_ms',0),4), 'parity', (c4.get('parity_check') or {}).get('ok'), str(c4.get('parallelism',''))[-50:])
except Exception as e:
    print('${tag} n=${n} FAILED', e); print(open('gpurun_out/scale_${tag}_n${n}_v2.err').read()[-1500:])
PY
}
run $NG auto
SLM_EXCHANGE_TWO_PHASE_MIN=0 run $NG onephase
run $NG nccl --exchange nccl
[ $NG -ge 8 ] && run 4 auto
[ $NG -ge 4 ] && run 2 auto
exit 0
