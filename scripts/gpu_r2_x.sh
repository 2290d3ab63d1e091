#!/bin/bash
# 4-GPU validation of the default sharded path (two-phase refine from 4 ranks on): default bench line with parity_check on
# every rank, then the one-phase A/B of config 4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
run() {  # tag, extra args
  local tag=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --steps 20 --warmup 3 "$@" > gpurun_out/scale_${tag}_n${NG}_x.json 2> gpurun_out/scale_${tag}_n${NG}_x.err; echo "bench $tag exit $?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/scale_${tag}_n${NG}_x.json') if l.startswith('{')][-1])
    print('${tag} n=${NG}', d['config']['workload'][:3], round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'min', round(d['ms_min'], 4), 'kernel', round(d['roofline']['kernel_ms'], 4), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check']['ok'], d['parity_check']['ranks_checked'])
    for k, c in d['configs'].items():
        if 'error' in c: print(k, 'ERROR', c['error']); continue
        print('    ', k, round(c['value'], 1), 'ms', round(c['ms_per_step'], 4), 'kernel', round(c.get('kernel_ms') or 0, 4), 'e2e', round(c['e2e']['value'], 1), 'parity', (c.get('parity_check') or {}).get('ok'), c['parallelism'][-80:])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/scale_${tag}_n${NG}_x.err').read()[-1500:])
PY
}
run auto
SLM_EXCHANGE_TWO_PHASE_MIN=0 run onephase --workload c4 --configs none --e2e-steps 1
exit 0
