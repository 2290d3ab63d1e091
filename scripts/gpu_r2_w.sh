#!/bin/bash
# 2-GPU validation of HEAD's exchange kernels over real NVLink: the multi-rank parity test on the NVLink exchange with the
# two-phase form forced at 2 ranks (one-phase for the c5-like case, two-phase for the c4-like one), then the default
# 2-rank bench line (c5 headline + train-sharded c4 + query-sharded c4q, parity_check on every rank).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
SLM_TEST_WORLDS=2 SLM_TEST_EXCHANGES=auto SLM_EXCHANGE_TWO_PHASE_WORLD=2 timeout 300 python -m pytest tests/test_multi_gpu.py -q -m gpu -s --timeout 250 > gpurun_out/pytest_multi_n2_twophase_w.txt 2>&1; echo "pytest multi exit $?"; grep -E "world|passed|failed|Error" gpurun_out/pytest_multi_n2_twophase_w.txt | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale_auto_n2_w.json 2> gpurun_out/scale_auto_n2_w.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/scale_auto_n2_w.json') if l.startswith('{')][-1])
    print('n=2 c5', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'min', round(d['ms_min'], 4), 'kernel', round(d['roofline']['kernel_ms'], 4), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check'], 'launches', d['gpu_launches'])
    for k, c in d['configs'].items():
        if 'error' in c: print(k, 'ERROR', c['error']); continue
        print('    ', k, round(c['value'], 1), 'ms', round(c['ms_per_step'], 4), 'kernel', round(c.get('kernel_ms') or 0, 4), 'e2e', round(c['e2e']['value'], 1), 'parity', (c.get('parity_check') or {}).get('ok'), c['parallelism'][-80:])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/scale_auto_n2_w.err').read()[-1500:])
PY
exit 0
