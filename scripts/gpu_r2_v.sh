#!/bin/bash
# Re-entry validation of HEAD on one GPU (the container was re-created; the last session's gpurun_out/ is gone): the whole
# GPU suite incl. the new masked-search tests, smoke, one default bench line, one rank's share of config 4 at 8 ranks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 700 python -m pytest tests -q -x -m gpu --timeout 200 > gpurun_out/pytest_gpu_v.txt 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_v.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_v.txt 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_v.txt
timeout 500 python bench.py > gpurun_out/bench_default_v.json 2> gpurun_out/bench_default_v.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_default_v.json') if l.startswith('{')][-1])
    print('c5', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'kernel', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3),
          'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check']['ok'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 2))
    for k, c in d['configs'].items():
        if 'error' in c: print(k, 'ERROR', c['error']); continue
        print(k, round(c['value'], 1), 'ms', round(c['ms_per_step'], 4), 'kernel', c.get('kernel'), c.get('kernel_ms'), 'frac', c.get('roofline_frac'), 'e2e', round(c['e2e']['value'], 1), 'parity', (c.get('parity_check') or {}).get('ok'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_default_v.err').read()[-2000:])
PY
SLM_EXCHANGE_TWO_PHASE_MIN=0 timeout 100 python scripts/c4_shard_profile.py 8 10
timeout 100 python scripts/c4_shard_profile.py 8 10
exit 0
