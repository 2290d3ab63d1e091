#!/bin/bash
# Validation of the host-path staging, knn2_batched, wide chi-square scan: their tests, then the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 500 python -m pytest tests/test_host_path_gpu.py tests/test_bow.py tests/test_batched_chain_gpu.py tests/test_masked_gpu.py -q -m gpu --timeout 200 > gpurun_out/pytest_gpu_y.txt 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu_y.txt | cut -c1-300
for th in 0 2 4 8 16; do
SLM_HOST_STAGE_THREADS=$th timeout 200 python - <<'PY'
import os, time, numpy as np, torch, slammatch
from slammatch import synth
q = synth.uniform(2000, 1); t = synth.uniform(10_000_000, 2)
for _ in range(2): slammatch.knn2(q, t)
t0 = time.perf_counter()
for _ in range(4): slammatch.knn2(q, t)
dt = (time.perf_counter() - t0) / 4
print('threads', os.environ['SLM_HOST_STAGE_THREADS'], 'pageable c5 call', round(dt * 1e3, 2), 'ms', round(2e10 / dt / 1e9, 1), 'Gcmp/s')
PY
done
timeout 500 python bench.py > gpurun_out/bench_default_y.json 2> gpurun_out/bench_default_y.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_default_y.json') if l.startswith('{')][-1])
    print('c5', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'pageable', round(d['e2e']['pageable']['value'], 1), 'parity', d['parity_check']['ok'])
    for k, c in d['configs'].items():
        if 'error' in c: print(k, 'ERROR', c['error']); continue
        print(k, round(c['value'], 1), 'ms', round(c['ms_per_step'], 4), 'e2e', round(c['e2e']['value'], 1), 'pageable', (c['e2e'].get('pageable') or {}).get('value'), 'parity', (c.get('parity_check') or {}).get('ok'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_default_y.err').read()[-2000:])
PY
exit 0
