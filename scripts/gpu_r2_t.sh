#!/bin/bash
# Quick validation of the two-phase kernels: loopback + config-4 regime tests (per-test timeout), shard profile.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 400 python -m pytest tests/test_exchange_loopback_gpu.py tests/test_config4_regime_gpu.py -q -x -m gpu --timeout 100 > gpurun_out/pytest_gpu_t.txt 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_t.txt
for w in 8; do
  SLM_EXCHANGE_TWO_PHASE_MIN=0 timeout 120 python scripts/c4_shard_profile.py $w 10
  timeout 120 python scripts/c4_shard_profile.py $w 10
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c4shard_twophase.csv python scripts/c4_shard_profile.py 8 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ('twophase',):
    rows = list(csv.reader(open(f'gpurun_out/launches_c4shard_{f}.csv', errors='ignore')))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
    agg = collections.defaultdict(list)
    for r in rows[h + 2:]:
        if len(r) > vi: agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    print(f)
    for k, v in agg.items(): print('   ', k, len(v), round(sum(v) / len(v) / 1000, 1), 'us')
PY
