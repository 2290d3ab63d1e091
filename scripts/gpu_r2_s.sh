#!/bin/bash
# After the refine rework (carry-save POPC, branch-free row runs, one fence per block): full GPU suite, config-4 shard
# profile (one- and two-phase launch lists), c3 / c4 / c5 bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu_s.txt 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_s.txt
for w in 8 4; do
  SLM_EXCHANGE_TWO_PHASE_MIN=0 python scripts/c4_shard_profile.py $w 10
  python scripts/c4_shard_profile.py $w 10
done
SLM_EXCHANGE_TWO_PHASE_MIN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4shard_onephase.csv python scripts/c4_shard_profile.py 8 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c4shard_twophase.csv python scripts/c4_shard_profile.py 8 3 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ('onephase', 'twophase'):
    rows = list(csv.reader(open(f'gpurun_out/launches_c4shard_{f}.csv', errors='ignore')))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
    agg = collections.defaultdict(list)
    for r in rows[h + 2:]:
        if len(r) > vi: agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    print(f)
    for k, v in agg.items(): print('   ', k, len(v), round(sum(v) / len(v) / 1000, 1), 'us')
PY
python bench.py --configs c3,c4 --steps 20 --warmup 3 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_s.json') if l.startswith('{')][-1])
print('c5', round(d['value'], 1), d['ms_per_step'], d['roofline']['kernel_ms'], d['parity_check']['ok'])
for k, c in d['configs'].items(): print(k, round(c['value'], 1), c['ms_per_step'], c.get('kernel_ms'), (c.get('parity_check') or {}).get('ok'))
PY
