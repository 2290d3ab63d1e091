#!/bin/bash
# One rank's share of config 4 on one GPU: step time and launch list (one-phase and two-phase), ncu capture of the refine kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
for w in 8 4; do
  SLM_EXCHANGE_TWO_PHASE_MIN=0 python scripts/c4_shard_profile.py $w 10
  python scripts/c4_shard_profile.py $w 10
done
SLM_EXCHANGE_TWO_PHASE_MIN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4shard_onephase.csv python scripts/c4_shard_profile.py 8 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c4shard_twophase.csv python scripts/c4_shard_profile.py 8 3 > /dev/null 2>&1
SLM_EXCHANGE_TWO_PHASE_MIN=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_refine_kernel -s 3 -c 1 -o gpurun_out/ncu_c4shard_refine -f python scripts/c4_shard_profile.py 8 2 > /dev/null 2>&1
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
python -m pytest tests/test_batched_chain_gpu.py tests/test_config4_regime_gpu.py -q -x -m gpu 2>&1 | tail -2
python bench.py --configs c3 --steps 10 --warmup 3 --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); c=d['configs']['c3']; print('c3', round(c['value'],1), c['ms_per_step'], c.get('kernel_ms'))"
