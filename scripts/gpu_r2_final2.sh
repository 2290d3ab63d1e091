#!/bin/bash
# Closing single-GPU evidence of round 2 (second session): whole GPU suite, smoke, default bench line, rates of the masked
# search and the wide chi-square scan, launch lists of the c5 and c3 bench commands, ncu --set full of the reworked
# shared-memory refine kernel (c3) and of the masked kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 700 python -m pytest tests -q -x -m gpu --timeout 200 > gpurun_out/pytest_gpu_final2.txt 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_final2.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_final2.txt 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final2.txt
timeout 200 python scripts/masked_chi2_rates.py > gpurun_out/masked_chi2_rates.txt 2>&1; echo "rates exit $?"; cat gpurun_out/masked_chi2_rates.txt
timeout 500 python bench.py > gpurun_out/bench_default_final2.json 2> gpurun_out/bench_default_final2.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_default_final2.json') if l.startswith('{')][-1])
    print('c5', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'kernel', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value'], 1), 'pageable', round(d['e2e']['pageable']['value'], 1), 'resident', round(d['e2e']['resident_db']['value'], 1), 'parity', d['parity_check']['ok'])
    for k, c in d['configs'].items():
        if 'error' in c: print(k, 'ERROR', c['error']); continue
        print(k, round(c['value'], 1), 'ms', round(c['ms_per_step'], 4), 'kernel', c.get('kernel'), round(c.get('kernel_ms') or 0, 4), 'frac', c.get('roofline_frac'), 'e2e', round(c['e2e']['value'], 1), 'parity', (c.get('parity_check') or {}).get('ok'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_default_final2.err').read()[-2000:])
PY
B="bench.py --configs none --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5_final2.csv python $B > /dev/null 2>&1; echo "ncu c5 list exit $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3_final2.csv python $B --workload c3 > /dev/null 2>&1; echo "ncu c3 list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_refine_frame_kernel -s 3 -c 1 -f -o gpurun_out/ncu_refine_frame_c3_final2 python $B --workload c3 > /dev/null 2>&1; echo "ncu refine exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:knn2_masked_kernel -s 3 -c 1 -f -o gpurun_out/ncu_masked_final2 python scripts/masked_chi2_rates.py > /dev/null 2>&1; echo "ncu masked exit $?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
python - <<'PY'
import csv, collections
for f in ('c5', 'c3'):
    try:
        rows = list(csv.reader(open(f'gpurun_out/launches_{f}_final2.csv', errors='ignore')))
        h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
        ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
        agg = collections.defaultdict(list)
        for r in rows[h + 2:]:
            if len(r) > vi: agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
        print(f)
        for k, v in agg.items(): print('   ', k, len(v), round(sum(v) / len(v) / 1000, 1), 'us')
    except Exception as e:
        print(f, 'launch list failed', e)
PY
exit 0
