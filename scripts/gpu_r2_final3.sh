#!/bin/bash
# Closing run after the chi-square rewrite: whole GPU suite, smoke, default bench line, ncu --set full of the wide chi-square scan.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 700 python -m pytest tests -q -x -m gpu --timeout 200 > gpurun_out/pytest_gpu_final3.txt 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_final3.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_final3.txt 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final3.txt
timeout 500 python bench.py > gpurun_out/bench_default_final3.json 2> gpurun_out/bench_default_final3.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_default_final3.json') if l.startswith('{')][-1])
    print('c5', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'kernel', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value'], 1), 'pageable', round(d['e2e']['pageable']['value'], 1), 'resident', round(d['e2e']['resident_db']['value'], 1), 'parity', d['parity_check']['ok'])
    for k, c in d['configs'].items():
        if 'error' in c: print(k, 'ERROR', c['error']); continue
        print(k, round(c['value'], 1), 'ms', round(c['ms_per_step'], 4), 'kernel', c.get('kernel'), round(c.get('kernel_ms') or 0, 4), 'frac', c.get('roofline_frac'), 'e2e', round(c['e2e']['value'], 1), 'parity', (c.get('parity_check') or {}).get('ok'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_default_final3.err').read()[-2000:])
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chi2_scan_wide_kernel -s 40 -c 1 -f -o gpurun_out/ncu_chi2_wide_final3 python scripts/masked_chi2_rates.py > /dev/null 2>&1; echo "ncu chi2 exit $?"
ls -la gpurun_out/ncu_chi2_wide_final3.ncu-rep
exit 0
