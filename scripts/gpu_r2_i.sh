#!/bin/bash
# Round 2 (2 GPUs): c4 sharded, NVLink exchange (coalesced peer stores) vs NCCL all-gather vs all-to-all; launch list of c4 on one GPU.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_exchange_loopback_gpu.py -q -m gpu > gpurun_out/pytest_multi2.txt 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_multi2.txt
for ex in auto nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload c4 --configs none --exchange $ex > gpurun_out/c4_${ex}_n2.json 2> gpurun_out/c4_${ex}_n2.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/c4_${ex}_n2.json') if l.startswith('{')][-1]); print('c4 n=2 $ex', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['parity_check']['ok'], d['config']['parallelism'][-50:])" || tail -5 gpurun_out/c4_${ex}_n2.err
done
C4="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs none --e2e-steps 1 --workload c4"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c4_tc4.csv $C4 > gpurun_out/ncu_list_c4.log 2>&1; echo "ncu list exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c4_tc4.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[1:][-12:]: print(r[k][:60], r[v])
PY
