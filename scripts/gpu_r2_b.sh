#!/bin/bash
# Round 2, 8-GPU box: config-4 scaling with the NCCL all-gather vs the all-to-all exchange.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for ex in nccl a2a; do
  for n in 1 2 4 8; do
    [ $n -le $NG ] || continue
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workload c4 > gpurun_out/c4_${ex}_n1.json 2> gpurun_out/c4_${ex}_n1.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workload c4 --exchange $ex > gpurun_out/c4_${ex}_n${n}.json 2> gpurun_out/c4_${ex}_n${n}.err
    fi
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/c4_${ex}_n${n}.json') if l.startswith('{')][-1]); print('$ex', $n, round(d['value'],1), round(d['ms_per_step'],4))"
  done
done
