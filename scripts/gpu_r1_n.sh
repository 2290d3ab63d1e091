#!/bin/bash
# 2-GPU check after this session's changes: NCCL / NVLink / all-to-all exchanges against the oracle, c5 and c4 lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s > gpurun_out/pytest_multi2.txt 2>&1; echo "pytest multi exit $?" >> gpurun_out/status.txt
run() { # $1 = tag, rest = bench args
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 "$@" > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err; echo "n2 $tag exit $?" >> gpurun_out/status.txt
}
run c5
run c4_nccl --workload c4 --exchange nccl
run c4_a2a --workload c4 --exchange a2a
cat gpurun_out/status.txt
tail -3 gpurun_out/pytest_multi2.txt
for t in c5 c4_nccl c4_a2a; do python -c "
import json
d=json.load(open('gpurun_out/n2_$t.json')); print('$t', d['value'], d['ms_per_step'], d['config']['parallelism'])"; done
