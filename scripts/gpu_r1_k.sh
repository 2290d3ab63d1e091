#!/bin/bash
# N-GPU: NVLink exchange validation (pytest multi) + c5 with both exchanges
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -s > gpurun_out/pytest_multi.txt 2>&1; echo "pytest multi exit $?" >> gpurun_out/status.txt
for ex in nccl auto; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu --exchange $ex > gpurun_out/scale_c5_n${NG}_$ex.json 2> gpurun_out/scale_c5_n${NG}_$ex.err; echo "bench c5 n$NG $ex exit $?" >> gpurun_out/status.txt
done
cat gpurun_out/status.txt; tail -5 gpurun_out/pytest_multi.txt
