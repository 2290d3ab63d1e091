#!/bin/bash
# multi-GPU: NCCL sharded path parity + scaling of c5 / c4 (run with gpurun --gpus N)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
NG=$(nvidia-smi -L | wc -l)
echo "gpus $NG" >> gpurun_out/status.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_multi.txt 2>&1; echo "pytest multi exit $?" >> gpurun_out/status.txt
for wl in c5 c4; do
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --workload $wl --no-cpu > gpurun_out/scale_${wl}_n1.json 2> gpurun_out/scale_${wl}_n1.err; echo "bench $wl n1 exit $?" >> gpurun_out/status.txt
  for n in 2 4 8; do
    if [ $n -le $NG ]; then
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --workload $wl --no-cpu > gpurun_out/scale_${wl}_n${n}.json 2> gpurun_out/scale_${wl}_n${n}.err; echo "bench $wl n$n exit $?" >> gpurun_out/status.txt
    fi
  done
done
cat gpurun_out/status.txt
