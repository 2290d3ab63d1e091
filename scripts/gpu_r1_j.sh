#!/bin/bash
# 8-GPU: c5 scaling after the start-up / in-place all-gather changes + a trace of rank 0's timeline at N=8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
NG=$(nvidia-smi -L | wc -l)
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_c5_n1.json 2> gpurun_out/scale_c5_n1.err; echo "bench c5 n1 exit $?" >> gpurun_out/status.txt
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_c5_n${n}.json 2> gpurun_out/scale_c5_n${n}.err; echo "bench c5 n$n exit $?" >> gpurun_out/status.txt
  fi
done
SLM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $NG --steps 6 --warmup 3 --no-cpu > gpurun_out/trace_n8.json 2> gpurun_out/trace_n8.txt; echo "trace exit $?" >> gpurun_out/status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu --workload c4 > gpurun_out/scale_c4_n8.json 2> gpurun_out/scale_c4_n8.err; echo "bench c4 n8 exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
