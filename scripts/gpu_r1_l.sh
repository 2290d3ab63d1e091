#!/bin/bash
# final c5 scaling line (NVLink exchange, default bench flags) on one 8-GPU box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
NG=$(nvidia-smi -L | wc -l)
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/final_c5_n1.json 2> gpurun_out/final_c5_n1.err; echo "n1 exit $?" >> gpurun_out/status.txt
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/final_c5_n${n}.json 2> gpurun_out/final_c5_n${n}.err; echo "n$n exit $?" >> gpurun_out/status.txt
  fi
done
cat gpurun_out/status.txt
