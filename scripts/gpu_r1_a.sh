#!/bin/bash
# Round-1 GPU call A: micro-benchmarks (pipe rates, tcgen05 probe), parity tests, first bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
MB=slam-1_b200/csrc/microbench/bin
timeout 120 $MB/pipe_rates > gpurun_out/pipe_rates.txt 2>&1; echo "pipe_rates exit $?" >> gpurun_out/status.txt
timeout 120 $MB/tc_probe > gpurun_out/tc_probe.txt 2>&1; rc=$?; echo "tc_probe exit $rc" >> gpurun_out/status.txt
if [ $rc -ne 0 ]; then
  timeout 60 $MB/tc_probe 2048 128 256 > gpurun_out/tc_probe_swapped.txt 2>&1; echo "tc_probe swapped exit $?" >> gpurun_out/status.txt
fi
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "not tensor and not auto and not bmma" > gpurun_out/pytest_popc.txt 2>&1; echo "pytest popc exit $?" >> gpurun_out/status.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "tensor" > gpurun_out/pytest_tensor.txt 2>&1; echo "pytest tensor exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 5 --warmup 3 --variant popc > gpurun_out/bench_popc_c5.json 2> gpurun_out/bench_popc_c5.err; echo "bench popc exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 5 --warmup 3 --variant tensor --no-cpu > gpurun_out/bench_tensor_c5.json 2> gpurun_out/bench_tensor_c5.err; echo "bench tensor exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
tail -5 gpurun_out/pipe_rates.txt gpurun_out/tc_probe.txt gpurun_out/pytest_popc.txt gpurun_out/pytest_tensor.txt
