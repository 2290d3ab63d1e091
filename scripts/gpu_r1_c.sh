#!/bin/bash
# Round-1 GPU call C: validate the 2-CTA tcgen05 kernel, A/B against the 1-CTA kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
MB=slam-1_b200/csrc/microbench/bin
timeout 120 $MB/tc_probe > gpurun_out/tc_probe.txt 2>&1; echo "tc_probe exit $?" >> gpurun_out/status.txt
timeout 600 python -m pytest tests -x -q -m gpu -k "tensor or auto" > gpurun_out/pytest_tensor.txt 2>&1; echo "pytest tensor exit $?" >> gpurun_out/status.txt
for wl in c5 c4 c3; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err; echo "bench $wl exit $?" >> gpurun_out/status.txt
  SLM_TC_1CTA=1 timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}_1cta.json 2> gpurun_out/bench_${wl}_1cta.err; echo "bench $wl 1cta exit $?" >> gpurun_out/status.txt
done
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
