#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
for wl in h1 h4 c5; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err; echo "bench $wl exit $?" >> gpurun_out/status.txt
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --workload h1"
timeout 600 $CMD > gpurun_out/plain_h1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_stream_kernel -s 3 -c 1 -o gpurun_out/prof_stream_h1 $CMD > gpurun_out/ncu_h1.log 2>&1; echo "ncu h1 exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
