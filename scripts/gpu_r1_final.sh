#!/bin/bash
# Final round-1 evidence run on 1 GPU: tests, smoke, every bench workload, launch list + full ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/status.txt
for wl in c4 c3 c2 c1 h1 h4; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err; echo "bench $wl exit $?" >> gpurun_out/status.txt
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_tc2_final $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
