#!/bin/bash
# Round-1 GPU call B: full parity suite, bench lines for every workload, ncu launch list + full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
MB=slam-1_b200/csrc/microbench/bin
timeout 120 $MB/tc_probe > gpurun_out/tc_probe.txt 2>&1; echo "tc_probe exit $?" >> gpurun_out/status.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
for wl in c5 c4 c3 c2 c1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err; echo "bench $wl exit $?" >> gpurun_out/status.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --workload c1 --variant popc --no-cpu > gpurun_out/bench_c1_popc.json 2>> gpurun_out/bench_c1.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --variant popc --no-cpu > gpurun_out/bench_c2_popc.json 2>> gpurun_out/bench_c2.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/status.txt
# ncu: launch list (cold-cache, serialised: compare shares) then one full capture of the dominant kernel
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc_kernel -s 3 -c 1 -o gpurun_out/prof_tc \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
