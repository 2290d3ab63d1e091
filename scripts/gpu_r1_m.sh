#!/bin/bash
# Round-1 closing evidence on 1 GPU: tests, smoke, reference arm, every bench workload, launch lists and full ncu
# captures of the kernels added in this session (frame kernel, chained tc2 kernel, frame refine).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/status.txt
for wl in c4 c3 c2 c1 h1; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err; echo "bench $wl exit $?" >> gpurun_out/status.txt
done
timeout 300 python scripts/cpu_overhead.py > gpurun_out/cpu_overhead.txt 2>&1; echo "cpu overhead exit $?" >> gpurun_out/status.txt
C1="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --workload c1"
C3="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --workload c3"
timeout 300 $C1 > gpurun_out/plain_c1.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c1.csv $C1 > gpurun_out/ncu_list_c1.log 2>&1; echo "ncu list c1 exit $?" >> gpurun_out/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_frame_kernel -s 3 -c 1 -o gpurun_out/prof_frame_c1 $C1 > gpurun_out/ncu_full_c1.log 2>&1; echo "ncu full frame exit $?" >> gpurun_out/status.txt
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3.csv $C3 > gpurun_out/ncu_list_c3.log 2>&1; echo "ncu list c3 exit $?" >> gpurun_out/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_tc2_kernel -s 2 -c 1 -o gpurun_out/prof_tc2_chain_c3 $C3 > gpurun_out/ncu_full_c3.log 2>&1; echo "ncu full tc2 chain exit $?" >> gpurun_out/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_refine_frame_kernel -s 2 -c 1 -o gpurun_out/prof_refine_frame_c3 $C3 > gpurun_out/ncu_full_c3r.log 2>&1; echo "ncu full refine frame exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
