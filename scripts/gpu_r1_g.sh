#!/bin/bash
# full parity suite (all variants incl. bmma, chunked host path), A/B of the three variants with ncu counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" >> gpurun_out/status.txt
for v in popc bmma; do
  timeout 600 python bench.py --steps 3 --warmup 3 --variant $v --no-cpu --e2e-steps 1 > gpurun_out/bench_c5_$v.json 2> gpurun_out/bench_c5_$v.err; echo "bench c5 $v exit $?" >> gpurun_out/status.txt
done
for wl in c4 c3 c2 c1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err; echo "bench $wl exit $?" >> gpurun_out/status.txt
done
# ncu: issue-slot / pipe counters for the A/B (one launch each), after the plain runs above exited 0
CMDP="python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --variant popc"
CMDB="python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --variant bmma"
timeout 600 $CMDP > gpurun_out/plain_popc.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_popc_kernel -s 3 -c 1 -o gpurun_out/prof_popc $CMDP > gpurun_out/ncu_popc.log 2>&1; echo "ncu popc exit $?" >> gpurun_out/status.txt
timeout 600 $CMDB > gpurun_out/plain_bmma.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_bmma_kernel -s 3 -c 1 -o gpurun_out/prof_bmma $CMDB > gpurun_out/ncu_bmma.log 2>&1; echo "ncu bmma exit $?" >> gpurun_out/status.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
