#!/bin/bash
# Round-1 GPU call D: ncu launch list + full capture of the 2-CTA tcgen05 kernel (c5), after a plain run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_tc2 $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt
