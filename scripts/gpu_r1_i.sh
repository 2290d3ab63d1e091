#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> gpurun_out/status.txt
timeout 300 python scripts/fixed_overhead.py > gpurun_out/fixed_overhead.txt 2>&1; echo "fixed overhead exit $?" >> gpurun_out/status.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 exit $?" >> gpurun_out/status.txt
# memcheck on a small parity subset (one sanitizer tool per call)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "ragged or empty or golden_bfmatcher" > gpurun_out/memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; cat gpurun_out/fixed_overhead.txt
