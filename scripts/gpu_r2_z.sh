#!/bin/bash
# Re-validation after the chi-square stack fix: the four touched test files.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 500 python -m pytest tests/test_bow.py tests/test_host_path_gpu.py tests/test_batched_chain_gpu.py tests/test_masked_gpu.py -q -m gpu --timeout 200 > gpurun_out/pytest_gpu_z.txt 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_z.txt | cut -c1-300
exit 0
