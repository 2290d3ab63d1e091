#!/bin/bash
# Validation of the 8-lane-group chi-square scan and the masked kernel's grid; their rates.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 300 python -m pytest tests/test_bow.py tests/test_masked_gpu.py -q -m gpu --timeout 200 > gpurun_out/pytest_gpu_z2.txt 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_z2.txt | cut -c1-300
timeout 200 python scripts/masked_chi2_rates.py > gpurun_out/masked_chi2_rates_v2.txt 2>&1; echo "rates exit $?"; cat gpurun_out/masked_chi2_rates_v2.txt
exit 0
