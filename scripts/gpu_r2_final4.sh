#!/bin/bash
# Last seconds of the round's GPU budget: the bag-of-words tests (chi-square term short path, count ranges) and the scan rates.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 75 python -m pytest tests/test_bow.py -q -x -m gpu --timeout 60 > gpurun_out/pytest_bow_final4.txt 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_bow_final4.txt
timeout 30 python scripts/masked_chi2_rates.py > gpurun_out/masked_chi2_rates_v3.txt 2>&1; echo "rates exit $?"; tail -5 gpurun_out/masked_chi2_rates_v3.txt
exit 0
