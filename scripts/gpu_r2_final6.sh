#!/bin/bash
# Very last GPU seconds of round 2: one-shot check of the chi-square scan with batched loads (tests of tests/test_bow.py called
# directly + rates), then the pytest run of the same file if the budget still allows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 60 python scripts/chi2_quick_check.py > gpurun_out/chi2_quick_stdout.txt 2>&1; echo "quick exit $?"; tail -14 gpurun_out/chi2_quick_stdout.txt
timeout 60 python -m pytest tests/test_bow.py tests/test_device_mirrors_gpu.py -q -x -m gpu --timeout 50 > gpurun_out/pytest_bow_final6.txt 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_bow_final6.txt
exit 0
