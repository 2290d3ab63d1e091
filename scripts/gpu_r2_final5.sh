#!/bin/bash
# Last GPU seconds of round 2: chi-square scan rates on dense (worst case) and sparse (what BoW.hist produces) histograms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=slam-1_b200
timeout 40 python scripts/masked_chi2_rates.py --chi2-only > gpurun_out/chi2_rates_final5.txt 2>&1; echo "rates exit $?"; tail -9 gpurun_out/chi2_rates_final5.txt
exit 0
