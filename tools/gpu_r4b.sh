#!/bin/bash
# one-launch mode search: parity with the host-driven search + the small-model public call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mode or cfg1 or cfg2 or anova or covariance or family or points or raw_build" > gpurun_out/r4b_tests.txt 2>&1; echo "tests exit $?"
tail -15 gpurun_out/r4b_tests.txt
python tools/diag/small_api.py > gpurun_out/r4b_small_api.txt 2>&1; echo "small_api exit $?"
tail -12 gpurun_out/r4b_small_api.txt
