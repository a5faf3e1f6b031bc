#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mode or cfg1 or cfg2 or anova or covariance or family or points or raw_build" > gpurun_out/r4o_tests.txt 2>&1; echo "tests exit $?"
tail -15 gpurun_out/r4o_tests.txt
JP_MODE_TRACE=1 timeout 300 python tools/diag/small_api.py > gpurun_out/r4o_small_api.txt 2>&1; echo "small_api exit $?"
grep -v "^jp_mode (one" gpurun_out/r4o_small_api.txt | tail -12
grep "^jp_mode (one" gpurun_out/r4o_small_api.txt | sort | uniq -c | sort -k5 | awk 'NR%8==1' | head
