#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_launch" > gpurun_out/r4q_tests.txt 2>&1; echo "tests exit $?"
tail -30 gpurun_out/r4q_tests.txt
