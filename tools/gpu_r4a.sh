#!/bin/bash
# small-model public-call breakdown
mkdir -p gpurun_out
python tools/diag/small_api.py > gpurun_out/r4a_small_api.txt 2>&1; echo "small_api exit $?"
cat gpurun_out/r4a_small_api.txt | tail -20
