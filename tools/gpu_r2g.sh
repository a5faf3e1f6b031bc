#!/bin/bash
# round 2, call G (1 GPU): TMEM-read micro-benchmark on B200, host-side breakdown of the e2e step
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 120 tools/ubench/tmem_ld > gpurun_out/${TAG}_tmem_ld.txt 2>&1; echo "tmem_ld exit $?"; cat gpurun_out/${TAG}_tmem_ld.txt
timeout 300 python tools/diag/e2e_breakdown.py > gpurun_out/${TAG}_e2e.txt 2>&1; echo "e2e exit $?"; tail -12 gpurun_out/${TAG}_e2e.txt
