#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
python tools/diag/trace_step.py cfg3 > gpurun_out/${TAG}_trace_cfg3.txt 2>&1; cat gpurun_out/${TAG}_trace_cfg3.txt | tail -20
python tools/diag/e2e_breakdown.py > gpurun_out/${TAG}_e2e_breakdown.txt 2>&1; tail -12 gpurun_out/${TAG}_e2e_breakdown.txt
