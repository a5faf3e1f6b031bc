#!/bin/bash
# round 2, call J (1 GPU): ncu --set full of the stage-4 and one-pass marginal kernels inside a cfg3 step (warm caches, as in situ)
TAG=${1:-r2j}; WL=${2:-cfg3}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong none --workload $WL"
$BENCH > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'jp_stage4_kernel|jp_marginal_onepass_kernel' -s 8 -c 2 -f -o gpurun_out/${TAG}_s45 $BENCH > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
