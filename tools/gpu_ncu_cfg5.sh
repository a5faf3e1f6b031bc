#!/bin/bash
# ncu --set full captures of the cfg5-sized kernels (one launch each): bash tools/gpu_ncu_cfg5.sh TAG
TAG=$1
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5"
$BENCH > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'jp_glm_tc_kernel|tc_obs_prep_kernel|jp_glm_partials_kernel|tc_split_x_kernel|tc_fold_kernel' -s 20 -c 8 -f -o gpurun_out/${TAG}_prof $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
