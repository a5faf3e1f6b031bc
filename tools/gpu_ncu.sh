#!/bin/bash
# ncu captures of one bench command: launch list + full capture of one kernel.  usage: bash tools/gpu_ncu.sh TAG KERNEL_REGEX [bench args]
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong none $*"
$BENCH > gpurun_out/${TAG}_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list exit $?"
$BENCH > gpurun_out/${TAG}_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
