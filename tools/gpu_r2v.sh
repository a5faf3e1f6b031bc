#!/bin/bash
# round 2, final 1-GPU measurements: default bench line (cpu baseline, strong object), reference arm, cfg1 / cfg2 latency lines,
# stage traces, ncu launch list + full capture of the TC kernel, full capture of stage 4 / one-pass marginal exported as text
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "reference arm exit $?"
for w in cfg1 cfg2; do timeout 300 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err; done
for w in cfg3 cfg4 cfg5; do python tools/diag/trace_step.py $w > gpurun_out/${TAG}_trace_$w.txt 2>&1; done
bash tools/gpu_ncu.sh ${TAG} jp_glm_tc_kernel
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong none"
ncu --set full --clock-control none --cache-control none -k regex:'jp_stage4_kernel|jp_marginal_onepass_kernel' -s 8 -c 2 -f -o /tmp/${TAG}_s45 $BENCH > gpurun_out/${TAG}_ncu_s45.log 2>&1; echo "ncu s45 exit $?"
ncu -i /tmp/${TAG}_s45.ncu-rep --page raw --csv > gpurun_out/${TAG}_s45_raw.csv 2>/dev/null
ls -la gpurun_out | awk '{s+=$5} END {print "gpurun_out bytes", s}'
