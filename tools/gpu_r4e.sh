#!/bin/bash
# ncu source-level view of the one-launch mode search (cfg1: binomial mixture d=3 = launches 0..12; cfg2: eight schools d=10 after them)
mkdir -p gpurun_out /tmp/ncu
for pair in cfg1:6 cfg2:19; do
  tag=${pair%%:*}; skip=${pair##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:jp_mode_dev_kernel --launch-skip $skip -c 1 \
    -o /tmp/ncu/r4e_$tag -f python tools/diag/small_api.py > gpurun_out/r4e_ncu_$tag.log 2>&1; echo "ncu $tag exit $?"
  ncu -i /tmp/ncu/r4e_$tag.ncu-rep --page raw --csv > gpurun_out/r4e_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/r4e_$tag.ncu-rep --page source --csv > gpurun_out/r4e_${tag}_source.csv 2>/dev/null
  ls -la /tmp/ncu/ gpurun_out/r4e_${tag}_*
done
