#!/bin/bash
# attribute the TC kernel's time: product vs no-series-arithmetic vs no-TMEM-loads (profiling build only)
# build it first (here, no GPU needed): make -C jointposteriors.jl_b200/csrc BUILD=build_prof OUT=../libjpcuda_prof.so EXTRA=-DJP_TC_PROFILING_MODES -j8
TAG=$1; shift
for m in ${MODES:-0 1 2 3 4}; do
  JPCUDA_LIB=$PWD/jointposteriors.jl_b200/libjpcuda_prof.so JP_TC_DEBUG_MODE=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_mode$m.json 2> gpurun_out/${TAG}_mode$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_mode$m.json")); r=d["roofline"]
print("mode $m kernel_ms %.3f pairs/s %.3e" % (r["kernel_ms"], r["kernel_pairs_per_s"]))
PY
done
