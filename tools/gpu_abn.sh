#!/bin/bash
# A/B/n builds of the library on the bench: bash tools/gpu_abn.sh TAG "bench args" libA.so libB.so ...
TAG=$1; ARGS=$2; shift 2
mkdir -p gpurun_out
for L in "$@"; do
  n=$(basename $L .so)
  JPCUDA_LIB=$PWD/$L timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none $ARGS > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
  echo "$n exit $?"; tail -2 gpurun_out/${TAG}_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$n.json"))
r=d["roofline"]
print("$n", "step_ms %.3f fit_ms %.3f kernel_ms %.3f pairs/s %.3e epi_frac %.3f e2e_ms %.3f" % (d["ms_per_step"], d["fit_ms"], r["kernel_ms"], r["kernel_pairs_per_s"], r.get("epilogue",{}).get("frac",0), d["e2e"]["ms_per_step"]))
PY
done
