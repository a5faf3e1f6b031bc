#!/bin/bash
TAG=${1:-r3c}
mkdir -p gpurun_out
export JPCUDA_LIB=$PWD/jointposteriors.jl_b200/libjpcuda_w24.so
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "tc or cfg3 or device_side or p2p" 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
echo "pytest(w24) exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/${TAG}_pytest.log
for w in cfg3 cfg5; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none --workload $w > gpurun_out/${TAG}_w24_$w.json 2> gpurun_out/${TAG}_w24_$w.err; echo "bench w24 $w exit $?"; tail -2 gpurun_out/${TAG}_w24_$w.err
done
unset JPCUDA_LIB
for w in cfg3 cfg5; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none --workload $w > gpurun_out/${TAG}_w32_$w.json 2> gpurun_out/${TAG}_w32_$w.err; echo "bench w32 $w exit $?"
done
python - <<PY
import json
for v in ("w24","w32"):
  for w in ("cfg3","cfg5"):
    d=json.load(open("gpurun_out/${TAG}_%s_%s.json"%(v,w))); r=d["roofline"]
    print(v, w, "step %.3f kernel %.3f value %.3e" % (d["ms_per_step"], r["kernel_ms"], d["value"]))
PY
