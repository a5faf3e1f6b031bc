#!/bin/bash
# round 2, call B: GPU tests (new Genz-Keister levels, full-size oracle parity), A/B of the epilogue tile loop
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -s 2>&1 | tail -150 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; grep -E "cfg[345]:|passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -30
bash tools/gpu_abn.sh ${TAG}_cfg3 "--workload cfg3" jointposteriors.jl_b200/libjpcuda_base.so jointposteriors.jl_b200/libjpcuda.so jointposteriors.jl_b200/libjpcuda_v2.so
bash tools/gpu_abn.sh ${TAG}_cfg4 "--workload cfg4 --steps 5" jointposteriors.jl_b200/libjpcuda.so jointposteriors.jl_b200/libjpcuda_v2.so
bash tools/gpu_abn.sh ${TAG}_cfg5 "--workload cfg5 --steps 3" jointposteriors.jl_b200/libjpcuda.so jointposteriors.jl_b200/libjpcuda_v2.so
# attribution of the host round trip of the series-length decision (diagnostic env: skip the bounds read-back)
JP_TC_ASSUME=4,1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/${TAG}_assume.json 2> gpurun_out/${TAG}_assume.err
python - <<PY
import json
for n in ("${TAG}_cfg3_libjpcuda","${TAG}_assume"):
    d=json.load(open("gpurun_out/%s.json"%n)); r=d["roofline"]
    print(n, "step %.3f fit %.3f marg %.3f kernel %.3f e2e %.3f api %s" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median")))
PY
