#!/bin/bash
# round 2, call F (1 GPU): GPU tests + smoke, the full default bench line, then the ncu launch list and one full capture of the TC kernel
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); r=d["roofline"]
print("cfg3 step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f api %s launches %d" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median"), d["gpu_launches"]))
for k,v in (d.get("strong") or {}).items(): print(k, {a:b for a,b in v.items() if a not in ("workload","api_what")})
print(d.get("cpu_baseline"))
PY
for w in cfg3 cfg4 cfg5; do python tools/diag/trace_step.py $w > gpurun_out/${TAG}_trace_$w.txt 2>&1; tail -20 gpurun_out/${TAG}_trace_$w.txt; done
bash tools/gpu_ncu.sh ${TAG} jp_glm_tc_kernel
