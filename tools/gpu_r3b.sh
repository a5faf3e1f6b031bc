#!/bin/bash
# round 2, last 1-GPU validation of the final code: whole GPU suite, smoke, default bench line, reference arm, launch list + full capture of the TC kernel
TAG=${1:-r3b}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "reference arm exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); r=d["roofline"]
print("cfg3 step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f api %s launches %d grid_build %.2f" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median"), d["gpu_launches"], d["grid_build_ms"]))
for k,v in (d.get("strong") or {}).items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","kernel_ms","api_fit_marginals_ms")})
PY
bash tools/gpu_ncu.sh ${TAG} jp_glm_tc_kernel
ls -la gpurun_out | awk '{s+=$5} END {print "gpurun_out bytes", s}'
