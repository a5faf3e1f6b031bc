#!/bin/bash
# round 2, call C (1 GPU): GPU tests + smoke on the fused stage 4 / side-stream prep, then the full default bench line (with strong)
TAG=${1:-r2c}
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
for w in cfg1 cfg2; do timeout 300 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_$w.json')); print('$w', 'step', d['ms_per_step'], 'api', d.get('api_fit_marginals'))"; done
