#!/bin/bash
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); r=d["roofline"]
print("cfg3 step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f api %s" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median")), d["e2e"]["host_phases_ms"])
PY
for w in cfg3 cfg4; do python tools/diag/trace_step.py $w > gpurun_out/${TAG}_trace_$w.txt 2>&1; tail -13 gpurun_out/${TAG}_trace_$w.txt; done
