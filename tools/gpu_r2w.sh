#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
JP_NO_DENSITY_PREFETCH=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench_nopf.json 2> gpurun_out/${TAG}_bench_nopf.err; echo "bench(no prefetch) exit $?"
python - <<PY
import json
for f in ("bench","bench_nopf"):
    d=json.load(open("gpurun_out/${TAG}_%s.json"%f)); r=d["roofline"]
    print(f, "cfg3 step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f api %s" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median")), d["e2e"]["host_phases_ms"])
PY
