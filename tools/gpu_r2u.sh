#!/bin/bash
# round 2, final 1-GPU validation: whole GPU suite, smoke, the default bench line (cpu baseline, strong object), the reference arm,
# ncu launch list and full captures (TC kernel; stage 4 / one-pass marginal with warm caches)
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); r=d["roofline"]
print("cfg3 step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f api %s launches %d grid_build %.2f" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median"), d["gpu_launches"], d["grid_build_ms"]))
print({k: r[k] for k in ("frac","achieved","traffic")}, r["epilogue"]["frac"], {k: round(v,3) for k,v in r["tile_model"].items() if isinstance(v,float)})
print(d.get("cpu_baseline")); print(d["api_fit_marginals"].get("smooth_cdf_marginal"))
for k,v in (d.get("strong") or {}).items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","kernel_ms","api_fit_marginals_ms")})
PY
for w in cfg1 cfg2; do timeout 300 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_$w.json')); print('$w', 'step', d['ms_per_step'], 'api', {k:v for k,v in d.get('api_fit_marginals',{}).items() if k in ('ms_median','ms_min')}, d['api_fit_marginals'].get('smooth_cdf_marginal',{}).get('ms'), d['api_fit_marginals'].get('smooth_cdf_marginal',{}).get('converged'))"; done
for w in cfg3 cfg4 cfg5; do python tools/diag/trace_step.py $w > gpurun_out/${TAG}_trace_$w.txt 2>&1; done
bash tools/gpu_ncu.sh ${TAG} jp_glm_tc_kernel
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong none"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'jp_stage4_kernel|jp_marginal_onepass_kernel' -s 8 -c 2 -f -o gpurun_out/${TAG}_s45 $BENCH > gpurun_out/${TAG}_ncu_s45.log 2>&1; echo "ncu s45 exit $?"
