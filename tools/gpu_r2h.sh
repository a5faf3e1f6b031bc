#!/bin/bash
# round 2, call H (N GPUs): 1-GPU bench (strong baseline for the same boot) then the N-GPU bench line
TAG=${1:-r2h}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 exit $?"
bash tools/gpu_multi.sh ${TAG} $N > gpurun_out/${TAG}_multi.log 2>&1; head -c 600 gpurun_out/${TAG}_multi.log
python - <<PY
import json
for n in (1, $N):
    d=json.load(open("gpurun_out/${TAG}_bench_n%d.json" % n))
    print("N=%d value %.3e step %.3f fit %.3f marg %.3f kernel %.3f e2e %.3f (%.3e) api %s prep %s" % (n, d["value"], d["ms_per_step"], d["fit_ms"], d["marginal_ms"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("api_fit_marginals",{}).get("ms_median"), d["config"]["prep"]))
    for k,v in (d.get("strong") or {}).items(): print("   ", k, {a:b for a,b in v.items() if a not in ("workload","api_what","unit","steps","warmup","scaling","nodes","obs","d")})
PY
