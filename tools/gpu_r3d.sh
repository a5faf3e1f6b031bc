#!/bin/bash
# the 4-GPU bench line exactly as the driver launches it (default arguments)
TAG=${1:-r3d}; N=${2:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "exit $?"; tail -3 gpurun_out/${TAG}_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.err; echo "reference arm exit $?"; cut -c1-200 gpurun_out/${TAG}_ref_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json"))
print("N=%d value %.3e step %.3f fit %.3f marg %.3f kernel %.3f e2e %.3f (%.3e) api %s prep %s" % ($N, d["value"], d["ms_per_step"], d["fit_ms"], d["marginal_ms"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("api_fit_marginals",{}).get("ms_median"), d["config"]["prep"]))
print(d["config"]["parallelism"], d["scaling"], d["clocks"])
for k,v in (d.get("strong") or {}).items(): print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","kernel_ms","prep","api_fit_marginals_ms")}, (v.get("global_quantile_parity") or {}).get("equal"))
PY
