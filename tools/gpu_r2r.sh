#!/bin/bash
# round 2, call R (N GPUs): 1-GPU bench (strong baselines), N-GPU bench line (observation-sharded headline + node-sharded strong
# configurations), the headline node-sharded, stage traces of both shardings
TAG=${1:-r2r}; N=${2:-8}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 exit $?"
bash tools/gpu_multi.sh ${TAG} $N > gpurun_out/${TAG}_multi.log 2>&1; echo "obs: $(head -1 gpurun_out/${TAG}_multi.log)"; tail -5 gpurun_out/${TAG}_bench_n$N.err
bash tools/gpu_multi.sh ${TAG}nodes $N --strong none --shard nodes > gpurun_out/${TAG}_multi_nodes.log 2>&1; echo "nodes: $(head -1 gpurun_out/${TAG}_multi_nodes.log)"
for sh in obs nodes; do SHARD=$sh timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/diag/trace_sharded.py cfg3 > gpurun_out/${TAG}_trace_cfg3_$sh.txt 2>&1; grep -A40 "ranks, rank 0" gpurun_out/${TAG}_trace_cfg3_$sh.txt; done
python - <<PY
import json
for tag, n in (("${TAG}", 1), ("${TAG}", $N), ("${TAG}nodes", $N)):
    try:
        d=json.load(open("gpurun_out/%s_bench_n%d.json" % (tag, n)))
    except Exception as e:
        print(tag, n, "no json", e); continue
    print("%s N=%d value %.3e step %.3f fit %.3f marg %.3f kernel %.3f e2e %.3f (%.3e) api %s prep %s" % (tag, n, d["value"], d["ms_per_step"], d["fit_ms"], d["marginal_ms"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("api_fit_marginals",{}).get("ms_median"), d["config"]["prep"]))
    print("    e2e phases", d["e2e"].get("host_phases_ms"))
    for k,v in (d.get("strong") or {}).items(): print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","kernel_ms","prep","api_fit_marginals_ms")}, (v.get("vs_1gpu") or {}).get("efficiency"), (v.get("global_quantile_parity") or {}).get("equal"))
PY
