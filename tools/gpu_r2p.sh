#!/bin/bash
# round 2, call P (2 GPUs): new 1-GPU tests, whole GPU suite, then the 2-GPU bench line in observation-sharded (default) and node-sharded mode
TAG=${1:-r2p}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "observation_sharded or raw_build or abs_is_not or p2p_sharded or device_side" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_new.log
echo "pytest(new) exit ${PIPESTATUS[0]}"; tail -30 gpurun_out/${TAG}_pytest_new.log
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 exit $?"
bash tools/gpu_multi.sh ${TAG} $N --strong none > gpurun_out/${TAG}_multi.log 2>&1; echo "obs: $(head -1 gpurun_out/${TAG}_multi.log)"; tail -5 gpurun_out/${TAG}_bench_n$N.err
bash tools/gpu_multi.sh ${TAG}nodes $N --strong none --shard nodes > gpurun_out/${TAG}_multi_nodes.log 2>&1; echo "nodes: $(head -1 gpurun_out/${TAG}_multi_nodes.log)"
python - <<PY
import json
for tag, n in (("${TAG}", 1), ("${TAG}", $N), ("${TAG}nodes", $N)):
    try:
        d=json.load(open("gpurun_out/%s_bench_n%d.json" % (tag, n)))
    except Exception as e:
        print(tag, n, "no json", e); continue
    print("%s N=%d value %.3e step %.3f fit %.3f marg %.3f kernel %.3f e2e %.3f (%.3e) api %s prep %s" % (tag, n, d["value"], d["ms_per_step"], d["fit_ms"], d["marginal_ms"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("api_fit_marginals",{}).get("ms_median"), d["config"]["prep"]))
    print("    e2e phases", d["e2e"].get("host_phases_ms"))
PY
