#!/bin/bash
# N = 2 bench as the driver launches it (adaptive e2e warm-up agreed over the ranks, fallback plumbing)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4x_n2.json 2> gpurun_out/r4x_n2.err; echo "n2 exit $?"
tail -c 600 gpurun_out/r4x_n2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r4x_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.3e step %.4f e2e %.4f (warmup %d) api %.3f"%(j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["e2e"]["warmup"], j["api_fit_marginals"]["ms_median"]), j["config"].get("exchanges"), j["e2e"]["per_step_ms"])
PY
