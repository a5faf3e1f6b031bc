#!/bin/bash
# N = 8 bench of the final tree as the driver launches it
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r5b_n8.json 2> gpurun_out/r5b_n8.err; echo "n8 exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r5b_n8.json").read().strip().splitlines()[-1])
print("N=8 value %.3e step %.4f kernel %.4f e2e %.4f (warmup %d) api %.3f"%(j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["ms_per_step"], j["e2e"]["warmup"], j["api_fit_marginals"]["ms_median"]), j["config"].get("exchanges"), j["config"].get("parallelism"))
print(j.get("strong"))
PY
