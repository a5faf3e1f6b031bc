#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mode or glm or sharded or p2p or cfg3" > gpurun_out/r4w_tests.txt 2>&1; echo "tests exit $?"
tail -3 gpurun_out/r4w_tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4w_cfg3.json 2> gpurun_out/r4w_cfg3.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r4w_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 step %.4f e2e %.4f api %.3f (min %.3f) mode_ms %.3f"%(j["ms_per_step"], j["e2e"]["ms_per_step"], j["api_fit_marginals"]["ms_median"], j["api_fit_marginals"]["ms_min"], j.get("mode_ms",-1)))
PY
