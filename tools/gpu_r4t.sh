#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smooth" > gpurun_out/r4u_tests.txt 2>&1; echo "tests exit $?"
tail -5 gpurun_out/r4u_tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4u_cfg3.json 2> gpurun_out/r4u_cfg3.err; echo "bench exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg1 > gpurun_out/r4u_cfg1.json 2> gpurun_out/r4u_cfg1.err; echo "bench cfg1 exit $?"
python - <<'PY'
import json
for w in ("cfg3","cfg1"):
    j=json.loads(open("gpurun_out/r4u_%s.json"%w).read().strip().splitlines()[-1])
    print(w, "step %.4f e2e %.4f api %.3f"%(j["ms_per_step"], j["e2e"]["ms_per_step"], j["api_fit_marginals"]["ms_median"]), j["api_fit_marginals"].get("smooth_cdf_marginal"))
PY
