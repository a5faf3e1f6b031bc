#!/bin/bash
# final 1-GPU validation of the tree: full GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r4y_tests.txt 2>&1; echo "tests exit $?"
tail -3 gpurun_out/r4y_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4y_smoke.txt 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r4y_smoke.txt
timeout 600 python bench.py > gpurun_out/r4y_bench.json 2> gpurun_out/r4y_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4y_ref.json 2> gpurun_out/r4y_ref.err; echo "ref exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r4y_bench.json").read().strip().splitlines()[-1])
print("value %.3e step %.4f kernel %.4f e2e %.4f (warmup %d) api %.3f launches %s"%(j["value"], j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["ms_per_step"], j["e2e"]["warmup"], j["api_fit_marginals"]["ms_median"], j["gpu_launches"]))
print(j["roofline"]); print(j["cpu_baseline"]); print(j["api_fit_marginals"].get("smooth_cdf_marginal"))
r=json.loads(open("gpurun_out/r4y_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"])
PY
