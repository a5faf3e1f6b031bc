#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc or cfg3 or cfg4 or cfg5 or sharded or p2p or decision or glm" > gpurun_out/r4z_tests.txt 2>&1; echo "tests exit $?"
tail -3 gpurun_out/r4z_tests.txt
timeout 300 python tools/diag/trace_step.py > gpurun_out/r4z_trace.txt 2>&1; echo "trace exit $?"; tail -16 gpurun_out/r4z_trace.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4z_cfg3.json 2> gpurun_out/r4z_cfg3.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r4z_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 step %.4f fit %.4f kernel %.4f e2e %.4f api %.3f"%(j["ms_per_step"], j["fit_ms"], j["roofline"]["kernel_ms"], j["e2e"]["ms_per_step"], j["api_fit_marginals"]["ms_median"]))
PY
