#!/bin/bash
# host-decided vs device-decided series length, end to end (fresh upload per step), alternating on one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smooth" > gpurun_out/r4v_tests.txt 2>&1; echo "tests exit $?"
tail -3 gpurun_out/r4v_tests.txt
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4v_host_$rep.json 2> gpurun_out/r4v_host_$rep.err; echo "host $rep exit $?"
  JP_TC_DEVICE_DECISION=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4v_dev_$rep.json 2> gpurun_out/r4v_dev_$rep.err; echo "dev $rep exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r4v_*.json")):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "step %.4f e2e %.4f api %.3f"%(j["ms_per_step"], j["e2e"]["ms_per_step"], j["api_fit_marginals"]["ms_median"]), j["e2e"]["host_phases_ms"])
PY
