#!/bin/bash
# round 2, call K (1 GPU): the new tests (device-side decision, emulated P2P ranks) first, then the whole GPU suite, bench, trace
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "device_side or p2p_sharded" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_new.log
echo "pytest(new) exit ${PIPESTATUS[0]}"; tail -30 gpurun_out/${TAG}_pytest_new.log
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
JP_TC_DEVICE_DECISION=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench_dev.json 2> gpurun_out/${TAG}_bench_dev.err; echo "bench(dev decision) exit $?"; tail -3 gpurun_out/${TAG}_bench_dev.err
python - <<PY
import json
for f in ("bench", "bench_dev"):
    d=json.load(open("gpurun_out/${TAG}_%s.json" % f)); r=d["roofline"]
    print(f, "cfg3 step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f api %s launches %d" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], d.get("api_fit_marginals",{}).get("ms_median"), d["gpu_launches"]))
PY
for w in cfg3; do python tools/diag/trace_step.py $w > gpurun_out/${TAG}_trace_$w.txt 2>&1; tail -13 gpurun_out/${TAG}_trace_$w.txt; JP_TC_DEVICE_DECISION=1 python tools/diag/trace_step.py $w > gpurun_out/${TAG}_trace_dev_$w.txt 2>&1; tail -13 gpurun_out/${TAG}_trace_dev_$w.txt; done
