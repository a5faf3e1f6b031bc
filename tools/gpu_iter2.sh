#!/bin/bash
# GPU tests + cfg3 / cfg4 / cfg5 bench lines (no launch lists): bash tools/gpu_iter2.sh TAG
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
summ() {
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$1.json")); r=d["roofline"]
print("$1", "value %.3e step %.3f fit %.3f kernel %.3f marg %.3f mode %.2f upload %.1f e2e %.3f NC %s fold %s trunc %.2e" % (d["value"], d["ms_per_step"], d["fit_ms"], r["kernel_ms"], d["marginal_ms"], d["mode_ms"], d["upload_ms"], d["e2e"]["ms_per_step"], d["tc_diagnostics"]["series_terms"], d["tc_diagnostics"].get("economised"), d["tc_diagnostics"]["truncation_bound"]))
PY
}
for w in cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err; echo "$w exit $?"; summ $w
done
