#!/bin/bash
# One gpurun call of a development iteration: GPU tests, cfg3 bench (+ no-fold A/B), cfg5 bench + its launch list.
# usage: bash tools/gpu_iter.sh TAG
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
summ() {
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$1.json")); r=d["roofline"]
print("$1", "step %.3f fit %.3f kernel %.3f marg %.3f mode %.2f upload %.1f e2e %.3f NC %s fold %s trunc %.2e" % (d["ms_per_step"], d["fit_ms"], r["kernel_ms"], d["marginal_ms"], d["mode_ms"], d["upload_ms"], d["e2e"]["ms_per_step"], d["tc_diagnostics"]["series_terms"], d["tc_diagnostics"].get("economised"), d["tc_diagnostics"]["truncation_bound"]))
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_cfg3.json 2> gpurun_out/${TAG}_cfg3.err; echo "cfg3 exit $?"; summ cfg3
JP_TC_NO_FOLD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg3nofold.json 2> gpurun_out/${TAG}_cfg3nofold.err; echo "cfg3 nofold exit $?"; summ cfg3nofold
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err; echo "cfg5 exit $?"; summ cfg5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_cfg5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/${TAG}_cfg5_ncu.log 2>&1; echo "ncu $?"
