#!/bin/bash
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "tc or cfg or device_side or p2p" 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/${TAG}_pytest.log
for w in cfg3 cfg4 cfg5; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "bench $w exit $?"; tail -2 gpurun_out/${TAG}_bench_$w.err
done
python - <<PY
import json
for w in ("cfg3","cfg4","cfg5"):
    d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%w)); r=d["roofline"]
    print(w, "step %.3f fit %.3f marg %.3f kernel %.3f value %.3e e2e %.3f cycles/tile %.0f" % (d["ms_per_step"], d["fit_ms"], d["marginal_ms"], r["kernel_ms"], d["value"], d["e2e"]["ms_per_step"], r["tile_model"]["cycles_per_tile"]))
PY
