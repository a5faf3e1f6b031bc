#!/bin/bash
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -s -k "smooth or covariance_matrix_and or log_density_points or raw_build" 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_new.log
echo "pytest(new) exit ${PIPESTATUS[0]}"; tail -45 gpurun_out/${TAG}_pytest_new.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --strong none > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["api_fit_marginals"])
PY
