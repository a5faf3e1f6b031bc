#!/bin/bash
# one-launch mode search, final: full GPU suite, small-model public-call breakdown, cfg1 / cfg2 bench lines
mkdir -p gpurun_out
TAG=r4j
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.txt 2>&1; echo "tests exit $?"
tail -5 gpurun_out/${TAG}_tests.txt
JP_MODE_TRACE=1 timeout 300 python tools/diag/small_api.py > gpurun_out/${TAG}_small_api.txt 2>&1; echo "small_api exit $?"
grep -v "^jp_mode (one" gpurun_out/${TAG}_small_api.txt | tail -12
grep "^jp_mode (one" gpurun_out/${TAG}_small_api.txt | sort | uniq -c | sort -k5 | awk 'NR%8==1' | head
JP_MODE_HOST=1 timeout 300 python tools/diag/small_api.py > gpurun_out/${TAG}_small_api_hostdriven.txt 2>&1; echo "small_api (host-driven) exit $?"
grep "total" gpurun_out/${TAG}_small_api_hostdriven.txt
for w in cfg1 cfg2; do timeout 300 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err; echo "bench $w exit $?"; done
python - <<'PY'
import json
for w in ("cfg1","cfg2"):
    try:
        j=json.loads(open("gpurun_out/r4j_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "step %.3f ms e2e %.3f api %.3f ms (min %.3f)"%(j["ms_per_step"], j["e2e"]["ms_per_step"], j["api_fit_marginals"]["ms_median"], j["api_fit_marginals"]["ms_min"]))
    except Exception as e: print(w, "ERR", e)
PY
