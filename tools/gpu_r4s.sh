#!/bin/bash
mkdir -p gpurun_out
for sl in 8192 6144 3072; do
  JP_BINS_SLICE=$sl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4s_cfg3_$sl.json 2> gpurun_out/r4s_cfg3_$sl.err; echo "cfg3 slice $sl exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r4s_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step %.4f fit %.4f marg %.4f kernel %.4f e2e %.4f"%(j["ms_per_step"], j["fit_ms"], j["marginal_ms"], j["roofline"]["kernel_ms"], j["e2e"]["ms_per_step"]))
    except Exception as e: print(f,"ERR",e)
PY
