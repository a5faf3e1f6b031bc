#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x --durations=5 -k "cfg4_full or cfg5_full" 2>&1 | tail -60 > gpurun_out/r05b_pytest.log
tail -30 gpurun_out/r05b_pytest.log
timeout 150 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -k regex:"bins|moments|prep|partials|expw|reduce_max|scale|fold|finish|combine|glm_tc" -c 300 --csv --log-file gpurun_out/r05b_warm_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r05b_ncu.log 2>&1
echo ncu exit $?
