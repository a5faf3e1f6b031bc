#!/bin/bash
# One gpurun call: all GPU tests, smoke, the default bench line and the (cold-cache) ncu launch list of the same bench command.
# usage: tools/gpu_retry.sh 900 "bash tools/gpu_full.sh <tag>"
TAG=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -s --durations=6 2>&1 | grep -v "^E    +" | tail -80 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_smoke.log; cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 200 ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
