#!/bin/bash
# Full validation of the current tree in one gpurun call: GPU tests, smoke, bench lines (cfg3 default, cfg1 API latency),
# warm-cache per-kernel durations of one bench step.
TAG=${1:-r05c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short --durations=8 2>&1 | tail -80 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_cfg1.json 2> gpurun_out/${TAG}_cfg1.err
echo "cfg1 exit $?" >> gpurun_out/${TAG}_cfg1.err
tail -25 gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_smoke.log; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_cfg1.err
timeout 200 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_warm_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
