#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a bench line, and the ncu launch list of the same bench command.
# usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -400 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log | tail -5; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "${NCU:-0}" = "1" ]; then
  BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong none --path ${NCU_PATH:-auto}"
  $BENCH > gpurun_out/${TAG}_ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_list.log 2>&1
  echo "ncu list exit $?"
  $BENCH > gpurun_out/${TAG}_ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-jp_fit_nodes_kernel} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
