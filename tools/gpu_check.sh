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
