#!/bin/bash
# round 2, call A: GPU tests on the new default build, TMEM-read floor on B200, A/B of the epilogue variants (cfg3 + cfg5)
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
./tools/ubench/tmem_ld > gpurun_out/${TAG}_tmem_ld.txt 2>&1; echo "tmem_ld exit $?"; cat gpurun_out/${TAG}_tmem_ld.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/${TAG}_pytest.log
bash tools/gpu_abn.sh ${TAG}_cfg3 "--workload cfg3" jointposteriors.jl_b200/libjpcuda_base.so jointposteriors.jl_b200/libjpcuda_peek.so jointposteriors.jl_b200/libjpcuda_peek16.so jointposteriors.jl_b200/libjpcuda_peekc.so
bash tools/gpu_abn.sh ${TAG}_cfg5 "--workload cfg5 --steps 3" jointposteriors.jl_b200/libjpcuda_base.so jointposteriors.jl_b200/libjpcuda_peek.so jointposteriors.jl_b200/libjpcuda_peek16.so jointposteriors.jl_b200/libjpcuda_peekc.so
