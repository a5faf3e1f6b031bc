#!/bin/bash
# round 2, call B: GPU tests (new Genz-Keister levels, full-size oracle parity), A/B of the epilogue tile loop
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -s 2>&1 | tail -150 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; grep -E "cfg[345]:|passed|failed|Error|error" gpurun_out/${TAG}_pytest.log | tail -30
bash tools/gpu_abn.sh ${TAG}_cfg3 "--workload cfg3" jointposteriors.jl_b200/libjpcuda_base.so jointposteriors.jl_b200/libjpcuda.so jointposteriors.jl_b200/libjpcuda_v2.so
bash tools/gpu_abn.sh ${TAG}_cfg4 "--workload cfg4 --steps 5" jointposteriors.jl_b200/libjpcuda.so jointposteriors.jl_b200/libjpcuda_v2.so
bash tools/gpu_abn.sh ${TAG}_cfg5 "--workload cfg5 --steps 3" jointposteriors.jl_b200/libjpcuda.so jointposteriors.jl_b200/libjpcuda_v2.so
