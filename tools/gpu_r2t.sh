#!/bin/bash
# round 2, call T (1 GPU): stage-1 parity tests, grid build timing, ncu launch list of the grid builds
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "grid or sort or marginal_buffer or sorted" 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
echo "pytest exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/${TAG}_pytest.log
python tools/diag/grid_build.py > gpurun_out/${TAG}_grid_build.txt 2>&1; cat gpurun_out/${TAG}_grid_build.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_stage1_launches.csv python tools/diag/grid_build.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu exit $?"
