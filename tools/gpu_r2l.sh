#!/bin/bash
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "device_side or p2p_sharded" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_new.log
echo "pytest(new) exit ${PIPESTATUS[0]}"; tail -30 gpurun_out/${TAG}_pytest_new.log
python tools/diag/trace_e2e.py > gpurun_out/${TAG}_e2e_host.txt 2>&1; cat gpurun_out/${TAG}_e2e_host.txt | tail -16
JP_TC_DEVICE_DECISION=1 python tools/diag/trace_e2e.py > gpurun_out/${TAG}_e2e_dev.txt 2>&1; cat gpurun_out/${TAG}_e2e_dev.txt | tail -16
