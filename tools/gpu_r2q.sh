#!/bin/bash
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -s -k "observation_sharded or raw_build or abs_is_not or p2p_sharded or device_side" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_new.log
echo "pytest(new) exit ${PIPESTATUS[0]}"; tail -30 gpurun_out/${TAG}_pytest_new.log
