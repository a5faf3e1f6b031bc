#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
(echo host; python tools/diag/e2e_loop.py; echo dev; JP_TC_DEVICE_DECISION=1 python tools/diag/e2e_loop.py; echo "host keep"; KEEP_RESIDENT=1 python tools/diag/e2e_loop.py; echo "dev keep"; KEEP_RESIDENT=1 JP_TC_DEVICE_DECISION=1 python tools/diag/e2e_loop.py) > gpurun_out/${TAG}_e2e_loop.txt 2>&1
cat gpurun_out/${TAG}_e2e_loop.txt
