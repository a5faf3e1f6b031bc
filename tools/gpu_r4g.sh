#!/bin/bash
mkdir -p gpurun_out
JP_MODE_TRACE=1 timeout 300 python tools/diag/small_api.py > gpurun_out/r4m_small_api.txt 2>&1; echo "small_api exit $?"
grep -v "^jp_mode (one" gpurun_out/r4m_small_api.txt | tail -12
grep "^jp_mode (one" gpurun_out/r4m_small_api.txt | sort | uniq -c | sort -k5 | awk 'NR%8==1' | head
