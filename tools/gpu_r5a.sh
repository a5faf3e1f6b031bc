#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r5a_tests.txt 2>&1; echo "tests exit $?"
tail -3 gpurun_out/r5a_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r5a_smoke.txt 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r5a_smoke.txt
