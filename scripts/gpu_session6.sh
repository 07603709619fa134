#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread --durations=15 2>&1 | tail -45 > gpurun_out/r2_s6_tests.log
tail -32 gpurun_out/r2_s6_tests.log | cut -c1-220
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
