#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 400 --timeout-method=thread --durations=8 2>&1 | tail -40 > gpurun_out/r2_s7_tests.log
tail -25 gpurun_out/r2_s7_tests.log | cut -c1-220
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s7_bench.json 2> gpurun_out/r2_s7_bench.err
tail -c 1500 gpurun_out/r2_s7_bench.json
