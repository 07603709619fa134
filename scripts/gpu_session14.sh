#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_gpu_sharded_layout.py::test_sharded_build_merges_into_the_reference_table" -m gpu -q -x --timeout 600 -k "pull" 2>&1 | tail -25 > gpurun_out/r2_s14_tests.log
tail -12 gpurun_out/r2_s14_tests.log | cut -c1-300
