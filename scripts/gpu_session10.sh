#!/bin/bash
# pipelined export + pipelined submit: new tests, then the e2e number
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_export_pipe.py "tests/test_gpu_build.py::test_pipelined_submit_matches_oracle" "tests/test_gpu_build.py::test_blocks_and_subblocks_do_not_change_the_result" "tests/test_gpu_build.py::test_medium_synthetic_matches_oracle" -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/r2_s10_tests.log
tail -12 gpurun_out/r2_s10_tests.log | cut -c1-300
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA node\(s\)"; lscpu | grep -o "avx512f" | head -1
for mode in pipe plain; do
  if [ $mode = plain ]; then export DBG_B200_EXPORT=plain DBG_B200_PIPELINE=0; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-micro > gpurun_out/r2_s10_bench_$mode.json 2> gpurun_out/r2_s10_bench_$mode.err
  tail -n 2 gpurun_out/r2_s10_bench_$mode.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_s10_bench_$mode.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("$mode", "ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2))
print("  e2e", round(e["ms_per_step"],2), "h2d", round(e["h2d_ms"],2), "d2h", round(e["d2h_ms"],2), "build", round(e["build_ms"],2), "layout", round(e["layout_ms"],2), e.get("export"))
PY
done
