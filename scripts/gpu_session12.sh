#!/bin/bash
# dbg_finish_export: tests, then e2e with the fused call vs the separate calls
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_gpu_build.py::test_finish_export_matches_oracle" "tests/test_gpu_build.py::test_pipelined_submit_matches_oracle" tests/test_export_pipe.py -m gpu -q -x --timeout 600 2>&1 | tail -30 > gpurun_out/r2_s12_tests.log
tail -14 gpurun_out/r2_s12_tests.log | cut -c1-400
for api in finish_export separate; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-micro --e2e-api $api > gpurun_out/r2_s12_bench_$api.json 2> gpurun_out/r2_s12_bench_$api.err
  tail -n 2 gpurun_out/r2_s12_bench_$api.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_s12_bench_$api.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("$api", "ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2))
print("  e2e", round(e["ms_per_step"],2), "h2d", round(e["h2d_ms"],2), "d2h", round(e["d2h_ms"],2), "build", round(e["build_ms"],2), "layout", round(e["layout_ms"],2), e.get("export"))
PY
done
