#!/bin/bash
# layout kernel v2 (warp-owned word pairs) vs v1; full GPU suite; C4 line
mkdir -p gpurun_out
for v in 2 1; do
  DBG_B200_LAYOUT_V=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-micro > gpurun_out/r2_s13_bench_v$v.json 2> gpurun_out/r2_s13_bench_v$v.err
  tail -n 2 gpurun_out/r2_s13_bench_v$v.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_s13_bench_v$v.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("layout v$v", "ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "e2e", round(e["ms_per_step"],2))
PY
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2_s13_tests.log
tail -6 gpurun_out/r2_s13_tests.log | cut -c1-300
timeout 300 python bench.py --steps 3 --warmup 1 --workload C4 > gpurun_out/r2_s13_bench_C4.json 2> gpurun_out/r2_s13_bench_C4.err
tail -n 2 gpurun_out/r2_s13_bench_C4.err | cut -c1-300; cut -c1-600 gpurun_out/r2_s13_bench_C4.json
