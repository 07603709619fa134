#!/bin/bash
# round 2, re-entry: full GPU suite + bench line + launch list on the restored head
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2_s9_tests.log
tail -8 gpurun_out/r2_s9_tests.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s9_bench.json 2> gpurun_out/r2_s9_bench.err
tail -n 3 gpurun_out/r2_s9_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_s9_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "e2e", d["e2e"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
PY
