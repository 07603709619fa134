#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_seedidx.py -m gpu -q -x --timeout 300 2>&1 | tail -30 > gpurun_out/r2_s8_seed.log
tail -25 gpurun_out/r2_s8_seed.log | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other > gpurun_out/r2_s8_bench.json 2> gpurun_out/r2_s8_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_s8_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "e2e", d["e2e"]["ms_per_step"])
PY
