#!/bin/bash
mkdir -p gpurun_out
for sh in 21; do
DBG_B200_PART_SHIFT=$sh timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-micro --e2e-api separate > gpurun_out/r2_s18_bench_sh$sh.json 2> gpurun_out/r2_s18_bench_sh$sh.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_s18_bench_sh$sh.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("N=1 shift $sh ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2))
PY
done
