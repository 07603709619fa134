#!/bin/bash
mkdir -p gpurun_out
for T in 192 320 448 704; do
  LIB=build/libdbgb200_l2t$T.so; if [ $T = 448 ]; then LIB=dbg_assembly_b200/libdbgb200.so; fi
  DBG_B200_LIB=$PWD/$LIB timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-micro > gpurun_out/r2_s17_bench_$T.json 2> gpurun_out/r2_s17_bench_$T.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_s17_bench_$T.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("L2T=$T ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "clear", round(r["clear_ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "d2h", round(e["d2h_ms"],1))
PY
done
