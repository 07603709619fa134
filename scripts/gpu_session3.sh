#!/bin/bash
# GPU session 3: quick parity first (fail fast), guard variants, then the whole suite -- every step under its own timeout
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_build.py -m gpu -x -q --timeout 180 -k "golden or dense or random_reads or table_full" 2>&1 | tail -15 > gpurun_out/r2_s3_quick.log
tail -3 gpurun_out/r2_s3_quick.log
for g in 0 1 2; do
  DBG_B200_LIB=$PWD/dbg_assembly_b200/variants/libdbg_guard$g.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro --no-other > gpurun_out/r2_s3_guard$g.json 2> gpurun_out/r2_s3_guard$g.err
done
python - <<'PY'
import json
for g in (0,1,2):
    try:
        d=json.load(open(f"gpurun_out/r2_s3_guard{g}.json"))
        r=d["roofline"]; print(g, round(d["ms_per_step"],3), "insert", round(r["kernel_ms_per_step"],3), "build", round(r["build_kernels_ms_per_step"],3), "layout", round(r["layout_ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(g, "ERR", e)
PY
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/r2_s3_tests.log
tail -12 gpurun_out/r2_s3_tests.log
