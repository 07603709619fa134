#!/bin/bash
# GPU session 2: sanitizer on the by_slice failure, guard variants, new layout kernel, full tests
mkdir -p gpurun_out
(timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_sharded_layout.py -m gpu -x -q -k "31-False-True-2" 2>&1 | grep -v "^$" | head -60) > gpurun_out/r2_s2_sanitizer.log 2>&1
for g in 0 1 2; do
  DBG_B200_LIB=$PWD/dbg_assembly_b200/variants/libdbg_guard$g.so python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/r2_s2_guard$g.json 2> gpurun_out/r2_s2_guard$g.err
done
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_s2_tests.log
python - <<'PY'
import json
for g in (0,1,2):
    try:
        d=json.load(open(f"gpurun_out/r2_s2_guard{g}.json"))
        r=d["roofline"]; print(g, round(d["ms_per_step"],3), "insert", round(r["kernel_ms_per_step"],3), "build", round(r["build_kernels_ms_per_step"],3), "layout", round(r["layout_ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(g, "ERR", e)
PY
tail -8 gpurun_out/r2_s2_tests.log; head -40 gpurun_out/r2_s2_sanitizer.log
