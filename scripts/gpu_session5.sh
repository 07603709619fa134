#!/bin/bash
# GPU session 5: find the hanging sharded-layout test (thread-method timeouts dump the Python stack even when the main thread
# sits in a C call), time the new tests, kernel variants
mkdir -p gpurun_out
timeout 200 python -m pytest "tests/test_gpu_sharded_layout.py::test_sharded_layout_on_dense_and_tiny_tables" -m gpu -q --timeout 40 --timeout-method=thread 2>&1 | tail -60 > gpurun_out/r2_s5_hang.log
grep -E "passed|failed|Timeout|File \"|line " gpurun_out/r2_s5_hang.log | tail -25
timeout 900 python -m pytest tests/test_gpu_sharded_layout.py -m gpu -q --timeout 60 --timeout-method=thread --durations=12 --deselect "tests/test_gpu_sharded_layout.py::test_sharded_layout_on_dense_and_tiny_tables" 2>&1 | tail -40 > gpurun_out/r2_s5_sharded.log
tail -22 gpurun_out/r2_s5_sharded.log
for v in main lt128 scat4; do
  L=$PWD/dbg_assembly_b200/variants/libdbg_$v.so; [ $v = main ] && L=$PWD/dbg_assembly_b200/libdbgb200.so
  E=""; [ $v = scat4 ] && E="DBG_B200_STAGE_CAP=1024"
  env $E DBG_B200_LIB=$L timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro --no-other > gpurun_out/r2_s5_$v.json 2> gpurun_out/r2_s5_$v.err
done
DBG_B200_STAGE_CAP=1024 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro --no-other > gpurun_out/r2_s5_cap1024.json 2> gpurun_out/r2_s5_cap1024.err
python - <<'PY'
import json
for g in ("main","lt128","scat4","cap1024"):
    try:
        d=json.load(open(f"gpurun_out/r2_s5_{g}.json")); r=d["roofline"]
        print(g, round(d["ms_per_step"],3), "insert", round(r["kernel_ms_per_step"],3), "build", round(r["build_kernels_ms_per_step"],3), "layout", round(r["layout_ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(g, "ERR", e)
PY
