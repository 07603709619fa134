#!/bin/bash
# GPU session 4: fail-fast parity subset, the N=1 bench line, ncu launch list + full capture of the three big kernels, whole suite
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_build.py -m gpu -x -q --timeout 120 -k "golden or dense or table_full or saturation" 2>&1 | tail -6 > gpurun_out/r2_s4_quick.log
tail -2 gpurun_out/r2_s4_quick.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s4_bench.json 2> gpurun_out/r2_s4_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_s4_bench.json")); r=d["roofline"]
    print("step", round(d["ms_per_step"],3), "insert", round(r["kernel_ms_per_step"],3), "build", round(r["build_kernels_ms_per_step"],3), "layout", round(r["layout_ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("value_t1"))
    print({k:(round(v.get("ms_per_step",0),2), round(v.get("value",0)/1e9,2)) if "ms_per_step" in v else v for k,v in d.get("other_workloads",{}).items()})
except Exception as e: print("bench ERR", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/r2_s4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-micro --no-other > gpurun_out/r2_s4_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_layout_clusters|k_insert_tuples|k_build' -c 3 -o gpurun_out/r2_s4_full python bench.py --steps 1 --warmup 1 --no-cpu --no-micro --no-other > gpurun_out/r2_s4_ncu2.log 2>&1
ls -la gpurun_out/r2_s4_full.ncu-rep 2>/dev/null
timeout 1300 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/r2_s4_tests.log
tail -6 gpurun_out/r2_s4_tests.log
