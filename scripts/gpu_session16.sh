#!/bin/bash
# validation of the head: full GPU suite, the bench line, launch list, full ncu capture of scatter / insert / layout v2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2_s16_tests.log
tail -6 gpurun_out/r2_s16_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s16_bench.json 2> gpurun_out/r2_s16_bench.err
tail -n 3 gpurun_out/r2_s16_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_s16_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "clear", round(r["clear_ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "cpu", d["cpu_baseline"].get("value"))
print({k: (round(v["ms_per_step"],2) if "ms_per_step" in v else v) for k,v in d["other_workloads"].items()})
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-other --no-micro"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/r2_s16_launches.csv $CMD > gpurun_out/r2_s16_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_s16_launches.csv 40 | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_layout_clusters2|k_insert_tuples|k_build' -s 4 -c 3 -o gpurun_out/r2_s16_prof -f $CMD > gpurun_out/r2_s16_ncu2.log 2>&1
ls -la gpurun_out/r2_s16_prof.ncu-rep
