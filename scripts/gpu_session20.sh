#!/bin/bash
# refresh of the launch list and the scatter capture with the final library; one more e2e sample
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-other --no-micro"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --no-micro > gpurun_out/r2_s20_bench.json 2> gpurun_out/r2_s20_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_s20_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("ms", round(d["ms_per_step"],2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "h2d", round(e["h2d_ms"],1), "d2h", round(e["d2h_ms"],1))
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/r2_s20_launches.csv $CMD > gpurun_out/r2_s20_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_s20_launches.csv 30 | head -16
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build' -s 1 -c 1 -o gpurun_out/r2_s20_prof -f $CMD > gpurun_out/r2_s20_ncu2.log 2>&1
ls -la gpurun_out/r2_s20_prof.ncu-rep
