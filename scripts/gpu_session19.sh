#!/bin/bash
# end-of-round validation of the head on one GPU: full suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2_s19_tests.log
tail -6 gpurun_out/r2_s19_tests.log | cut -c1-300
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_s19_bench.json 2> gpurun_out/r2_s19_bench.err
tail -n 3 gpurun_out/r2_s19_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_s19_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "clear", round(r["clear_ms_per_step"],2), "e2e", round(e["ms_per_step"],2), "d2h", round(e["d2h_ms"],1), "cpu", d["cpu_baseline"].get("value"))
print({k: (round(v["ms_per_step"],2) if "ms_per_step" in v else v) for k,v in d["other_workloads"].items()})
PY
