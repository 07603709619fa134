#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29552 bench.py --gpus 4 --steps 10 --warmup 3 --no-micro > gpurun_out/r2_n4_bench.json 2> gpurun_out/r2_n4_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_n4_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("N=4 ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "fb", d.get("optimistic_exchange_fallbacks"), "nvlink", (d.get("nvlink") or {}).get("gbs_in"), "scat", (d.get("nvlink") or {}).get("scatter_kernel_ms_per_step"), "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else d.get("e2e_error"))
except Exception as e: print("ERR", e)
PY
tail -n 3 gpurun_out/r2_n4_bench.err | cut -c1-300
