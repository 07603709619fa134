#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --workload C3 --no-micro > gpurun_out/r2_n2f_bench_C3.json 2> gpurun_out/r2_n2f_bench_C3.err
timeout 300 $TR --master-port 29541 tests/multigpu_check.py --exchange pull --K 63 > gpurun_out/r2_n2f_check63.json 2> gpurun_out/r2_n2f_check63.err; tail -n 1 gpurun_out/r2_n2f_check63.json | cut -c1-300
python - <<'PY'
import json
for f in ["r2_n2f_bench_C3.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "fb", d.get("optimistic_exchange_fallbacks"), "nvlink", (d.get("nvlink") or {}).get("gbs_in"), "scat", (d.get("nvlink") or {}).get("scatter_kernel_ms_per_step"), "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else d.get("e2e_error"))
    except Exception as e: print(f, "ERR", e)
PY
