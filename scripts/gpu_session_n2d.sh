#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tests/multigpu_check.py --exchange pull > gpurun_out/r2_n2d_check.json 2> gpurun_out/r2_n2d_check.err; tail -n 1 gpurun_out/r2_n2d_check.json | cut -c1-400
timeout 300 $TR --master-port 29542 tests/multigpu_check.py --exchange pull --K 63 > gpurun_out/r2_n2d_check63.json 2> gpurun_out/r2_n2d_check63.err; tail -n 1 gpurun_out/r2_n2d_check63.json | cut -c1-400
for X in pull peer; do
timeout 400 $TR --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 --sub-blocks 1 --exchange $X --no-micro > gpurun_out/r2_n2d_bench_$X.json 2> gpurun_out/r2_n2d_bench_$X.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_n2d_bench_$X.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$X ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "clear", round(r["clear_ms_per_step"],2), "fb", d.get("optimistic_exchange_fallbacks"), "scat", d.get("nvlink",{}).get("scatter_kernel_ms_per_step"), "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else d.get("e2e_error"))
except Exception as e: print("$X ERR", e)
PY
tail -n 3 gpurun_out/r2_n2d_bench_$X.err | cut -c1-300
done
