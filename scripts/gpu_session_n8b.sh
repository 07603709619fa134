#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tests/multigpu_check.py --reads 400000 --slots 24000000 --exchange pull > gpurun_out/r2_n8b_check.json 2> gpurun_out/r2_n8b_check.err; tail -n 1 gpurun_out/r2_n8b_check.json | cut -c1-500
for X in pull peer; do
S=4; if [ $X = peer ]; then S=1; fi
timeout 500 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --exchange $X --sub-blocks $S --no-micro > gpurun_out/r2_n8b_bench_$X.json 2> gpurun_out/r2_n8b_bench_$X.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_n8b_bench_$X.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$X ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "fb", d.get("optimistic_exchange_fallbacks"), "nvlink", d.get("nvlink"), "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else d.get("e2e_error"))
except Exception as e: print("$X ERR", e)
PY
tail -n 3 gpurun_out/r2_n8b_bench_$X.err | cut -c1-300
done
