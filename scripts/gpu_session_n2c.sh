#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for S in 1 4; do
timeout 400 $TR --master-port 2952$S bench.py --gpus 2 --steps 10 --warmup 3 --sub-blocks $S --no-micro > gpurun_out/r2_n2c_bench_S$S.json 2> gpurun_out/r2_n2c_bench_S$S.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_n2c_bench_S$S.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("S=$S ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "clear", round(r["clear_ms_per_step"],2), "Sused", d.get("sub_blocks_used"), "scat", d.get("nvlink",{}).get("scatter_kernel_ms_per_step"), "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else d.get("e2e_error"))
PY
done
