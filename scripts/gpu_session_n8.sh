#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tests/multigpu_check.py --reads 400000 --slots 24000000 --front-end > gpurun_out/r2_n8_check.json 2> gpurun_out/r2_n8_check.err; tail -n 1 gpurun_out/r2_n8_check.json | cut -c1-600
timeout 500 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_n8_bench.json 2> gpurun_out/r2_n8_bench.err
timeout 600 $TR --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 2 --workload C5s > gpurun_out/r2_n8_bench_C5s.json 2> gpurun_out/r2_n8_bench_C5s.err
python - <<'PY'
import json
for f in ["r2_n8_bench.json","r2_n8_bench_C5s.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "ms", round(d["ms_per_step"],2), "G/s", round(d["value"]/1e9,2), "insert", round(r["kernel_ms_per_step"],2), "build", round(r["build_kernels_ms_per_step"],2), "layout", round(r["layout_ms_per_step"],2), "S", d.get("sub_blocks_used"), "nvlink", d.get("nvlink",{}).get("gbs_out") or d.get("nvlink",{}).get("gbs_in"), "scat", d.get("nvlink",{}).get("scatter_kernel_ms_per_step"), "fb", d.get("optimistic_exchange_fallbacks"), "e2e", d["e2e"]["ms_per_step"] if d.get("e2e") else d.get("e2e_error"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 4 gpurun_out/r2_n8_*.err | cut -c1-300
