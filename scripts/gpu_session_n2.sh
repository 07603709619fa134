#!/bin/bash
# 2-GPU session: parity of the default exchange + layout + merged image on real GPUs, the drop-in front end on 2 GPUs,
# K=63, the exact exchange, then the bench lines (C2 and C3 at N=2)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu_check.py --front-end > gpurun_out/r2_n2_check_k31.json 2> gpurun_out/r2_n2_check_k31.err; tail -1 gpurun_out/r2_n2_check_k31.json | cut -c1-700
timeout 400 $TR --master-port 29512 tests/multigpu_check.py --K 63 > gpurun_out/r2_n2_check_k63.json 2> gpurun_out/r2_n2_check_k63.err; tail -1 gpurun_out/r2_n2_check_k63.json | cut -c1-400
timeout 400 $TR --master-port 29513 tests/multigpu_check.py --exchange peer_exact > gpurun_out/r2_n2_check_exact.json 2> gpurun_out/r2_n2_check_exact.err; tail -1 gpurun_out/r2_n2_check_exact.json | cut -c1-400
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; tail -1 gpurun_out/r2_n2_bench.json | cut -c1-1500
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 2 --workload C3 > gpurun_out/r2_n2_bench_C3.json 2> gpurun_out/r2_n2_bench_C3.err; tail -1 gpurun_out/r2_n2_bench_C3.json | cut -c1-900
timeout 400 $TR --master-port 29516 bench.py --gpus 2 --steps 10 --warmup 3 --exchange peer_exact > gpurun_out/r2_n2_bench_exact.json 2> gpurun_out/r2_n2_bench_exact.err; tail -1 gpurun_out/r2_n2_bench_exact.json | cut -c1-600
tail -3 gpurun_out/r2_n2_*.err | cut -c1-300
