"""Sweep of the pipelined export's knobs on the C2 table (one build, many exports): chunk size, pinned ring slots, host threads.
Prints ms per export and the compact/plain mix.  Usage: python scripts/export_sweep.py [reps]"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dbg_assembly_b200 as dbg
from dbg_assembly_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = synth.CONFIGS["C2"]
n, L, K = cfg["n_reads"], cfg["read_len"], cfg["K"]
p = synth.make_params(cfg["seed"], cfg["genome_len"], L, cfg["insert"], cfg["err"], cfg["n_rate"])
dev = torch.device("cuda:0")
d_bases = torch.empty(n * L, dtype=torch.uint8, device=dev)
synth.reads_device(p, 0, n, d_bases.data_ptr(), device=0)
d_offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
torch.cuda.synchronize()
b = dbg.DBGBuilder(K=K, max_read_len=cfg["max_read_len"], init_slots=int(cfg["init_g"] * 1e9), device=0, track_order=True)
b.submit_device(d_bases.data_ptr(), d_offs.data_ptr(), n, 0, n * L, first_read_index=0)
st = b.finalize()
P = st["array_size"]
h_arr = dbg.capi.PinnedBuffer(P * 16); h_nul = dbg.capi.PinnedBuffer(P // 8 + 1)
L_ = dbg.capi.load()

def run(env):
    for k in ("DBG_B200_EXPORT", "DBG_B200_EXPORT_CHUNK", "DBG_B200_EXPORT_SLOTS", "DBG_B200_EXPORT_THREADS", "DBG_B200_EXPORT_NO_DIRECT", "DBG_B200_EXPORT_PLAIN_PCT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        dbg.capi.check(L_.dbg_export_kmerset(b.h, h_arr.ptr, h_nul.ptr), "export")
        ts.append((time.perf_counter() - t0) * 1e3)
    info = b.export_info()
    print(json.dumps({"env": env, "ms": [round(t, 1) for t in ts[1:]], "first_ms": round(ts[0], 1), "compact": info["chunks_compact"], "plain": info["chunks_plain"],
                      "link_GB": round(info["link_bytes"] / 1e9, 2)}), flush=True)

run({"DBG_B200_EXPORT": "plain"})
sweeps = []
for chunk in (1 << 16, 1 << 18, 1 << 20):
    for slots in (2, 4, 8, 19):
        for thr in (4, 8, 15):
            sweeps.append({"DBG_B200_EXPORT_CHUNK": str(chunk), "DBG_B200_EXPORT_SLOTS": str(slots), "DBG_B200_EXPORT_THREADS": str(thr)})
for pct in (50, 70, 85):
    for thr in (8, 15):
        sweeps.append({"DBG_B200_EXPORT_PLAIN_PCT": str(pct), "DBG_B200_EXPORT_THREADS": str(thr)})
for e in sweeps:
    run(e)
run({"DBG_B200_EXPORT": "plain"})
