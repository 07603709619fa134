#!/usr/bin/env python
"""bench.py -- canonical k-mers/s into the De Bruijn graph (count + edges) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]

A "step" is one complete build of the graph from the synthetic read set of the workload:
table clear -> fused encode/extract/insert kernels over ALL reads -> reference-layout image + k-mer-0 node.
`value` has the reads resident in HBM; `e2e` goes through the host-buffer C ABI (dbg_submit_reads ...
dbg_export_kmerset) with the H2D copy of the reads and the D2H copy of the finished KmerSet image inside
the timed region.  N>1: one process per GPU (torchrun), reads dealt in contiguous blocks, k-mers sharded by
owner slot range with one NCCL all-to-all per step; weak scaling (genome, reads and table grow with N).

--impl reference times the reference's own CPU build (oracle/_ref/ref_build_driver = the unmodified
DBG_contig sources) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "canonical k-mers/sec into DBG (count+edges)"
UNIT = "k-mers/s"


def workload(name, n_gpus, scale=1.0):
    from dbg_assembly_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    cfg["n_reads"] = int(cfg["n_reads"] * scale) // 2 * 2
    cfg["genome_len"] = max(int(cfg["genome_len"] * scale), 4 * cfg["insert"])
    cfg["init_g"] = cfg["init_g"] * scale
    # weak scaling: per-GPU reads fixed, genome and table grow with N
    cfg["genome_len_total"] = cfg["genome_len"] * n_gpus
    cfg["n_reads_total"] = cfg["n_reads"] * n_gpus
    cfg["init_g_total"] = cfg["init_g"] * n_gpus
    return cfg


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.index, self.stop_flag, self.th = index, False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True); self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own build on the host cores, on the FULL named workload
# ---------------------------------------------------------------------------------------------------
SYNTH_FASTA = os.path.join(REPO, "oracle", "_bin", "synth_fasta")


def write_workload_fasta(cfg, n_reads, path, raw=False):
    """the workload's reads as one-line FASTA, written by the stand-alone generator (oracle/tools/synth_fasta.cpp: same
    counter-based generator as the library, but nothing of the product is loaded by this process)"""
    if not os.access(SYNTH_FASTA, os.X_OK):
        subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "tools"], check=True, stdout=subprocess.DEVNULL)
    err = int(round(cfg["err"] * (1 << 24))); nr = int(round(cfg["n_rate"] * (1 << 24)))
    cmd = [SYNTH_FASTA, str(cfg["seed"]), str(cfg["genome_len"]), str(cfg["read_len"]), str(cfg["insert"]), str(err), str(nr), "0",
           str(n_reads), path] + (["raw"] if raw else [])
    subprocess.run(cmd, check=True)


def scratch_dir(need_bytes=0):
    """tmpfs when it has room (the reference reads the file through the page cache either way)"""
    for d in ("/dev/shm", None):
        try:
            if d is None or (os.path.isdir(d) and os.statvfs(d).f_bavail * os.statvfs(d).f_frsize > need_bytes + (2 << 30)):
                return tempfile.TemporaryDirectory(dir=d)
        except Exception:
            pass
    return tempfile.TemporaryDirectory()


def cpu_reference_run(cfg, n_reads, threads, steps, warmup, name, budget_s=None):
    """The reference build (oracle/_ref/ref_build_driver: the unmodified DBG_contig sources around build_debruijn_graph)
    on the first n_reads reads of the workload (n_reads == cfg["n_reads"]: the whole workload, its own -i).
    Returns (values [k-mers/s per timed step], info).  budget_s bounds the wall clock: at least one warm-up (if asked
    for) and two timed builds always run, further ones only while the budget lasts."""
    from oracle import oracle as orc          # test infrastructure: only this leg may touch it
    occ = n_reads * (cfg["read_len"] - cfg["K"] + 1)
    frac = n_reads / cfg["n_reads"]
    init_g = cfg["init_g"] * frac             # == cfg["init_g"] for the whole workload
    vals = []
    t_start = time.perf_counter()
    with scratch_dir(n_reads * (cfg["read_len"] + 14)) as td:
        if orc.have_reference():
            kind = "reference"
            path = os.path.join(td, "reads.fa")
            write_workload_fasta(cfg, n_reads, path)          # once, outside the timed builds
            done_w = 0
            while True:
                elapsed = time.perf_counter() - t_start
                over = budget_s is not None and elapsed > budget_s
                # warm-ups: all that were asked for, but beyond the first only while they fit a quarter of the budget
                if done_w < warmup and (done_w == 0 or budget_s is None or elapsed < budget_s / 4):
                    orc.run_ref_build([path], cfg["K"], cfg["max_read_len"], init_g, threads=threads, dump=False, timeout=3000)
                    done_w += 1
                    continue
                if len(vals) >= steps or (over and len(vals) >= min(2, steps)):
                    break
                stats, _ = orc.run_ref_build([path], cfg["K"], cfg["max_read_len"], init_g, threads=threads, dump=False, timeout=3000)
                vals.append(occ / stats["wall_s"])
        else:
            kind, threads = "port", 1
            path = os.path.join(td, "reads.raw")
            write_workload_fasta(cfg, n_reads, path, raw=True)
            bases = np.fromfile(path, dtype=np.uint8)
            offs = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(cfg["read_len"])
            for i in range(min(warmup, 1) + min(steps, 2)):
                t0 = time.perf_counter()
                o = orc.OracleGraph(cfg["K"], cfg["max_read_len"], int(init_g * 1e9), 0.7, 10, 10000)
                o.add_file(bases, offs); o.finish()
                dt = time.perf_counter() - t0
                o.close()
                if i >= min(warmup, 1):
                    vals.append(occ / dt)
    whole = n_reads == cfg["n_reads"]
    info = {"kind": kind, "cores": threads, "steps_run": len(vals),
            "sample": (f"the WHOLE workload {name}: {n_reads} reads, {occ} k-mer occurrences, -i {init_g:g}" if whole else
                       f"first {n_reads} reads of {name} ({occ} k-mer occurrences, {frac:.3f} of the workload), table -i {init_g:.4g} (scaled with the sample)")
                      + ", one-line FASTA written once before the timed builds, whole build_debruijn_graph() incl. file read, wall clock"}
    return vals, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # the same configuration as the B200 arm at this N: weak scaling, so at --gpus N the reference builds the N-GPU
    # workload (N x genome, N x reads, N x table) -- on the host cores, it has no other way to use the box
    n_g = max(1, args.gpus)
    cfg_n = workload(args.workload, n_g, args.scale)
    cfg = dict(cfg_n, genome_len=cfg_n["genome_len_total"], n_reads=cfg_n["n_reads_total"], init_g=cfg_n["init_g_total"])
    threads = os.cpu_count() or 1
    n_reads = cfg["n_reads"] if args.ref_sample_reads <= 0 else min(cfg["n_reads"], args.ref_sample_reads)
    vals, info = cpu_reference_run(cfg, n_reads, threads, args.steps, args.warmup, args.workload, budget_s=args.ref_budget_s)
    v = float(np.mean(vals))
    occ = n_reads * (cfg["read_len"] - cfg["K"] + 1)
    cd = config_dict(args, cfg_n, n_g)
    if n_reads != cfg["n_reads"]:
        cd["reference_sample_reads"] = n_reads
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * occ / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": cd,
            "cpu_baseline": dict(info, value=v, unit=UNIT),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "steps_run": info["steps_run"], "per_step_values": vals}
    print(json.dumps(line))
    return 0


def config_dict(args, cfg, n):
    return {"workload": f"{args.workload}: synthetic {cfg['genome_len'] / 1e6:.2f} Mb genome x{n} GPUs, "
                        f"{cfg['n_reads']} x {cfg['read_len']} bp paired reads per GPU ({cfg['n_reads'] * cfg['read_len'] / cfg['genome_len']:.0f}x), "
                        f"{cfg['err'] * 100:g}% substitutions, {cfg['n_rate'] * 100:g}% N, K={cfg['K']}, -r {cfg['max_read_len']}, -i {cfg['init_g'] * n:g}",
            "K": cfg["K"], "reads_per_gpu": cfg["n_reads"], "read_len": cfg["read_len"], "table_slots_request": int(cfg["init_g"] * n * 1e9),
            "parallelism": f"hash-sharded x{n}" if n > 1 else "single GPU",
            "l2_policy": "inputs (reads 460 MB, table 6.4 GB per GPU) far exceed the 126 MB L2; table re-zeroed every step",
            "layout_parity": "track_order=1 (reference -t 1 slot layout)" if n == 1 else
                             "track_order=1; every rank lays out its slice of the reference table after the cross-shard hand-off of boundary clusters (same work per GPU as N=1)"}


def short_workload(args, wname, local, dev, torch, dbg, synth, steps=3, warmup=2):
    """one of the other named workloads on one GPU, device-resident reads: ms per build, k-mers/s, per-phase times, and the
    end-to-end figure through host buffers (1 step)"""
    from dbg_assembly_b200.graph import torch_stream_handle
    cfg = workload(wname, 1, 1.0)
    n, L, K = cfg["n_reads"], cfg["read_len"], cfg["K"]
    p = synth.make_params(cfg["seed"], cfg["genome_len"], L, cfg["insert"], cfg["err"], cfg["n_rate"])
    d_bases = torch.empty(n * L, dtype=torch.uint8, device=dev)
    synth.reads_device(p, 0, n, d_bases.data_ptr(), device=local)
    d_offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    out = {"config": config_dict(argparse.Namespace(workload=wname), cfg, 1)["workload"]}
    with dbg.DBGBuilder(K=K, max_read_len=cfg["max_read_len"], init_slots=int(cfg["init_g"] * 1000000000), device=local, track_order=True) as b:
        b.set_stream(torch_stream_handle(dev))
        for _ in range(warmup):
            b.reset(); b.submit_device(d_bases.data_ptr(), d_offs.data_ptr(), n, 0, n * L, first_read_index=0); st = b.finalize()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            b.reset(); b.submit_device(d_bases.data_ptr(), d_offs.data_ptr(), n, 0, n * L, first_read_index=0); st = b.finalize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tm = b.timings()
        out.update({"ms_per_step": ms, "value": st["occurrences"] / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": warmup,
                    "occurrences_per_step": int(st["occurrences"]), "nodes": int(st["count"]), "dtype": "u128" if K > 31 else "u64",
                    "clear_ms": tm["clear_ms"], "build_kernels_ms": tm["build_ms"], "insert_ms": tm["insert_ms"], "layout_ms": tm["layout_ms"],
                    "paths": b.path_counts()})
    del d_bases, d_offs
    torch.cuda.empty_cache()
    return out


def e2e_sharded(args, cfg, sb, d_bases, d_offs, n, L, K, first_read, occ_rank, st, rank, world, dev, dist, torch):
    import dbg_assembly_b200 as dbg
    from dbg_assembly_b200.sharded import SharedImage
    Lb = dbg.capi.load()
    P = st["array_size"]
    wide = K > 31
    nb = 32 if wide else 16
    path = [f"/dev/shm/dbg_b200_e2e_{os.getpid()}.img"] if rank == 0 else [None]
    dist.broadcast_object_list(path, src=0)
    path = path[0]
    ok = torch.tensor([1 if (rank != 0 or SharedImage.room_for(P, wide)) else 0], device=dev)
    dist.broadcast(ok, src=0)
    if int(ok.item()) == 0:
        raise RuntimeError(f"/dev/shm cannot hold the {P * nb / 1e9:.1f} GB shared table image")
    img = SharedImage(path, P, wide, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        img = SharedImage(path, P, wide, create=False)
    # page-lock only what this rank writes: its slice of the image (it may wrap past slot P-1)
    g_first, n_slots, _ = sb.b.shard_slice_info()
    regs = []

    def reg(off, ln):
        if ln <= 0:
            return
        a0 = off // 4096 * 4096
        a1 = min(img.total, (off + ln + 4095) // 4096 * 4096)
        dbg.capi.check(Lb.dbg_host_register(img.base_ptr + a0, a1 - a0), "dbg_host_register")
        regs.append(img.base_ptr + a0)
    first_len = min(n_slots, P - g_first)
    reg(g_first * nb, first_len * nb)
    if first_len < n_slots:
        reg(0, (n_slots - first_len) * nb)
    h_bases = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    h_offs = (torch.arange(n + 1, dtype=torch.int64) * L).pin_memory()
    h_bases.copy_(d_bases.cpu())
    d_b2, d_o2 = torch.empty_like(d_bases), torch.empty_like(d_offs)

    def e2e_step():
        d_b2.copy_(h_bases, non_blocking=True)
        d_o2.copy_(h_offs, non_blocking=True)
        sb.b.reset()
        sb.add_reads_device(d_b2, d_o2, n, 0, n * L, first_read, occ_rank)
        s = sb.finalize(layout=True)
        sb.export_into(img, s)
        return s
    e2e_step()
    torch.cuda.synchronize(); dist.barrier()
    e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        s2 = e2e_step()
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / e_steps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    # cheap integrity figure of the merged image (rank 0): filled slots in the first 8 Mi slots vs the bitmap's count
    chk = None
    if rank == 0:
        m = min(P, 8 << 20)
        chk = {"slots_checked": int(m), "nonzero_kmers": int(np.count_nonzero(img.arr["kmer"][:m])),
               "nul_bits": int(np.unpackbits(img.nul[: (m + 7) // 8])[:m].sum())}
    out = {"value": s2["global_occurrences"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(n * L + (n + 1) * 8) * world,
           "d2h_bytes_per_step": int(img.img_bytes + img.nul_bytes), "ms_per_step": dt * 1e3, "steps": e_steps,
           "what": "per rank: pinned host reads -> H2D -> fused exchange -> insert -> cross-shard hand-off -> layout -> D2H of the rank's slice into ONE "
                   "shared host table image (/dev/shm); rank 0 fixes the shared nul_flag bytes and adds the k-mer-0 node; wall clock, max over ranks",
           "merged_image_slots": int(P), "merged_image_check": chk}
    for r in regs:
        Lb.dbg_host_unregister(r)
    dist.barrier()
    img.close(unlink=rank == 0)
    return out


# ---------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------
def kfreq_measure(cfg, world, rank, local, dev, steps, warmup, dist, torch, synth):
    """C4: the K-mer frequency table of correct_error (kfreq_* C ABI).  A step = zero this rank's part of the 4^K-entry table,
    all-gather the ranks' reads (N > 1; 1.25 B per occurrence over NVLink), extract every canonical k-mer and count the ones
    in the rank's index range (table sharded by .cz block), read the counters back.  Device-timed, max over ranks."""
    from dbg_assembly_b200.kfreq import KmerFreq
    n, L, K = cfg["n_reads"], cfg["read_len"], cfg["K"]
    p = synth.make_params(cfg["seed"], cfg["genome_len_total"], L, cfg["insert"], cfg["err"], cfg["n_rate"])
    mine = torch.empty(n * L, dtype=torch.uint8, device=dev)
    synth.reads_device(p, rank * n, n, mine.data_ptr(), device=local)
    everything = torch.empty(world * n * L, dtype=torch.uint8, device=dev) if world > 1 else mine
    d_offs = torch.arange(world * n + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    kf = KmerFreq(K=K, device=local, block_rank=rank, block_count=world)

    def step():
        kf.reset()
        if world > 1:
            dist.all_gather_into_tensor(everything, mine)
            torch.cuda.synchronize()
        kf.submit_device(everything.data_ptr(), d_offs.data_ptr(), world * n, 0, world * n * L)
        return kf.finalize()
    for _ in range(warmup):
        st = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st = step()
    e1.record(); torch.cuda.synchronize()
    # (the library works on its own stream and synchronises inside kfreq_finalize: the events bracket whole steps)
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    hist = kf.histogram()
    species = int(hist[1:].sum())
    lo, hi = kf.index_range()
    if world > 1:
        t = torch.tensor([species], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        species = int(t.item())
    occ = int(st["occurrences"])          # every rank extracts every read: the global count
    kf.close()
    del mine, everything, d_offs
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "value": occ / (ms * 1e-3), "unit": UNIT, "occurrences_per_step": occ, "species": species,
            "table_entries_per_gpu": int(hi - lo), "table_bytes_per_gpu": int(hi - lo) * 4, "steps": steps, "warmup": warmup, "dtype": "u32",
            "config": f"C4: K={K} direct-index frequency table (4^{K} u32 counters, sharded by .cz block over {world} GPUs), "
                      f"{cfg['genome_len_total'] / 1e6:.1f} Mb genome, {world * n} x {L} bp reads ({world * n * L / cfg['genome_len_total']:.0f}x), "
                      f"{cfg['err'] * 100:g}% substitutions",
            "what": "zero the table + (N > 1: all-gather of the reads) + k_build<FreqSink> (fused 2-bit pack, canonical k-mer, RED.ADD into the "
                    "rank's index range) + counters; every rank extracts all reads and keeps its range (DESIGN.md section 6)"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import dbg_assembly_b200 as dbg
    from dbg_assembly_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py --gpus {args.gpus} must be launched with torchrun (one rank per GPU)", file=sys.stderr)
            return 2
    if dbg.capi.device_count() == 0:
        print("bench.py: no CUDA device -- libdbgb200 has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # keep stdout to the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1 behind python's back, so
    # fd 1 is pointed at stderr for the whole run and the result goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload(args.workload, world, args.scale)
    if args.workload == "C4":
        # the other front end of the north star: correct_error's K-mer frequency table, sharded by index range
        sampler = ClockSampler(local)
        sampler.start()
        r = kfreq_measure(cfg, world, rank, local, dev, args.steps, args.warmup, dist, torch, synth)
        clocks = sampler.stop()
        peak, peak_src = measured_peaks()
        ach = r["occurrences_per_step"] * 8 / (r["ms_per_step"] * 1e-3) / 1e9
        line = {"metric": "canonical k-mers/sec into the K-mer frequency table (count)", "value": r["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": {"workload": r["config"], "K": cfg["K"],
                "l2_policy": "the 68.7 GB table (whole, or its per-GPU shard) and the reads far exceed the 126 MB L2; table re-zeroed every step"},
                "clocks": clocks, "gpu_launches": 3 * args.steps, "occurrences_per_step": r["occurrences_per_step"], "species": r["species"],
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "kernel": "k_build<FreqSink> (random 4-B RED.ADD per occurrence into the direct-index table)",
                             "algorithmic_bytes_per_occurrence": 8, "peak_source": peak_src},
                "e2e": None, "what": r["what"]}
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        sys.stdout.flush()
        if rank == 0:
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        os.close(real_stdout)
        return 0
    if args.init_g:
        cfg["init_g"] = args.init_g; cfg["init_g_total"] = args.init_g * world
    n, L, K = cfg["n_reads"], cfg["read_len"], cfg["K"]
    occ_rank = n * (L - K + 1)
    p = synth.make_params(cfg["seed"], cfg["genome_len_total"], L, cfg["insert"], cfg["err"], cfg["n_rate"])
    d_bases = torch.empty(n * L, dtype=torch.uint8, device=dev)
    first_read = rank * n
    synth.reads_device(p, first_read, n, d_bases.data_ptr(), device=local)
    d_offs = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    from dbg_assembly_b200.graph import torch_stream_handle
    stream = torch_stream_handle(dev)
    init_slots = int(cfg["init_g_total"] * 1000000000)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    extra = {}
    if world == 1:
        b = dbg.DBGBuilder(K=K, max_read_len=cfg["max_read_len"], init_slots=init_slots, device=local, track_order=True)
        b.set_stream(stream)

        def step():
            b.reset()
            b.submit_device(d_bases.data_ptr(), d_offs.data_ptr(), n, 0, n * L, first_read_index=0)
            return b.finalize()
        closer = b
        launch_count = lambda: b.launches  # noqa: E731
        timings = b.timings
    else:
        from dbg_assembly_b200.sharded import ShardedBuilder
        sb = ShardedBuilder(K=K, max_read_len=cfg["max_read_len"], init_slots=init_slots, device=local, track_order=True,
                            exchange=args.exchange, sub_blocks=args.sub_blocks)
        extra["exchange"] = sb.exchange
        extra["sub_blocks"] = sb.sub_blocks
        sb.b.set_stream(stream)

        def step():
            sb.b.reset()
            sb.add_reads_device(d_bases, d_offs, n, 0, n * L, first_read, occ_rank)
            return sb.finalize(layout=True)
        closer = sb
        launch_count = lambda: sb.b.launches  # noqa: E731
        timings = sb.b.timings

    for _ in range(args.warmup):
        st = step()
    barrier()
    l0 = launch_count()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    build_ms, clear_ms, layout_ms, insert_ms, scatter_ms = [], [], [], [], []
    for _ in range(args.steps):
        st = step()
        tm = timings()
        build_ms.append(tm["build_ms"]); clear_ms.append(tm["clear_ms"]); layout_ms.append(tm["layout_ms"]); insert_ms.append(tm["insert_ms"])
        scatter_ms.append(tm.get("scatter_ms", 0.0))
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        occ_total = st["global_occurrences"]
        nodes_total = st["global_count"]
        extra["exchange_bytes_per_step_rank0"] = sb.exchange_bytes // (args.steps + args.warmup)
        extra["optimistic_exchange_fallbacks"] = sb.opt_fallbacks
        extra["sub_blocks_used"] = sb.sub_blocks_used
        sc_ms = float(np.mean(scatter_ms))
        if sc_ms > 0:
            # NVLink figure: bytes this rank moved to/from its peers over the kernel that moves them (rank 0, CUDA events) --
            # pull: the owners' reads inside k_insert_tuples_pull (by symmetry what rank 0 reads = what it serves);
            # peer: the stores of the fused scatter kernel into the owners' receive regions
            xb = extra["exchange_bytes_per_step_rank0"]
            if sb.exchange == "pull" and not sb.opt_fallbacks:
                im = float(np.mean(insert_ms))
                extra["nvlink"] = {"scatter_kernel_ms_per_step": sc_ms, "insert_kernel_ms_per_step": im, "bytes_in_per_step": xb,
                                   "gbs_in": xb / (im * 1e-3) / 1e9 if im > 0 else None,
                                   "what": "pull exchange: k_build<StagedScatterSink<OPT>> partitions by (owner, table slice) into the LOCAL send buffer (no NVLink); "
                                           "k_insert_tuples_pull reads the peers' regions over NVLink peer mappings while it inserts"}
            else:
                extra["nvlink"] = {"scatter_kernel_ms_per_step": sc_ms, "bytes_out_per_step": xb, "gbs_out": xb / (sc_ms * 1e-3) / 1e9,
                                   "what": "k_build<StagedScatterSink<OPT>> by owner: extraction fused with stores into the owners' receive regions over NVLink peer mappings"}
    else:
        occ_total = st["occurrences"]
        nodes_total = st["count"]
    ms_per_step = ms_total / args.steps
    value = occ_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: the bucketed hash-insert kernel (k_insert_tuples) when the
    # ---- partitioned path ran, else the fused direct-insert build kernel.  achieved = occurrences x 32 B
    # ---- (one 16-B slot read + one 16-B slot write) / that kernel's CUDA-event time.
    peak, peak_src = measured_peaks()
    kb = float(np.mean(build_ms))
    ki = float(np.mean(insert_ms))
    dom_ms = ki if ki > 0 else kb
    achieved = st["occurrences"] * 32 / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
    if world == 1:
        kname = ("k_insert_tuples (bucket-ordered hash insert, L2-resident table slices)" if ki > 0
                 else "k_build<InsertSink> (fused 2-bit pack + canonical k-mer + direct hash insert)")
    else:
        kname = "k_insert_tuples (owner-side insert of exchanged tuples)"
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": None, "kernel": kname, "algorithmic_bytes_per_occurrence": 32, "kernel_ms_per_step": dom_ms,
            "build_kernels_ms_per_step": kb, "clear_ms_per_step": float(np.mean(clear_ms)),
            "layout_ms_per_step": float(np.mean(layout_ms)), "peak_source": peak_src}
    traffic_file = os.path.join(REPO, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu capture
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            if world == 1 and args.scale == 1.0 and ki > 0:
                roof["traffic"] = tj.get("k_insert_tuples_dram_bytes_per_launch")
                roof["traffic_source"] = tj.get("source")
        except Exception:
            pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": config_dict(args, cfg, world), "clocks": clocks, "gpu_launches": int(launches),
            "occurrences_per_step": int(occ_total), "nodes": int(nodes_total), "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "roofline": roof}
    line.update(extra)

    # ---- random-access denominator measured on the spot (rank 0) ----
    if rank == 0 and not args.no_micro:
        try:
            import ctypes as C
            Lb = dbg.capi.load()
            res = {}
            for mode, tag in ((0, "load_store"), (1, "load_cas")):
                ms = C.c_float(0)
                n_ops = 1 << 28
                dbg.capi.check(Lb.dbg_measure_random_rmw(local, 6 << 30, n_ops, mode, C.byref(ms)), "dbg_measure_random_rmw")
                res[tag + "_gbs_32B"] = n_ops * 32 / (ms.value * 1e-3) / 1e9      # algorithmic 32 B per op, like `achieved`
            line["random_access"] = dict(res, note="uniform random 32-B sector RMW over a 6 GiB table, 2^28 ops, same units as roofline.achieved")
            if achieved:
                line["roofline"]["frac_of_random_rmw"] = achieved / res["load_cas_gbs_32B"]
        except Exception as e:      # the microbenchmark must never take the headline down
            line["random_access"] = {"error": str(e)}

    # ---- e2e through the host-buffer C ABI (N=1: full export; N>1: host submit is single-GPU only) ----
    if world == 1:
        closer.close()
        P = st["array_size"]
        nbytes_node = 32 if K > 31 else 16
        bufs = []
        try:
            h_bases = dbg.capi.PinnedBuffer(n * L); bufs.append(h_bases)
            h_offs = dbg.capi.PinnedBuffer((n + 1) * 8); bufs.append(h_offs)
            h_bases.array[:] = d_bases.cpu().numpy()
            np.frombuffer(h_offs.array, dtype=np.uint64)[:] = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
            h_arr = dbg.capi.PinnedBuffer(P * nbytes_node); bufs.append(h_arr)
            h_nul = dbg.capi.PinnedBuffer(P // 8 + 1); bufs.append(h_nul)
            del d_bases
            torch.cuda.empty_cache()
            Lb = dbg.capi.load()

            def measure(api):
                b2 = dbg.DBGBuilder(K=K, max_read_len=cfg["max_read_len"], init_slots=init_slots, device=local, track_order=True)
                try:
                    def e2e_step():
                        b2.reset()
                        if api == "finish_export":
                            return b2.finish_export_ptr(h_bases.ptr, h_offs.ptr, n, h_arr.ptr, h_nul.ptr)
                        b2.submit_ptr(h_bases.ptr, h_offs.ptr, n)
                        s = b2.finalize()
                        dbg.capi.check(Lb.dbg_export_kmerset(b2.h, h_arr.ptr, h_nul.ptr), "dbg_export_kmerset")
                        return s
                    for _ in range(max(1, min(args.warmup, 2))):
                        e2e_step()
                    torch.cuda.synchronize()
                    e_steps = max(1, min(args.steps, 5))
                    t0 = time.perf_counter()
                    for _ in range(e_steps):
                        s2 = e2e_step()
                    torch.cuda.synchronize()
                    dt = (time.perf_counter() - t0) / e_steps
                    tm = b2.timings()
                    xi = b2.export_info()
                finally:
                    b2.close()
                # the image on the host is a table: over its first 8 Mi slots every set nul_flag bit has a node and vice versa
                m = min(P, 1 << 23) // 8 * 8
                img = h_arr.array[: m * nbytes_node].view(np.uint64).reshape(m, nbytes_node // 8)
                occupied = (img != 0).any(axis=1)
                bits = np.unpackbits(h_nul.array[: m // 8]).astype(bool)
                check = {"slots_checked": int(m), "nodes": int(occupied.sum()), "nul_bits": int(bits.sum()),
                         "consistent": bool((occupied & ~bits).sum() == 0 and (bits & ~occupied).sum() <= 1)}      # (the k-mer-0 node is all zero)
                return {"value": s2["occurrences"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(n * L + (n + 1) * 8),
                        "d2h_bytes_per_step": int(xi["link_bytes"]), "ms_per_step": dt * 1e3, "steps": e_steps,
                        "h2d_ms": tm["h2d_ms"], "d2h_ms": tm["d2h_ms"], "build_ms": tm["build_ms"], "layout_ms": tm["layout_ms"],
                        "result_bytes_on_host": int(P * nbytes_node + P // 8 + 1), "export": xi, "api": api, "host_image_check": check,
                        "nodes": int(s2["count"]),
                        "what": ("dbg_reset + dbg_finish_export(pinned host reads -> the whole P-slot KmerSet image + nul_flag in pinned host memory): one "
                                 "call = dbg_submit_reads + dbg_finalize + dbg_export_kmerset, pipelined (H2D || extraction; insert / layout / D2H slice "
                                 "group by slice group: export.chunks_plain windows); wall clock" if api == "finish_export" else
                                 "dbg_reset + dbg_submit_reads (pinned host reads) -> dbg_finalize -> dbg_export_kmerset (KmerSet image into pinned host memory); wall clock")}
            try:
                line["e2e"] = measure(args.e2e_api)
            except Exception as e1:
                if args.e2e_api != "finish_export":
                    raise
                # the fused call failed on this box: the three separate calls are the same contract
                line["e2e_finish_export_error"] = f"{type(e1).__name__}: {e1}"
                line["e2e"] = measure("separate")
        except Exception as e:
            line["e2e"] = None
            line["e2e_error"] = f"{type(e).__name__}: {e}"
        for hb in bufs:
            hb.close()
    else:
        # ---- e2e at N>1: every rank's reads start in ITS pinned host memory; H2D, exchange, insert, cross-shard hand-off,
        # ---- layout; every rank copies its slice into ONE shared host table image (a /dev/shm mapping = the KmerSet a
        # ---- single consumer process would traverse); rank 0 fixes the shared nul_flag bytes and adds the k-mer-0 node.
        try:
            line["e2e"] = e2e_sharded(args, cfg, sb, d_bases, d_offs, n, L, K, first_read, occ_rank, st, rank, world, dev, dist, torch)
        except Exception as e:
            line["e2e"] = None
            line["e2e_error"] = f"{type(e).__name__}: {e}"
        closer.close()

    # ---- the other named single-GPU workloads, short (device-resident, 3 steps): C1 bundled-test shape, C3 K=63 ----
    if world == 1 and rank == 0 and not args.no_other and args.scale == 1.0 and args.workload == "C2":
        line["other_workloads"] = {}
        for wname in ("C1", "C3"):
            try:
                line["other_workloads"][wname] = short_workload(args, wname, local, dev, torch, dbg, synth)
            except Exception as e:
                line["other_workloads"][wname] = {"error": f"{type(e).__name__}: {e}"}
        try:      # C4's per-GPU share of reads against the WHOLE 4^17 table on this one GPU (68.7 GB)
            line["other_workloads"]["C4"] = kfreq_measure(workload("C4", 1, 1.0), 1, 0, local, dev, 2, 1, dist, torch, synth)
        except Exception as e:
            line["other_workloads"]["C4"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference on the WHOLE workload, all host threads, one build;
    # ---- and with -t 1 ----
    if world == 1 and rank == 0 and not args.no_cpu:
        try:
            threads = os.cpu_count() or 1
            cfg1 = workload(args.workload, 1, args.scale)
            nr = cfg1["n_reads"] if args.ref_sample_reads <= 0 else min(cfg1["n_reads"], args.ref_sample_reads)
            vals, info = cpu_reference_run(cfg1, nr, threads, 1, 0, args.workload)
            line["cpu_baseline"] = dict(info, value=float(np.mean(vals)), unit=UNIT)
            if info["kind"] == "reference" and threads > 1 and not args.no_cpu_t1:
                try:
                    vals1, _ = cpu_reference_run(cfg1, nr, 1, 1, 0, args.workload)
                    line["cpu_baseline"]["value_t1"] = float(np.mean(vals1))
                except Exception as e1:
                    line["cpu_baseline"]["value_t1_error"] = str(e1)
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    os.close(real_stdout)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only; numbers at scale != 1 are not the metric)")
    ap.add_argument("--init-g", type=float, default=None, help="override the table size -i (experiments only)")
    ap.add_argument("--ref-sample-reads", type=int, default=0,
                    help="reference arm / cpu_baseline: 0 = the whole workload (default), N = only its first N reads (said so in the output)")
    ap.add_argument("--ref-budget-s", type=float, default=900.0,
                    help="reference arm: wall-clock budget; at least 1 warm-up and 2 timed full builds run, more while it lasts")
    ap.add_argument("--exchange", default="pull", choices=["pull", "peer", "peer_exact", "peer_sliced", "nccl"],
                    help="multi-GPU: 'pull' (default) = ONE extraction pass partitions by (owner, table slice) into the source's own send buffer, "
                         "the owners read their regions over NVLink inside the bucketed insert (falls back to 'peer' beyond 4096 buckets); "
                         "'peer' = ONE extraction pass stores tuples into fixed regions of the owners' buffers over NVLink, then an owner-side "
                         "partition + insert; 'peer_exact' = count pass + exact offsets + scatter; 'nccl' = pack + send/recv")
    ap.add_argument("--sub-blocks", type=int, default=4, help="multi-GPU 'peer': sub-blocks per step (scatter k+1 overlaps insert k)")
    ap.add_argument("--e2e-api", default="finish_export", choices=["finish_export", "separate"],
                    help="N=1 e2e leg: the fused, pipelined dbg_finish_export (default) or the three separate calls")
    ap.add_argument("--no-other", action="store_true", help="N=1: skip the short C1 / C3 lines (other_workloads)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cpu-t1", action="store_true", help="skip the -t 1 run of the reference in cpu_baseline")
    ap.add_argument("--no-micro", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and args.scale == 1.0:
        print("note: fewer than 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
