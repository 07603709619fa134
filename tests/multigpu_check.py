"""Multi-GPU parity check on real GPUs (not collected by pytest: needs N GPUs and torchrun; the same flow runs on ONE
device inside pytest, tests/test_gpu_sharded_layout.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [--reads 200000] [--K 31] [--exchange pull|peer|peer_exact|peer_sliced|nccl] [--front-end]

Every rank extracts + exchanges + inserts its block of C2-shaped synthetic reads (ShardedBuilder, two calls so that the
exchange repeats), the boundary clusters are handed around the ring, every rank lays out its slice and exports it into
ONE shared host table image; rank 0 compares that image with the oracle's table (== the reference with -t 1): filled
slots, k-mers, link words, slot layout, k-mer-0 node -- bit for bit.  --front-end additionally runs the relinked reference
front end (oracle/_ref/debruijn_contig_b200) with DBG_B200_GPUS=N on the same reads and compares its output files with
the single-GPU run's and, where the reference binary is present, the reference's.  Prints one JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

SUF = (".contig.seq.fa", ".contig.small.fa", ".contig.kmer.freq", ".contig.tip.fa", ".contig.bubble.fa", ".contig.lowedge.fa",
       ".contig.seq.depth", ".contig.small.depth")


def front_end_check(world, hb, ho, K, L):
    """the drop-in binary on N GPUs from ONE process (dbg_mg_*) vs on one GPU vs the reference program"""
    from oracle import oracle as orc
    exe = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_b200")
    ref = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_ref")
    if not os.access(exe, os.X_OK):
        return {"skipped": "oracle/_ref/debruijn_contig_b200 not present"}
    out = {}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "reads.fa"); orc.write_fasta(fa, hb, ho)
        lib = os.path.join(td, "reads.lib"); open(lib, "w").write(fa + "\n")
        files = {}
        for tag, binary, env in (("gpus1", exe, {}), (f"gpus{world}", exe, {"DBG_B200_GPUS": str(world)}), ("reference", ref, {})):
            if not os.access(binary, os.X_OK):
                continue
            pre = os.path.join(td, tag)
            r = subprocess.run([binary, "-k", str(K), "-r", str(L), "-f", "2", "-t", "1", "-i", "0.012", "-M", "100", "-o", pre, lib],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900, env=dict(os.environ, **env))
            if r.returncode != 0:
                return {"error": f"{tag}: rc {r.returncode}: {r.stderr.decode()[-400:]}"}
            files[tag] = {s: open(pre + s, "rb").read() for s in SUF}
        out["multi_equals_single"] = files[f"gpus{world}"] == files["gpus1"]
        if "reference" in files:
            out["multi_equals_reference"] = files[f"gpus{world}"] == files["reference"]
        out["contig_bytes"] = len(files["gpus1"][".contig.seq.fa"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200_000)
    ap.add_argument("--K", type=int, default=31)
    ap.add_argument("--slots", type=int, default=12_000_000)
    ap.add_argument("--exchange", default="peer", choices=["pull", "peer", "peer_exact", "peer_sliced", "nccl"])
    ap.add_argument("--front-end", action="store_true")
    a = ap.parse_args()
    import dbg_assembly_b200 as dbg
    from dbg_assembly_b200 import synth
    from dbg_assembly_b200.sharded import ShardedBuilder, SharedImage
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = 150 if a.K <= 31 else 100
    p = synth.make_params(seed=5, genome_len=300_000, read_len=L, insert=400, err=0.01, n_rate=0.001)
    per = a.reads // world
    first = rank * per
    d_bases = torch.empty(per * L, dtype=torch.uint8, device=dev)
    synth.reads_device(p, first, per, d_bases.data_ptr(), device=local)
    d_offs = torch.arange(per + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    sb = ShardedBuilder(K=a.K, max_read_len=L, init_slots=a.slots, device=local, exchange=a.exchange, sub_blocks=3)
    from dbg_assembly_b200.graph import torch_stream_handle
    sb.b.set_stream(torch_stream_handle(dev))
    half = per // 2
    occ_upper = per * L
    sb.add_reads_device(d_bases, d_offs, half, 0, half * L, first, occ_upper)
    sb.add_reads_device(d_bases, d_offs[half:], per - half, half * L, (per - half) * L, first + half, occ_upper)
    st = sb.finalize(layout=True)
    P, wide = st["array_size"], a.K > 31
    path = [f"/dev/shm/dbg_b200_mgcheck_{os.getpid()}.img"] if rank == 0 else [None]
    dist.broadcast_object_list(path, src=0)
    img = SharedImage(path[0], P, wide, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        img = SharedImage(path[0], P, wide, create=False)
    sb.export_into(img, st)
    ok, msg, fe = True, "", None
    if rank == 0:
        from oracle import oracle as orc
        hb, ho = synth.reads_host(p, 0, per * world)
        o = orc.OracleGraph(a.K, L, a.slots, 0.7, 10, 1 << 40, wide=wide)
        o.add_file(hb, ho); o.finish()
        e = o.dump()
        try:
            bits = np.unpackbits(img.nul)[:P]
            slot = np.nonzero(bits)[0].astype(np.uint64)
            sel = slot.astype(np.int64)
            assert st["global_count"] == o.count and st["global_occurrences"] == o.occurrences, (st["global_count"], o.count)
            assert np.array_equal(slot, e["slot"]), "slot layout differs from the oracle"
            assert np.array_equal(img.arr["kmer"][sel], e["kmer"]), "k-mers differ"
            if wide:
                assert np.array_equal(img.arr["kmer_hi"][sel], e["kmer_hi"])
            assert np.array_equal(img.arr["l_link"][sel], e["l"]) and np.array_equal(img.arr["r_link"][sel], e["r"]), "link words differ"
            mask = np.ones(P, dtype=bool); mask[sel] = False
            assert not img.arr["kmer"][mask].any() and not img.arr["l_link"][mask].any(), "unfilled slots are not zero"
        except AssertionError as ex:
            ok, msg = False, str(ex)
        if a.front_end and a.K <= 31:
            fe = front_end_check(world, hb, ho, a.K, L)
            if fe.get("error") or fe.get("multi_equals_single") is False or fe.get("multi_equals_reference") is False:
                ok = False
        print(json.dumps({"multigpu_check": "ok" if ok else "FAILED", "n_gpus": world, "K": a.K, "reads": per * world,
                          "nodes": int(st["global_count"]), "occurrences": st["global_occurrences"], "detail": msg,
                          "exchange_bytes_rank0": sb.exchange_bytes, "exchange": sb.exchange, "optimistic_fallbacks": sb.opt_fallbacks,
                          "front_end": fe}))
        o.close()
    dist.barrier()
    img.close(unlink=rank == 0)
    sb.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
