"""Multi-GPU check of the sharded K-mer frequency table (BASELINE config 4 shape; not collected by pytest).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_kfreq_check.py [--K 13] [--reads 200000]

The table is sharded by whole `.cz` blocks (8 Mi k-mers): rank r owns a contiguous run of blocks and writes them
itself.  Reads are dealt to the ranks in contiguous blocks, generated on the device, and **all-gathered** (1.25 B
per occurrence over NVLink) instead of exchanging k-mers (8 B per occurrence): extraction costs ~5 ps per
occurrence, so every rank re-extracts all reads and keeps the k-mers of its own index range.

Checks: every rank's 8-bit image equals the oracle's counts over its index range; the per-rank spectra add up to
the oracle's; the concatenation of the per-rank `.cz` / `.cz.len` files loads (with the restated correct_error
loader) to the oracle's bits.  With --K 17 (68.7 GB table over all ranks) the oracle side is sparse.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=13)
    ap.add_argument("--reads", type=int, default=200_000)
    ap.add_argument("--genome", type=int, default=300_000)
    ap.add_argument("--cutoff", type=int, default=1)
    a = ap.parse_args()
    from dbg_assembly_b200 import synth
    from dbg_assembly_b200.kfreq import KmerFreq
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, L = a.K, 150
    p = synth.make_params(seed=4, genome_len=a.genome, read_len=L, insert=400, err=0.01, n_rate=0.001)
    per = a.reads // world
    mine = torch.empty(per * L, dtype=torch.uint8, device=dev)
    synth.reads_device(p, rank * per, per, mine.data_ptr(), device=local)
    torch.cuda.synchronize()
    everything = torch.empty(world * per * L, dtype=torch.uint8, device=dev)
    t0 = time.perf_counter()
    dist.all_gather_into_tensor(everything, mine)
    d_offs = torch.arange(world * per + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    kf = KmerFreq(K=K, device=local, block_rank=rank, block_count=world)
    kf.submit_device(everything.data_ptr(), d_offs.data_ptr(), world * per, 0, world * per * L)
    st = kf.finalize()
    t1 = time.perf_counter()
    lo, hi = kf.index_range()
    hist = kf.histogram()
    tmp = tempfile.mkdtemp(prefix="kfreq_mg_")
    prefix = os.path.join(tmp, f"shard{rank}")
    kf.write_cz(prefix, bits=1, cutoff=a.cutoff)
    img8 = kf.export(bits=8) if K <= 14 else None
    bits = kf.export(bits=1, cutoff=0)
    nz_local = np.nonzero(np.unpackbits(bits))[0].astype(np.uint64) + np.uint64(lo) if hi > lo else np.zeros(0, np.uint64)
    cz = open(prefix + ".kmer.freq.cz", "rb").read() if hi > lo else b""
    lens = open(prefix + ".kmer.freq.cz.len").read() if hi > lo else ""
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(lo=lo, hi=hi, hist=hist, img8=img8, cz=cz, lens=lens, occ=st["occurrences"], keys=nz_local))
    ok, msg = True, ""
    if rank == 0:
        from oracle import oracle as orc
        hb, ho = synth.reads_host(p, 0, world * per)
        try:
            assert all(g["occ"] == world * per * (L - K + 1) for g in gathered), "every rank sees every occurrence"
            assert gathered[0]["lo"] == 0 and gathered[-1]["hi"] == 4 ** K
            assert all(gathered[i]["hi"] == gathered[i + 1]["lo"] for i in range(world - 1)), "index ranges tile 4^K"
            if K <= 14:
                counts = orc.kfreq_count(hb, ho, K)
                for g in gathered:
                    assert np.array_equal(g["img8"], np.minimum(counts[g["lo"]:g["hi"]], 255).astype(np.uint8)), "8-bit image"
                exp = np.bincount(np.minimum(counts[counts > 0], 65535), minlength=65536).astype(np.uint64)
                assert np.array_equal(sum(g["hist"] for g in gathered)[1:], exp[1:]), "spectrum"
                # merged files == what one context would have written; load them the way correct_error does
                mp = os.path.join(tmp, "merged")
                open(mp + ".kmer.freq.cz", "wb").write(b"".join(g["cz"] for g in gathered))
                open(mp + ".kmer.freq.cz.len", "w").write("".join(g["lens"] for g in gathered))
                both, canon = orc.load_cz_1bit(mp, K)
                assert np.array_equal(canon, (counts > a.cutoff).astype(np.uint8)), "merged .cz"
            else:
                # sparse oracle: canonical k-mers through the parse restatement on a sample of the reads
                import ctypes as C
                Lb = orc.lib()
                klo = np.zeros(L, np.uint64); lb = np.zeros(L, np.uint8); rb = np.zeros(L, np.uint8)
                u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
                got = np.concatenate([g["keys"] for g in gathered])
                assert (np.diff(got.astype(np.int64)) > 0).all(), "shards in index order, no overlap"
                sample = []
                for i in range(0, world * per, max(1, world * per // 3000)):
                    rd = hb[int(ho[i]):int(ho[i + 1])].tobytes()
                    m = Lb.orc64_parse_read(rd, len(rd), K, 65535, klo.ctypes.data_as(u64p), lb.ctypes.data_as(u8p), rb.ctypes.data_as(u8p))
                    sample.append(klo[:m].copy())
                sample = np.unique(np.concatenate(sample))
                assert np.isin(sample, got).all(), "every sampled canonical k-mer is present in its owner's range"
                species = int(sum(int(g["hist"][1:].sum()) for g in gathered))
                assert species == len(got), (species, len(got))
                n_lens = sum(len(g["lens"].split()) for g in gathered)
                assert n_lens == 4 ** K // (8 << 20), n_lens
        except AssertionError as e:
            ok, msg = False, repr(e)
        print(json.dumps({"multigpu_kfreq_check": "ok" if ok else "FAILED", "n_gpus": world, "K": K, "reads": world * per,
                          "occurrences": gathered[0]["occ"], "allgather_count_ms": round((t1 - t0) * 1e3, 2), "detail": msg}))
    kf.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
