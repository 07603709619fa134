"""Multi-GPU parity check (not collected by pytest: needs N GPUs and torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [--reads 200000] [--K 31]

Every rank extracts + exchanges + inserts its block of C2-shaped synthetic reads (ShardedBuilder); the union
of the shard dumps must equal the oracle's node multiset bit for bit, every node must sit on its owner, and
replaying the keys in first-occurrence order must reproduce the oracle's slot layout.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200_000)
    ap.add_argument("--K", type=int, default=31)
    ap.add_argument("--slots", type=int, default=12_000_000)
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer_sliced", "nccl"])
    a = ap.parse_args()
    import dbg_assembly_b200 as dbg
    from dbg_assembly_b200 import synth
    from dbg_assembly_b200.sharded import ShardedBuilder, shard_size
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
    sb = ShardedBuilder(K=a.K, max_read_len=L, init_slots=a.slots, device=local, exchange=a.exchange)
    from dbg_assembly_b200.graph import torch_stream_handle
    sb.b.set_stream(torch_stream_handle(dev))
    # two blocks per rank, to exercise repeated exchanges
    half = per // 2
    occ_upper = per * L
    sb.add_reads_device(d_bases, d_offs, half, 0, half * L, first, occ_upper)
    sb.add_reads_device(d_bases, d_offs[half:], per - half, half * L, (per - half) * L, first + half, occ_upper)
    st = sb.finalize()
    shard = sb.b.dump_shard()
    polyA = sb.b.get_polyA_counts()
    P = st["array_size"]
    # every node sits on its owner
    homes = np.array([dbg.capi.hash_code(int(k)) % P for k in shard["kmer"][:2000]]) if a.K <= 31 else None
    if homes is not None:
        ss = shard_size(P, world)
        assert ((homes // ss) == rank).all(), "node stored on the wrong shard"
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v for k, v in shard.items()})
    ok = True
    msg = ""
    if rank == 0:
        from oracle import oracle as orc
        hb, ho = synth.reads_host(p, 0, per * world)
        o = orc.OracleGraph(a.K, L, a.slots, 0.7, 10, 1 << 40, wide=a.K > 31)
        o.add_file(hb, ho); o.finish()
        e = o.dump()
        nz = (e["kmer"] != 0) | (e["kmer_hi"] != 0)
        kk = np.concatenate([g["kmer"] for g in gathered]); kh = np.concatenate([g["kmer_hi"] for g in gathered])
        ll = np.concatenate([g["l"] for g in gathered]); rr = np.concatenate([g["r"] for g in gathered]); oo = np.concatenate([g["ord"] for g in gathered])
        so = np.lexsort((kk, kh)); eo = np.lexsort((e["kmer"][nz], e["kmer_hi"][nz]))
        try:
            assert len(kk) == int(nz.sum()) == st["global_count"] - 1, (len(kk), int(nz.sum()), st["global_count"])
            assert np.array_equal(kk[so], e["kmer"][nz][eo]) and np.array_equal(kh[so], e["kmer_hi"][nz][eo])
            assert np.array_equal(ll[so], e["l"][nz][eo]) and np.array_equal(rr[so], e["r"][nz][eo])
            assert st["global_occurrences"] == o.occurrences
            pa = np.minimum(polyA, 255).astype(np.uint64)
            assert int(e["l"][~nz][0]) == (int(pa[0]) << 24 | int(pa[1]) << 16 | int(pa[2]) << 8 | int(pa[3]))
            assert int(e["r"][~nz][0]) == (int(pa[4]) << 24 | int(pa[5]) << 16 | int(pa[6]) << 8 | int(pa[7]))
            # first-occurrence order == oracle insertion order: the oracle's slot of a key is increasing in
            # "ordinal" only within a probe cluster, so check through a replay on a sample-free full pass
            order = np.argsort(oo)
            seq_lo, seq_hi = kk[order], kh[order]
            Pn = o.size
            occ = np.zeros(Pn, dtype=bool)
            slot_of = np.empty(len(seq_lo), dtype=np.int64)
            hh = [dbg.capi.hash_code_wide(int(x), int(y)) % Pn if a.K > 31 else dbg.capi.hash_code(int(x)) % Pn for x, y in zip(seq_lo.tolist(), seq_hi.tolist())]
            for i, h in enumerate(hh):
                while occ[h]:
                    h = h + 1 if h + 1 < Pn else 0
                occ[h] = True; slot_of[i] = h
            exp_slot = dict(zip(zip(e["kmer"][nz].tolist(), e["kmer_hi"][nz].tolist()), e["slot"][nz].tolist()))
            got_slot = [exp_slot[(x, y)] for x, y in zip(seq_lo.tolist(), seq_hi.tolist())]
            assert np.array_equal(slot_of, np.array(got_slot)), "ordinals do not reproduce the reference layout"
        except AssertionError as ex:
            ok, msg = False, str(ex)
        print(json.dumps({"multigpu_check": "ok" if ok else "FAILED", "n_gpus": world, "K": a.K, "reads": per * world,
                          "nodes": int(len(kk)) + 1, "occurrences": st["global_occurrences"], "detail": msg,
                          "exchange_bytes_rank0": sb.exchange_bytes, "exchange": sb.exchange}))
        o.close()
    dist.barrier()
    sb.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
