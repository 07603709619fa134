"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU build (dbg_assembly_b200/sharded.py).

The kernels need a GPU, so here the extract / insert steps are stand-ins built from the oracle's
parse_read and a python dict; what is under test is the plumbing every rank runs around them: the owner
function, the size exchange, the all-to-all(v) of tuples with ragged buckets, the poly-A all-reduce, and
that the union of the shards equals the single-process build."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO, random_reads, reads_to_arrays


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, K, R, P, out_dir):
    import sys
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dbg_assembly_b200 import capi
    from dbg_assembly_b200.sharded import Exchange, owner_of, shard_size
    from oracle import oracle as orc

    reads = random_reads(123, 600, 10, 120, genome_len=4000) + [b"A" * 50] * 7
    per = (len(reads) + world - 1) // world
    mine = reads[rank * per:(rank + 1) * per]         # contiguous block of the global read sequence
    first_read = rank * per

    # stand-in for dbg_extract_tuples_device: tuples {kmer, ord<<8 | rb<<4 | lb} bucketed by owner
    buckets = [[] for _ in range(world)]
    polyA = np.zeros(8, dtype=np.uint64)
    for i, rd in enumerate(mine):
        lo, hi, lb, rb = orc.parse_read(rd, K, R)
        for j in range(len(lo)):
            k = int(lo[j])
            if k == 0:
                if lb[j] < 4: polyA[lb[j]] += 1
                if rb[j] < 4: polyA[4 + rb[j]] += 1
                continue
            home = capi.hash_code(k) % P
            q = owner_of(home, P, world)
            meta = (((first_read + i) << 16 | j) << 8) | (int(rb[j]) << 4) | int(lb[j])
            buckets[q].append((k, meta))
    ex = Exchange()
    send_counts = torch.tensor([len(b) for b in buckets], dtype=torch.int64)
    recv_counts = ex.exchange_counts(send_counts)
    # int64 view of u64 tuples, [c, 2]
    tb = [torch.from_numpy(np.array(b, dtype=np.uint64).reshape(-1, 2).view(np.int64)) for b in buckets]
    recv, total = ex.exchange_payload(tb, recv_counts.tolist(), 2, torch.zeros(1, dtype=torch.int64))
    assert total == int(recv_counts.sum())
    got = recv.numpy().view(np.uint64)
    # stand-in for dbg_insert_tuples_device: this rank's shard
    lo_slot, hi_slot = rank * shard_size(P, world), min(P, (rank + 1) * shard_size(P, world))
    shard = {}
    for k, meta in got.tolist():
        home = capi.hash_code(k) % P
        assert lo_slot <= home < hi_slot, "tuple routed to the wrong owner"
        lbv, rbv, od = meta & 15, (meta >> 4) & 15, meta >> 8
        e = shard.setdefault(k, [[0] * 4, [0] * 4, od])
        if lbv < 4: e[0][lbv] = min(255, e[0][lbv] + 1)
        if rbv < 4: e[1][rbv] = min(255, e[1][rbv] + 1)
        e[2] = min(e[2], od)
    polyA_sum = ex.allreduce_sum_u64(polyA, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"),
             kmer=np.array(sorted(shard), dtype=np.uint64),
             l=np.array([sum(min(c, 255) << (24 - 8 * b) for b, c in enumerate(shard[k][0])) for k in sorted(shard)], dtype=np.uint32),
             r=np.array([sum(min(c, 255) << (24 - 8 * b) for b, c in enumerate(shard[k][1])) for k in sorted(shard)], dtype=np.uint32),
             ord=np.array([shard[k][2] for k in sorted(shard)], dtype=np.uint64), polyA=polyA_sum, n_reads=len(reads))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_exchange_equals_single_process(tmp_path, oracle_mod):
    K, R = 21, 100
    from dbg_assembly_b200 import capi
    P = capi.find_next_prime(60_000)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), K, R, P, str(tmp_path)), nprocs=world, join=True)
    shards = [np.load(tmp_path / f"shard{r}.npz") for r in range(world)]
    reads = random_reads(123, 600, 10, 120, genome_len=4000) + [b"A" * 50] * 7
    assert int(shards[0]["n_reads"]) == len(reads)
    o = oracle_mod.OracleGraph(K, R, 60_000, 0.7, 10, 1 << 40)
    o.add_file(*reads_to_arrays(reads)); o.finish()
    e = o.dump()
    nz = e["kmer"] != 0
    order = np.argsort(e["kmer"][nz])
    kk = np.concatenate([s["kmer"] for s in shards]); ll = np.concatenate([s["l"] for s in shards]); rr = np.concatenate([s["r"] for s in shards])
    so = np.argsort(kk)
    assert np.array_equal(kk[so], e["kmer"][nz][order])
    assert np.array_equal(ll[so], e["l"][nz][order]) and np.array_equal(rr[so], e["r"][nz][order])
    # shards are disjoint, poly-A counters agree on both ranks and reproduce the oracle's k-mer-0 node
    assert len(np.intersect1d(shards[0]["kmer"], shards[1]["kmer"])) == 0
    assert np.array_equal(shards[0]["polyA"], shards[1]["polyA"])
    pa = np.minimum(shards[0]["polyA"], 255).astype(np.uint64)
    exp_l = int(pa[0]) << 24 | int(pa[1]) << 16 | int(pa[2]) << 8 | int(pa[3])
    exp_r = int(pa[4]) << 24 | int(pa[5]) << 16 | int(pa[6]) << 8 | int(pa[7])
    assert int(e["l"][~nz][0]) == exp_l and int(e["r"][~nz][0]) == exp_r
    # first-occurrence ordinals order the keys exactly like the oracle's sequential insertion: replaying the
    # keys in ordinal order through linear probing must reproduce the oracle's slot layout
    oo = np.concatenate([s["ord"] for s in shards])
    seq = kk[np.argsort(oo)]
    P_ = o.size
    table = {}
    for k in seq.tolist():
        h = capi.hash_code(k) % P_
        while h in table:
            h = (h + 1) % P_
        table[h] = k
    h = capi.hash_code(0) % P_
    while h in table:
        h = (h + 1) % P_
    table[h] = 0
    slots = np.array(sorted(table), dtype=np.uint64)
    assert np.array_equal(slots, e["slot"])
    assert np.array_equal(np.array([table[s] for s in sorted(table)], dtype=np.uint64), e["kmer"])
    o.close()


def test_owner_function_matches_library_shard_ranges():
    from dbg_assembly_b200.sharded import owner_of, shard_size
    for P, n in ((1009, 2), (200000033, 8), (100000007, 3), (7, 4)):
        ss = shard_size(P, n)
        assert ss * n >= P
        for home in (0, ss - 1, ss, P - 1):
            if home < P:
                q = owner_of(home, P, n)
                assert 0 <= q < n and q * ss <= home < (q + 1) * ss
