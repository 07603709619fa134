"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(include/dbg_b200.h via dbg_assembly_b200.capi); the oracle (oracle/) is only the checker.

Bar: bit-exact -- same nodes, same (l_link, r_link), same SLOT LAYOUT as the reference run with -t 1.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_NO_ENLARGE, REPO, load_golden, random_reads, reads_to_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dbg():
    import dbg_assembly_b200 as m
    if m.capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (no CPU fallback exists)")
    return m


@pytest.fixture(params=["direct", "partitioned"], autouse=True)
def build_path(request, monkeypatch):
    """every test runs through both build paths: the fused direct-insert kernel, and the radix-partitioned
    path (count / scan / scatter / insert) forced on with tiny buckets and tiny host batches so that many
    buckets, many batches and the double-buffered batch hand-off are exercised even at test sizes"""
    if request.param == "direct":
        monkeypatch.setenv("DBG_B200_PARTITION", "0")
    else:
        monkeypatch.setenv("DBG_B200_PARTITION", "1")
        monkeypatch.setenv("DBG_B200_PART_SHIFT", "8")
        monkeypatch.setenv("DBG_B200_BATCH_BASES", "150000")
        monkeypatch.setenv("DBG_B200_BATCH_READS", "3000")
    return request.param


def gpu_build(dbg, files, K, R, init_slots, load=0.7, track=True, force_wide=False):
    with dbg.DBGBuilder(K=K, max_read_len=R, init_slots=init_slots, load_factor=load, track_order=track,
                        force_wide=force_wide) as b:
        for bases, offs in files:
            b.submit(bases, offs)
        st = b.finalize()
        arr, nul = b.export_kmerset()
        return st, arr, nul


def image_to_dump(arr, nul, P):
    bits = np.unpackbits(nul)[:P]
    slot = np.nonzero(bits)[0].astype(np.uint64)
    sel = slot.astype(np.int64)
    d = dict(slot=slot, kmer=arr["kmer"][sel], l=arr["l_link"][sel], r=arr["r_link"][sel])
    if "kmer_hi" in arr.dtype.names:
        d["kmer_hi"] = arr["kmer_hi"][sel]
    # everything outside the filled slots must be zero (reference: calloc'ed array)
    mask = np.ones(P, dtype=bool); mask[sel] = False
    assert not arr["kmer"][mask].any() and not arr["l_link"][mask].any() and not arr["r_link"][mask].any()
    return d


def oracle_build(orc, files, K, R, init_slots, load=0.7, wide=False):
    o = orc.OracleGraph(K, R, init_slots, load, 10, 1 << 40, wide=wide)   # one block per file: never enlarges
    for bases, offs in files:
        o.add_file(bases, offs)
    o.finish()
    return o


# ---------------------------------------------------------------------------------------------------
# golden vectors produced by the reference itself
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_NO_ENLARGE)
@pytest.mark.parametrize("force_wide", [False, True])
def test_golden_reference_tables(dbg, name, force_wide):
    g = load_golden(name)
    st, arr, nul = gpu_build(dbg, g["files"], g["K"], g["R"], g["init_slots"], g["load"], force_wide=force_wide)
    assert st["array_size"] == g["size"]
    assert st["max_cutoff"] == g["max"]
    assert st["count"] == g["count"]
    assert st["reads"] == g["reads"]
    assert st["kmers_logged"] == g["kmers_logged"]
    d = image_to_dump(arr, nul, g["size"])
    assert np.array_equal(d["slot"], g["slot"]), "slot layout differs from the reference (-t 1)"
    assert np.array_equal(d["kmer"], g["kmer"])
    assert np.array_equal(d["l"], g["l"])
    assert np.array_equal(d["r"], g["r"])
    if force_wide:
        assert not d["kmer_hi"].any()


@pytest.mark.parametrize("name", ["ragged_k31", "saturate_k21", "tiny_k3", "contig_k31"])
def test_golden_tables_global_layout_method(dbg, name, monkeypatch):
    """the fallback layout method (atomicMin priority probing over the whole table) gives the same image as
    the default cluster-local method"""
    monkeypatch.setenv("DBG_B200_LAYOUT", "global")
    g = load_golden(name)
    st, arr, nul = gpu_build(dbg, g["files"], g["K"], g["R"], g["init_slots"], g["load"])
    d = image_to_dump(arr, nul, g["size"])
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(d[k], g[k]), k


@pytest.mark.parametrize("version", ["1", "2"])
@pytest.mark.parametrize("force_wide", [False, True])
@pytest.mark.parametrize("name", ["ragged_k31", "saturate_k21", "contig_k31", "even_k16_two_files"])
def test_both_versions_of_the_cluster_layout_pass(dbg, name, force_wide, version, monkeypatch):
    """k_layout_clusters (version 1, the default for 64-B nodes) and k_layout_clusters2 (version 2, the default for 32-B
    nodes) on both node widths: the reference's slot layout either way"""
    monkeypatch.setenv("DBG_B200_LAYOUT_V", version)
    g = load_golden(name)
    st, arr, nul = gpu_build(dbg, g["files"], g["K"], g["R"], g["init_slots"], g["load"], force_wide=force_wide)
    d = image_to_dump(arr, nul, g["size"])
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(d[k], g[k]), k


def test_dense_table_long_clusters_and_wraparound(dbg, oracle_mod):
    """load factor 0.97: clusters of hundreds of slots (region path / global fallback) and a probe chain that
    runs over the end of the table into the first slots"""
    reads = random_reads(81, 1500, 60, 100, genome_len=20000, err=0.03)
    files = [reads_to_arrays(reads)]
    probe = oracle_build(oracle_mod, files, 31, 100, 2_000_000)
    n_nodes = probe.count
    probe.close()
    for load in (0.97, 0.999):
        init_slots = int(n_nodes / load)
        o = oracle_build(oracle_mod, files, 31, 100, init_slots)
        st, arr, nul = gpu_build(dbg, files, 31, 100, init_slots)
        d, e = image_to_dump(arr, nul, o.size), o.dump()
        for k in ("slot", "kmer", "l", "r"):
            assert np.array_equal(d[k], e[k]), (load, k)
        o.close()


# ---------------------------------------------------------------------------------------------------
# seeded random inputs against the oracle
# ---------------------------------------------------------------------------------------------------
CASES = [
    # seed, K, R, n_reads, len_lo, len_hi, genome, load target
    (11, 31, 150, 2000, 20, 200, 20000, 0.5),
    (12, 31, 100, 3000, 90, 110, 3000, 0.9),     # heavy collisions, long clusters, wrap-around
    (13, 15, 80, 1500, 1, 100, 5000, 0.6),
    (14, 16, 60, 1500, 10, 90, 2000, 0.6),       # even K: palindromes
    (15, 1, 20, 200, 0, 30, 100, 0.3),
    (16, 2, 20, 200, 0, 30, 100, 0.5),
    (17, 27, 250, 800, 200, 400, 30000, 0.4),    # reads longer than -r get trimmed
    (18, 32, 120, 1500, 20, 150, 8000, 0.6),     # 128-bit path from here on
    (19, 33, 120, 1500, 20, 150, 8000, 0.6),
    (20, 47, 150, 1500, 40, 180, 8000, 0.8),
    (21, 63, 100, 3000, 60, 120, 10000, 0.6),
]


@pytest.mark.parametrize("seed,K,R,n_reads,len_lo,len_hi,genome,load_target", CASES)
def test_random_reads_match_oracle(dbg, oracle_mod, seed, K, R, n_reads, len_lo, len_hi, genome, load_target):
    reads = random_reads(seed, n_reads, len_lo, len_hi, genome_len=genome)
    reads += [b"", b"A" * (K + 3), b"T" * (K + 1), b"ACGT" * 20, b"N" * (K + 2)]
    files = [reads_to_arrays(reads[: len(reads) // 2]), reads_to_arrays(reads[len(reads) // 2:])]
    wide = K > 31
    probe = oracle_build(oracle_mod, files, K, R, 50_000_000 // 100, wide=wide)
    n_nodes = probe.count
    probe.close()
    init_slots = max(3, int(n_nodes / load_target))
    o = oracle_build(oracle_mod, files, K, R, init_slots, wide=wide)
    st, arr, nul = gpu_build(dbg, files, K, R, init_slots)
    assert st["array_size"] == o.size and st["count"] == o.count and st["max_cutoff"] == o.max
    assert st["occurrences"] == o.occurrences and st["kmers_logged"] == o.kmers_logged
    d = image_to_dump(arr, nul, o.size)
    e = o.dump()
    assert np.array_equal(d["slot"], e["slot"])
    assert np.array_equal(d["kmer"], e["kmer"])
    if wide:
        assert np.array_equal(d["kmer_hi"], e["kmer_hi"])
    assert np.array_equal(d["l"], e["l"]) and np.array_equal(d["r"], e["r"])
    # full image equality incl. the nul_flag bytes
    assert np.array_equal(nul, o.nul_flag())
    o.close()


def test_blocks_and_subblocks_do_not_change_the_result(dbg, oracle_mod, monkeypatch):
    """many small submits, and forced tiny internal sub-blocks (double buffering), equal one big submit"""
    reads = random_reads(31, 4000, 30, 160, genome_len=15000)
    bases, offs = reads_to_arrays(reads)
    o = oracle_build(oracle_mod, [(bases, offs)], 31, 150, 400_000)
    e = o.dump()
    for sub_bases, step in ((None, 4000), (None, 37), (4096, 4000), (1000, 500)):
        if sub_bases:
            monkeypatch.setenv("DBG_B200_SUB_BASES", str(sub_bases))
        else:
            monkeypatch.delenv("DBG_B200_SUB_BASES", raising=False)
        with dbg.DBGBuilder(K=31, max_read_len=150, init_slots=400_000) as b:
            for r0 in range(0, len(reads), step):
                r1 = min(len(reads), r0 + step)
                b.submit(bases, offs[r0:r1 + 1])      # offsets keep indexing the whole base array
            st = b.finalize()
            arr, nul = b.export_kmerset()
        d = image_to_dump(arr, nul, o.size)
        for k in ("slot", "kmer", "l", "r"):
            assert np.array_equal(d[k], e[k]), (sub_bases, step, k)
        assert st["occurrences"] == o.occurrences
    o.close()


def test_saturation_and_polyA_under_contention(dbg, oracle_mod):
    """thousands of concurrent updates of the same few nodes must still give min(255, count) per lane"""
    reads = [b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTAC"] * 3000 + [b"A" * 80] * 2000 + [b"T" * 40] * 100 + [b"G" * 64] * 1500
    reads += [b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAC", b"GAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"] * 7
    files = [reads_to_arrays(reads)]
    for K in (21, 31, 35):
        o = oracle_build(oracle_mod, files, K, 100, 5000, wide=K > 31)
        st, arr, nul = gpu_build(dbg, files, K, 100, 5000)
        d, e = image_to_dump(arr, nul, o.size), o.dump()
        for k in ("slot", "kmer", "l", "r"):
            assert np.array_equal(d[k], e[k]), (K, k)
        assert (e["l"] == 0xFF000000).any() or (e["l"] >> 24 == 255).any()   # the case really saturates
        o.close()


def test_untracked_mode_same_nodes_valid_layout(dbg, oracle_mod):
    """track_order=0: same node multiset; layout is a valid linear-probing layout (every key reachable
    from hash%P without crossing an empty slot), like the reference with -t > 1"""
    reads = random_reads(41, 3000, 50, 150, genome_len=10000)
    files = [reads_to_arrays(reads)]
    o = oracle_build(oracle_mod, files, 31, 150, 150_000)
    e = o.dump()
    st, arr, nul = gpu_build(dbg, files, 31, 150, 150_000, track=False)
    d = image_to_dump(arr, nul, o.size)
    assert np.array_equal(d["slot"], e["slot"])       # occupancy of linear probing is order independent
    a, b = np.argsort(d["kmer"]), np.argsort(e["kmer"])
    for k in ("kmer", "l", "r"):
        assert np.array_equal(d[k][a], e[k][b])
    P = o.size
    filled = np.zeros(P, dtype=bool); filled[d["slot"].astype(np.int64)] = True
    L = oracle_mod.lib()
    for s, k in zip(d["slot"].tolist()[:5000], d["kmer"].tolist()[:5000]):
        h = L.orc_hash_code(k) % P
        while h != s:
            assert filled[h]
            h = (h + 1) % P
    o.close()


def test_empty_input_still_has_the_polyA_node(dbg, oracle_mod):
    """DBGgraph.cpp:418: the k-mer-0 node is inserted always, even when nothing was read"""
    empty = (np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    st, arr, nul = gpu_build(dbg, [empty], 31, 100, 1000)
    o = oracle_build(oracle_mod, [empty], 31, 100, 1000)
    d, e = image_to_dump(arr, nul, o.size), o.dump()
    assert st["count"] == 1 == o.count
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(d[k], e[k])
    # reads all shorter than K
    short = reads_to_arrays([b"ACG", b"", b"ACGTACGTAC"])
    st, arr, nul = gpu_build(dbg, [short], 31, 100, 1000)
    assert st["count"] == 1 and st["occurrences"] == 0 and st["reads"] == 3
    o.close()


def test_links_pass_and_compact_dump(dbg, oracle_mod):
    """calculate_kmer_links (contig.cpp:107-205) on the device + slot-ordered survivor dump"""
    g = load_golden("contig_k31")
    o = oracle_build(oracle_mod, g["files"], g["K"], g["R"], g["init_slots"], g["load"])
    with dbg.DBGBuilder(K=g["K"], max_read_len=g["R"], init_slots=g["init_slots"], load_factor=g["load"]) as b:
        for bases, offs in g["files"]:
            b.submit(bases, offs)
        b.finalize()
        for cutoff in (2, 0, 7):
            lk, ek = b.export_links(cutoff), o.kmer_links(cutoff)
            assert np.array_equal(lk["klink"], ek["klink"])
            assert np.array_equal(lk["del_flag"], ek["del_flag"])
            assert np.array_equal(lk["depth_stat"], ek["depth_stat"])
            assert np.array_equal(lk["tips"], ek["tips"]) and np.array_equal(lk["branches"], ek["branches"])
            assert (lk["total"], lk["deleted"], lk["linear"]) == (ek["total"], ek["deleted"], ek["linear"])
            dump = b.dump_compact(cutoff)
            e = o.dump()
            bits = np.unpackbits(ek["del_flag"])[: o.size]
            keep = bits[e["slot"].astype(np.int64)] == 0
            for k in ("slot", "kmer", "l", "r"):
                assert np.array_equal(dump[k], e[k][keep]), (cutoff, k)
        full = b.dump_compact(-1)
        e = o.dump()
        for k in ("slot", "kmer", "l", "r"):
            assert np.array_equal(full[k], e[k])
    # the reference's own .contig.kmer.freq rows equal the device histogram
    txt = g["file_contig_kmer_freq"].tobytes().decode().splitlines()
    rows = [tuple(map(int, line.split("\t"))) for line in txt[1:]]
    with dbg.DBGBuilder(K=g["K"], max_read_len=g["R"], init_slots=g["init_slots"], load_factor=g["load"]) as b:
        for bases, offs in g["files"]:
            b.submit(bases, offs)
        b.finalize()
        lk = b.export_links(2, lists=False)
    assert rows == [(d, int(lk["depth_stat"][d])) for d in range(1, 256)]
    o.close()


def test_device_resident_input_and_synth_parity(dbg, oracle_mod):
    """reads generated in HBM by the device generator == host generator; submit_device == submit"""
    import torch
    from dbg_assembly_b200 import synth
    p = synth.make_params(seed=7, genome_len=200_000, read_len=150, insert=400, err=0.01, n_rate=0.001)
    n = 40_000
    hb, ho = synth.reads_host(p, 0, n)
    db = torch.empty(n * 150, dtype=torch.uint8, device="cuda")
    synth.reads_device(p, 0, n, db.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(db.cpu().numpy(), hb)
    do = torch.from_numpy(ho.astype(np.int64)).cuda()
    o = oracle_build(oracle_mod, [(hb, ho)], 31, 150, 2_000_000)
    e = o.dump()
    with dbg.DBGBuilder(K=31, max_read_len=150, init_slots=2_000_000) as b:
        # two device blocks, the second one starting at a non-16-aligned base offset
        half = 20_001
        b.submit_device(db.data_ptr(), do.data_ptr(), half, 0, half * 150)
        b.submit_device(db.data_ptr(), do.data_ptr() + 8 * half, n - half, half * 150, (n - half) * 150)
        st = b.finalize()
        arr, nul = b.export_kmerset()
        assert b.launches >= 4
    d = image_to_dump(arr, nul, o.size)
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(d[k], e[k])
    assert st["occurrences"] == n * 120 == o.occurrences
    o.close()


def test_reset_rebuilds_identically(dbg):
    reads = random_reads(51, 2000, 40, 120)
    files = [reads_to_arrays(reads)]
    with dbg.DBGBuilder(K=25, max_read_len=120, init_slots=200_000) as b:
        b.submit(*files[0]); b.finalize(); a1, n1 = b.export_kmerset()
        b.reset()
        b.submit(*files[0]); st = b.finalize(); a2, n2 = b.export_kmerset()
    assert np.array_equal(a1, a2) and np.array_equal(n1, n2) and st["reads"] == len(reads)


def test_table_full_is_reported_not_hung(dbg):
    reads = random_reads(61, 3000, 100, 150, genome_len=100000, err=0.05)
    bases, offs = reads_to_arrays(reads)
    with dbg.DBGBuilder(K=31, max_read_len=150, init_slots=5000) as b:   # ~3e5 distinct k-mers into 5003 slots (+margin)
        with pytest.raises(dbg.capi.DbgError) as ei:      # reported by the submit that follows the overfull batch, or by finalize
            b.submit(bases, offs)
            b.finalize()
        assert ei.value.code == dbg.capi.DBG_ERR_TABLE_FULL


def test_table_full_is_reported_promptly_on_a_large_table(dbg, build_path, monkeypatch):
    """a table of 2 M slots fed 6 M distinct k-mers in several host batches: DBG_ERR_TABLE_FULL comes back from a
    submit (or the finalize) within seconds -- probes are capped, a full table ends the kernels early, and the counters
    travel to the host behind every batch"""
    import time
    from dbg_assembly_b200 import synth
    if build_path != "direct":
        pytest.skip("batching set by this test")
    monkeypatch.setenv("DBG_B200_BATCH_BASES", "40000000")
    p = synth.make_params(seed=5, genome_len=50_000_000, read_len=150, insert=400, err=0.0, n_rate=0.0)
    hb, ho = synth.reads_host(p, 0, 400_000)
    t0 = time.perf_counter()
    with pytest.raises(dbg.capi.DbgError) as ei:
        with dbg.DBGBuilder(K=31, max_read_len=150, init_slots=2_000_000) as b:
            for r0 in range(0, 400_000, 50_000):
                b.submit(hb[r0 * 150:(r0 + 50_000) * 150], ho[: 50_001])
            b.finalize()
    assert ei.value.code == dbg.capi.DBG_ERR_TABLE_FULL
    assert time.perf_counter() - t0 < 60


@pytest.mark.parametrize("optimistic,K", [("1", 31), ("0", 31), ("overflow", 31), ("1", 63), ("overflow", 63)])
def test_sharded_tuple_path_single_gpu(dbg, oracle_mod, monkeypatch, optimistic, K):
    """multi-GPU building blocks on one device: extract tuples bucketed by owner shard, insert each bucket
    into that shard's context, union of shard dumps == oracle node multiset (ranks emulated in sequence).
    The owner-side partition of the received tuples runs optimistically (fixed bucket regions), exactly, and
    through the overflow fallback (regions forced too small)."""
    import torch
    if optimistic == "overflow":
        monkeypatch.setenv("DBG_B200_OPT_CAPB", "512")
    else:
        monkeypatch.setenv("DBG_B200_OPTIMISTIC", optimistic)
    reads = random_reads(71, 6000, 40, 150, genome_len=30000) + [b"A" * 60] * 50
    bases, offs = reads_to_arrays(reads)
    P_req = 300_000
    wide = K > 31
    tb = 32 if wide else 16
    o = oracle_build(oracle_mod, [(bases, offs)], K, 150, P_req, wide=wide)
    e = o.dump()
    db = torch.from_numpy(bases).cuda()
    do = torch.from_numpy(offs.astype(np.int64)).cuda()
    n_occ = o.occurrences
    for n_parts in (2, 3):
        shards = [dbg.DBGBuilder(K=K, max_read_len=150, init_slots=P_req, shard_rank=r, shard_count=n_parts) for r in range(n_parts)]
        cap = int(offs[-1])
        tuples = torch.empty(cap * (tb // 8), dtype=torch.int64, device="cuda")
        counts = torch.zeros(n_parts, dtype=torch.int64, device="cuda")
        # every rank extracts its own half of the reads (here: one extractor context does both halves)
        half = len(reads) // 2
        polyA = np.zeros(8, dtype=np.uint64)
        for (r0, r1) in ((0, half), (half, len(reads))):
            ex = shards[0]
            ex.extract_tuples_device(db.data_ptr(), do.data_ptr() + 8 * r0, r1 - r0, int(offs[r0]), int(offs[r1] - offs[r0]), r0,
                                     n_parts, tuples.data_ptr(), cap, counts.data_ptr())
            torch.cuda.synchronize()
            c = counts.cpu().numpy()
            off = 0
            for q in range(n_parts):          # tuples are packed by owner: offsets = prefix sums of the counts
                shards[q].insert_tuples_device(tuples.data_ptr() + tb * off, int(c[q]))
                off += int(c[q])
            torch.cuda.synchronize()
        polyA += shards[0].get_polyA_counts()
        total_nodes = 0
        merged = {"kmer": [], "kmer_hi": [], "l": [], "r": []}
        for q, s in enumerate(shards):
            st = s.finalize()
            total_nodes += st["count"]
            d = s.dump_shard()
            assert len(d["kmer"]) == st["count"]
            if not wide:
                homes = np.array([dbg.capi.hash_code(int(k)) % o.size for k in d["kmer"][:500].tolist()], dtype=np.int64)
                assert ((homes // ((o.size + n_parts - 1) // n_parts)) == q).all()
            for k in merged:
                merged[k].append(d[k])
            s.close()
        assert total_nodes == o.count - 1            # shards do not carry the k-mer-0 node
        nzm = (e["kmer"] != 0) | (e["kmer_hi"] != 0)
        so = np.lexsort((np.concatenate(merged["kmer"]), np.concatenate(merged["kmer_hi"])))
        eo = np.lexsort((e["kmer"][nzm], e["kmer_hi"][nzm]))
        for k in ("kmer", "kmer_hi", "l", "r"):
            assert np.array_equal(np.concatenate(merged[k])[so], e[k][nzm][eo]), (n_parts, k)
        exp_l = np.minimum(polyA[:4], 255); exp_r = np.minimum(polyA[4:], 255)
        zero = ~nzm
        assert int(e["l"][zero][0]) == int(exp_l[0]) << 24 | int(exp_l[1]) << 16 | int(exp_l[2]) << 8 | int(exp_l[3])
        assert int(e["r"][zero][0]) == int(exp_r[0]) << 24 | int(exp_r[1]) << 16 | int(exp_r[2]) << 8 | int(exp_r[3])
    # tuples of ALL occurrences into one unsharded context == the fused path == the oracle (layout too)
    with dbg.DBGBuilder(K=K, max_read_len=150, init_slots=P_req) as ex, dbg.DBGBuilder(K=K, max_read_len=150, init_slots=P_req) as ins:
        cap = int(offs[-1])
        tuples = torch.empty(cap * (tb // 8), dtype=torch.int64, device="cuda")
        counts = torch.zeros(1, dtype=torch.int64, device="cuda")
        ex.extract_tuples_device(db.data_ptr(), do.data_ptr(), len(reads), 0, int(offs[-1]), 0, 1, tuples.data_ptr(), cap, counts.data_ptr())
        torch.cuda.synchronize()
        ins.insert_tuples_device(tuples.data_ptr(), int(counts.item()))
        ins.set_polyA_counts(ex.get_polyA_counts())
        st = ins.finalize()
        arr, nul = ins.export_kmerset()
    d = image_to_dump(arr, nul, o.size)
    for k in ("slot", "kmer", "l", "r") + (("kmer_hi",) if wide else ()):
        assert np.array_equal(d[k], e[k])
    o.close()


def test_medium_synthetic_matches_oracle(dbg, oracle_mod):
    """C2-shaped reads (PE150, 1 % errors, 0.1 % N, K=31) at a size the oracle finishes in seconds"""
    from dbg_assembly_b200 import synth
    p = synth.make_params(seed=2, genome_len=460_000, read_len=150, insert=500, err=0.01, n_rate=0.001)
    n = 306_666
    hb, ho = synth.reads_host(p, 0, n)
    init_slots = 20_000_000
    o = oracle_build(oracle_mod, [(hb, ho)], 31, 150, init_slots)
    st, arr, nul = gpu_build(dbg, [(hb, ho)], 31, 150, init_slots)
    assert st["count"] == o.count and st["occurrences"] == o.occurrences == n * 120
    e = o.dump()
    d = image_to_dump(arr, nul, o.size)
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(d[k], e[k])
    o.close()


@pytest.mark.parametrize("K", [31, 55])
@pytest.mark.parametrize("variant", ["optimistic", "overflow_fallback", "exact", "unstaged"])
def test_partition_variants_match_oracle(dbg, oracle_mod, build_path, monkeypatch, variant, K):
    """the partitioned build's variants on the medium C2-shaped case: the optimistic single-pass partition (fixed bucket
    regions), its overflow fallback (regions forced too small, so every block is redone by the exact two-pass
    partition -- and the side counters must not be double counted), the exact partition alone, and the scatter
    that stores tuples one by one; all bit-identical to the oracle, layout included"""
    if build_path != "partitioned":
        pytest.skip("partitioned path only")
    for k in ("DBG_B200_PART_SHIFT", "DBG_B200_BATCH_BASES", "DBG_B200_BATCH_READS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("DBG_B200_PART_SHIFT", "14")
    if variant == "overflow_fallback":
        monkeypatch.setenv("DBG_B200_OPT_CAPB", "1000")
    elif variant == "exact":
        monkeypatch.setenv("DBG_B200_OPTIMISTIC", "0")
    elif variant == "unstaged":
        monkeypatch.setenv("DBG_B200_STAGE_CAP", "0")
    from dbg_assembly_b200 import synth
    p = synth.make_params(seed=12, genome_len=300_000, read_len=150, insert=500, err=0.01, n_rate=0.002)
    n = 120_000
    hb, ho = synth.reads_host(p, 0, n)
    # poly-A reads: the k-mer-0 side counters are bumped by the scatter pass, which the fallback must undo
    extra = [b"A" * 150] * 40
    eb = np.frombuffer(b"".join(extra), dtype=np.uint8)
    hb = np.concatenate([hb, eb]); ho = np.concatenate([ho, ho[-1] + 150 * np.arange(1, len(extra) + 1, dtype=ho.dtype)])
    init_slots = 12_000_000
    o = oracle_build(oracle_mod, [(hb, ho)], K, 150, init_slots, wide=K > 31)
    with dbg.DBGBuilder(K=K, max_read_len=150, init_slots=init_slots, load_factor=0.7) as b:
        half = (n // 2)
        b.submit(hb, ho[: half + 1])
        b.submit(hb, ho[half:])
        st = b.finalize()
        arr, nul = b.export_kmerset()
        pc = b.path_counts()
    assert pc["direct"] == 0
    if variant == "optimistic":
        assert pc["optimistic"] >= 1 and pc["overflows"] == 0 and pc["exact"] == 0, pc
    elif variant == "overflow_fallback":
        assert pc["overflows"] >= 1 and pc["exact"] == pc["overflows"] and pc["optimistic"] == 0, pc
    else:
        assert pc["exact"] >= 1 and pc["optimistic"] == 0 and pc["overflows"] == 0, pc
    assert st["count"] == o.count and st["occurrences"] == o.occurrences and st["kmers_logged"] == o.kmers_logged
    e = o.dump()
    d = image_to_dump(arr, nul, o.size)
    for k in ("slot", "kmer", "l", "r") + (("kmer_hi",) if K > 31 else ()):
        assert np.array_equal(d[k], e[k]), k
    o.close()


@pytest.mark.parametrize("K", [31, 55])
@pytest.mark.parametrize("variant", ["optimistic", "overflow_fallback"])
def test_pipelined_submit_matches_oracle(dbg, oracle_mod, build_path, monkeypatch, variant, K):
    """a dbg_submit_reads call that is a partitioned block by itself is pipelined: the scatter of sub-block i runs while
    sub-block i+1 is copied, one insert follows (submit_pipelined, csrc/dbg_build.cu).  Small sub-blocks make every call a
    dozen launches; a second call appends to a table that already holds nodes; the overflow variant forces the regions too
    small, so the whole resident batch is redone by the exact partition.  Bit-identical to the oracle, layout included."""
    if build_path != "partitioned":
        pytest.skip("partitioned path only")
    for k in ("DBG_B200_PART_SHIFT", "DBG_B200_BATCH_BASES", "DBG_B200_BATCH_READS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("DBG_B200_PART_SHIFT", "14")
    monkeypatch.setenv("DBG_B200_SUB_BASES", "700000")
    monkeypatch.setenv("DBG_B200_SUB_READS", "5000")
    if variant == "overflow_fallback":
        monkeypatch.setenv("DBG_B200_OPT_CAPB", "1000")
    from dbg_assembly_b200 import synth
    p = synth.make_params(seed=21, genome_len=300_000, read_len=150, insert=500, err=0.01, n_rate=0.002)
    n = 100_000
    hb, ho = synth.reads_host(p, 0, n)
    extra = [b"A" * 150] * 40 + [b"ACGT" * 10]          # poly-A side counters; a read shorter than K=55
    eb = np.frombuffer(b"".join(extra), dtype=np.uint8)
    hb = np.concatenate([hb, eb])
    ho = np.concatenate([ho, ho[-1] + np.cumsum(np.array([len(x) for x in extra], dtype=ho.dtype))])
    n_all = len(ho) - 1
    init_slots = 10_000_000
    o = oracle_build(oracle_mod, [(hb, ho)], K, 150, init_slots, wide=K > 31)
    launches = {}
    for pipe in ("1", "0"):
        monkeypatch.setenv("DBG_B200_PIPELINE", pipe)
        with dbg.DBGBuilder(K=K, max_read_len=150, init_slots=init_slots, load_factor=0.7) as b:
            cut = n_all * 2 // 3
            b.submit(hb, ho[: cut + 1])
            b.submit(hb, ho[cut:])
            st = b.finalize()
            arr, nul = b.export_kmerset()
            pc = b.path_counts()
            launches[pipe] = b.launches
        assert pc["direct"] == 0
        blocks = 2 if pipe == "1" else 1       # pipelined: every call is its own block; otherwise both calls share one batch
        if variant == "optimistic":
            assert pc["optimistic"] == blocks and pc["overflows"] == 0 and pc["exact"] == 0, (pc, pipe)
        else:
            assert pc["overflows"] == blocks and pc["exact"] == blocks and pc["optimistic"] == 0, (pc, pipe)
        assert st["count"] == o.count and st["occurrences"] == o.occurrences and st["kmers_logged"] == o.kmers_logged
        assert st["reads"] == n_all
        e = o.dump()
        d = image_to_dump(arr, nul, o.size)
        for k in ("slot", "kmer", "l", "r") + (("kmer_hi",) if K > 31 else ()):
            assert np.array_equal(d[k], e[k]), (k, pipe)
    assert launches["1"] > launches["0"] + 10, launches      # the pipelined calls really ran sub-block by sub-block
    o.close()


FINISH_CASES = [
    # name, K, reads, init_slots, groups, earlier block, expect
    ("medium_k31", 31, 100_000, 10_000_000, 8, False, "grouped"),
    ("medium_k55", 55, 100_000, 10_000_000, 5, False, "grouped"),
    ("second_block", 31, 100_000, 10_000_000, 8, True, "grouped"),
    ("wrap_and_long_clusters", 31, 12_000, 1_000_200, 3, False, "grouped"),       # 2 keys wrap past slot P-1, 42 clusters > 64 slots
    ("dense_patches", 31, 12_000, 700_700, 2, False, "grouped"),                   # load 0.96: ~1800 long clusters re-copied
    ("too_dense_redone", 31, 36_000, 1_600_000, 4, False, "plain"),                # > 4096 long clusters: one-pass layout + plain copy
    ("small_call", 31, 1_500, 1_000_200, 3, False, "plain"),                       # not a partitioned block: the plain sequence
]


@pytest.mark.parametrize("name,K,n,init_slots,groups,earlier,expect", FINISH_CASES, ids=[c[0] for c in FINISH_CASES])
def test_finish_export_matches_oracle(dbg, oracle_mod, build_path, monkeypatch, name, K, n, init_slots, groups, earlier, expect):
    """dbg_finish_export = submit + finalize + export in one call; where the last block is a partitioned block it runs
    insert -> layout -> copy slice group by slice group (grouped_finish, csrc/dbg_build.cu).  The table image handed to
    the caller must be the reference's, bit for bit: windows, the wrap-around region, long clusters (patched after their
    window left), the k-mer-0 node (inserted on the host copy), and the fall-backs."""
    if build_path != "partitioned":
        pytest.skip("partitioned path only")
    for k in ("DBG_B200_PART_SHIFT", "DBG_B200_BATCH_BASES", "DBG_B200_BATCH_READS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("DBG_B200_PART_SHIFT", "14")
    monkeypatch.setenv("DBG_B200_SUB_BASES", "300000")
    monkeypatch.setenv("DBG_B200_SUB_READS", "5000")
    monkeypatch.setenv("DBG_B200_FINISH_GROUPS", str(groups))
    from dbg_assembly_b200 import synth
    seed = 33 if init_slots < 5_000_000 else 21
    glen = 250_000 if init_slots < 5_000_000 else 300_000
    p = synth.make_params(seed=seed, genome_len=glen, read_len=150, insert=500, err=0.01, n_rate=0.002)
    hb, ho = synth.reads_host(p, 0, n)
    if init_slots >= 5_000_000:
        extra = [b"A" * 150] * 40 + [b"T" * 150] * 3      # the k-mer-0 node with saturating lanes
        eb = np.frombuffer(b"".join(extra), dtype=np.uint8)
        hb = np.concatenate([hb, eb])
        ho = np.concatenate([ho, ho[-1] + np.cumsum(np.array([len(x) for x in extra], dtype=ho.dtype))])
    n_all = len(ho) - 1
    o = oracle_build(oracle_mod, [(hb, ho)], K, 150, init_slots, load=0.9, wide=K > 31)
    with dbg.DBGBuilder(K=K, max_read_len=150, init_slots=init_slots, load_factor=0.9) as b:
        cut = n_all // 4 if earlier else 0
        if earlier:
            b.submit(hb, ho[: cut + 1])
        P = b.get_stats()["array_size"]
        pa = dbg.capi.PinnedBuffer(P * (32 if K > 31 else 16)); pn = dbg.capi.PinnedBuffer(P // 8 + 1)
        try:
            from dbg_assembly_b200.graph import NODE16, NODE32
            arr = pa.array.view(NODE32 if K > 31 else NODE16); nul = pn.array
            arr.view(np.uint8)[:] = 0xEE; nul[:] = 0xEE
            st, _, _ = b.finish_export(hb, ho[cut:], arr, nul)
            info = b.export_info()
            if expect == "grouped":
                assert info["chunks_plain"] >= 2 and info["chunks_compact"] == 0, info          # windows
            else:
                assert info["chunks_plain"] == 1, info
            assert st["count"] == o.count and st["occurrences"] == o.occurrences and st["kmers_logged"] == o.kmers_logged
            assert st["reads"] == n_all
            e = o.dump()
            d = image_to_dump(arr, nul, o.size)
            for k in ("slot", "kmer", "l", "r") + (("kmer_hi",) if K > 31 else ()):
                assert np.array_equal(d[k], e[k]), k
            # and the device image the link pass works on is the same table
            arr2, nul2 = b.export_kmerset()
            assert np.array_equal(nul2, nul) and arr2.tobytes() == arr.tobytes()
            # n_reads == 0 on a finalized context: export only
            arr.view(np.uint8)[:] = 0x11; nul[:] = 0x11
            b.finish_export(hb[:0], ho[:1], arr, nul)
            assert np.array_equal(nul2, nul) and arr2.tobytes() == arr.tobytes()
        finally:
            pa.close(); pn.close()
    o.close()


def test_full_size_properties_C2(dbg, build_path, monkeypatch):
    """BASELINE config C2 at full size (3.07 M reads, 3.68e8 occurrences): size-independent properties
    -- conservation of occurrences in the link lanes, idempotent rebuild, and the direct and the
    partitioned build paths agreeing bit for bit (layout included)."""
    import torch
    if build_path != "direct":
        pytest.skip("runs once; compares both paths itself")
    for k in ("DBG_B200_PARTITION", "DBG_B200_PART_SHIFT", "DBG_B200_BATCH_BASES", "DBG_B200_BATCH_READS"):
        monkeypatch.delenv(k, raising=False)      # library defaults: auto path selection (partitioned at this size)
    from dbg_assembly_b200 import synth
    cfg = synth.CONFIGS["C2"]
    p = synth.make_params(cfg["seed"], cfg["genome_len"], cfg["read_len"], cfg["insert"], cfg["err"], cfg["n_rate"])
    n, L, K = cfg["n_reads"], cfg["read_len"], cfg["K"]
    db = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    synth.reads_device(p, 0, n, db.data_ptr())
    do = (torch.arange(n + 1, dtype=torch.int64, device="cuda") * L)
    torch.cuda.synchronize()
    with dbg.DBGBuilder(K=K, max_read_len=L, init_g=cfg["init_g"]) as b:
        b.submit_device(db.data_ptr(), do.data_ptr(), n, 0, n * L)
        st = b.finalize()
        assert st["occurrences"] == n * (L - K + 1)
        assert st["array_size"] == 200000033
        d1 = b.dump_compact(-1)
        lk = b.export_links(2, lists=False)
        # every occurrence adds one left and one right link except at read ends: sum over lanes (unsaturated
        # nodes) == 2*occ - 2*reads; saturation only lowers it
        lanes = sum(int(((d1["l"] >> s) & 0xFF).sum()) + int(((d1["r"] >> s) & 0xFF).sum()) for s in (24, 16, 8, 0))
        sat = int(sum(((d1[x] >> s) & 0xFF == 255).sum() for x in ("l", "r") for s in (24, 16, 8, 0)))
        assert lanes <= 2 * st["occurrences"] - 2 * n
        if sat == 0:
            assert lanes == 2 * st["occurrences"] - 2 * n
        assert lk["total"] == st["count"] == len(d1["kmer"])
        assert int(lk["depth_stat"].sum()) == 8 * st["count"]
        # the image is a valid table: slots strictly increasing, keys unique
        assert (np.diff(d1["slot"].astype(np.int64)) > 0).all()
        assert len(np.unique(d1["kmer"])) == len(d1["kmer"])
        # rebuild gives the same image (deterministic, layout included)
        b.reset()
        b.submit_device(db.data_ptr(), do.data_ptr(), n, 0, n * L)
        st2 = b.finalize()
        d2 = b.dump_compact(-1)
        assert st2["count"] == st["count"]
        for k in ("slot", "kmer", "l", "r"):
            assert np.array_equal(d1[k], d2[k])
    # the direct (fused random-access) path produces the identical image
    monkeypatch.setenv("DBG_B200_PARTITION", "0")
    with dbg.DBGBuilder(K=K, max_read_len=L, init_g=cfg["init_g"]) as b:
        b.submit_device(db.data_ptr(), do.data_ptr(), n, 0, n * L)
        st3 = b.finalize()
        d3 = b.dump_compact(-1)
    assert st3["count"] == st["count"] and st3["occurrences"] == st["occurrences"]
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(d1[k], d3[k])


# ---------------------------------------------------------------------------------------------------
# the drop-in: reference front end + host traversal on the GPU-built table, byte-identical outputs
# ---------------------------------------------------------------------------------------------------
B200_CONTIG = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_b200")
OUT_SUFFIXES = (".contig.seq.fa", ".contig.small.fa", ".contig.kmer.freq", ".contig.tip.fa", ".contig.bubble.fa",
                ".contig.lowedge.fa", ".contig.seq.depth", ".contig.small.depth")


@pytest.mark.skipif(not os.access(B200_CONTIG, os.X_OK), reason="oracle/_ref/debruijn_contig_b200 not built (needs /root/reference at build time)")
def test_contig_files_byte_identical_to_reference(dbg, oracle_mod, tmp_path):
    """debruijn_contig (reference main.cpp + contig.cpp) linked against libdbgb200 through
    integration/DBGgraph_b200.cpp writes the same nine files as the reference itself (-t 1)."""
    g = load_golden("contig_k31")
    paths = []
    for i, (bases, offs) in enumerate(g["files"]):
        p = str(tmp_path / f"f{i}.fa")
        oracle_mod.write_fasta(p, bases, offs)
        paths.append(p)
    lib = str(tmp_path / "reads.lib")
    open(lib, "w").write("\n".join(paths) + "\n")
    pre = str(tmp_path / "b200")
    r = subprocess.run([B200_CONTIG, "-k", str(g["K"]), "-r", str(g["R"]), "-f", "2", "-t", "1", "-i", repr(g["init_g"]),
                        "-l", repr(g["load"]), "-M", "100", "-o", pre, lib], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    for suf in OUT_SUFFIXES:
        got = open(pre + suf, "rb").read()
        exp = g["file" + suf.replace(".", "_")].tobytes()
        assert got == exp, f"{suf} differs from the reference"
    log = r.stderr.decode()
    assert f"array_size:\t{g['size']}" in log and f"count:\t{g['count']}" in log
    # and, where the reference binary itself travelled, a fresh input through both programs
    ref = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_ref")
    if os.access(ref, os.X_OK):
        reads = random_reads(99, 5000, 100, 100, genome_len=15000, err=0.004, n_rate=0.0, lower=0.0)
        bases, offs = reads_to_arrays(reads)
        p = str(tmp_path / "fresh.fa"); oracle_mod.write_fasta(p, bases, offs)
        lib2 = str(tmp_path / "fresh.lib"); open(lib2, "w").write(p + "\n")
        outs = {}
        for tag, exe in (("ref", ref), ("b200", B200_CONTIG)):
            pre2 = str(tmp_path / ("fresh_" + tag))
            rr = subprocess.run([exe, "-k", "25", "-r", "100", "-f", "2", "-t", "1", "-i", "0.0005", "-M", "100", "-o", pre2, lib2],
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
            assert rr.returncode == 0, rr.stderr.decode()[-2000:]
            outs[tag] = {suf: open(pre2 + suf, "rb").read() for suf in OUT_SUFFIXES}
        assert outs["ref"] == outs["b200"]
        assert len(outs["ref"][".contig.seq.fa"]) > 1000


def test_contig_files_identical_when_the_reference_enlarges(dbg, oracle_mod, build_path, tmp_path):
    """-i far too small, -b 100: the reference enlarges its hash three times (DBGgraph.cpp:337-351).  The GPU front end
    rebuilds with a device table that can hold the nodes and lays the KmerSet out with the host-side growth replay
    (dbg_replay_growth): all eight output files must still be byte-identical to the reference's."""
    if build_path != "direct":
        pytest.skip("front-end binary: build path chosen by the library")
    ref = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_ref")
    if not (os.access(ref, os.X_OK) and os.access(B200_CONTIG, os.X_OK)):
        pytest.skip("oracle/_ref binaries not present")
    reads = random_reads(98, 3000, 100, 100, genome_len=15000, err=0.004, n_rate=0.0, lower=0.0)
    paths = []
    for i, part in enumerate((reads[:1800], reads[1800:])):
        bases, offs = reads_to_arrays(part)
        p = str(tmp_path / f"g{i}.fa"); oracle_mod.write_fasta(p, bases, offs); paths.append(p)
    lib = str(tmp_path / "grow.lib"); open(lib, "w").write("\n".join(paths) + "\n")
    outs, logs = {}, {}
    for tag, exe in (("ref", ref), ("b200", B200_CONTIG)):
        pre = str(tmp_path / ("grow_" + tag))
        rr = subprocess.run([exe, "-k", "25", "-r", "100", "-f", "2", "-t", "1", "-i", "0.000008", "-b", "100", "-e", "10", "-M", "100", "-o", pre, lib],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert rr.returncode == 0, rr.stderr.decode()[-2000:]
        outs[tag] = {suf: open(pre + suf, "rb").read() for suf in OUT_SUFFIXES}
        logs[tag] = rr.stderr.decode()
    assert "array_size:\t64151" in logs["ref"] and "array_size:\t64151" in logs["b200"]
    assert outs["ref"] == outs["b200"]
    assert len(outs["ref"][".contig.seq.fa"]) > 1000


def test_python_mirror_reproduces_the_enlarged_reference_table(dbg, build_path):
    """golden enlarge_k25: the reference enlarged twice (-i 4e-6, -b 20); the mirror of build_debruijn_graph returns the
    KmerSet in the reference's post-growth layout.  golden maxmem_k25 (-e 1 exhausted: the reference dropped reads,
    DBGgraph.cpp:346-350) is reproduced too: the mirror rebuilds on exactly the reads the reference used."""
    if build_path != "direct":
        pytest.skip("path chosen by the library")
    from dbg_assembly_b200.graph import build_debruijn_graph
    g = load_golden("enlarge_k25")
    ks = build_debruijn_graph(g["files"], KmerSize=g["K"], maxReadLen=g["R"], initHashSize=g["init_g"], hashLoadFactor=g["load"],
                              BufferNum=g["B"], maxDoubleHashTimes=g["max_double"])
    assert (ks.size, ks.max, ks.count) == (g["size"], g["max"], g["count"])
    slot = ks.filled_slots()
    assert np.array_equal(slot, g["slot"]) and np.array_equal(ks.array["kmer"][slot.astype(np.int64)], g["kmer"])
    assert np.array_equal(ks.array["l_link"][slot.astype(np.int64)], g["l"]) and np.array_equal(ks.array["r_link"][slot.astype(np.int64)], g["r"])
    g2 = load_golden("maxmem_k25")
    ks2 = build_debruijn_graph(g2["files"], KmerSize=g2["K"], maxReadLen=g2["R"], initHashSize=g2["init_g"], hashLoadFactor=g2["load"],
                               BufferNum=g2["B"], maxDoubleHashTimes=g2["max_double"])
    assert (ks2.size, ks2.max, ks2.count, ks2.Total_reads_num) == (g2["size"], g2["max"], g2["count"], g2["reads"])
    slot2 = ks2.filled_slots()
    assert np.array_equal(slot2, g2["slot"]) and np.array_equal(ks2.array["kmer"][slot2.astype(np.int64)], g2["kmer"])
    assert np.array_equal(ks2.array["l_link"][slot2.astype(np.int64)], g2["l"]) and np.array_equal(ks2.array["r_link"][slot2.astype(np.int64)], g2["r"])


@pytest.mark.skipif(not os.access(B200_CONTIG, os.X_OK), reason="oracle/_ref/debruijn_contig_b200 not built (needs /root/reference at build time)")
def test_full_size_C2_contig_files_byte_identical_to_reference(dbg, build_path, tmp_path):
    """the configuration the metric is quoted on, at FULL size, through files: the reference program (-t 1: its
    deterministic slot layout) and the front end relinked to libdbgb200 on the same 460 MB FASTA (C2: 4.6 Mb genome,
    3 066 666 x 150 bp reads, 1 % errors, K=31, -i 0.2) -- all eight output files byte for byte.  About a minute, most
    of it the reference's CPU build."""
    if build_path != "direct":
        pytest.skip("front-end binaries: the build path is chosen by the library")
    from dbg_assembly_b200 import synth
    ref = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_ref")
    gen = os.path.join(REPO, "oracle", "_bin", "synth_fasta")
    if not (os.access(ref, os.X_OK) and os.access(gen, os.X_OK)):
        pytest.skip("oracle/_ref/debruijn_contig_ref or oracle/_bin/synth_fasta not present")
    c = synth.CONFIGS["C2"]
    p = synth.make_params(c["seed"], c["genome_len"], c["read_len"], c["insert"], c["err"], c["n_rate"])
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize > (2 << 30) else str(tmp_path)
    fa = os.path.join(d, f"c2_full_{os.getpid()}.fa")
    try:
        subprocess.run([gen, str(p.seed), str(p.genome_len), str(p.read_len), str(p.insert), str(p.err_per_2p24), str(p.n_per_2p24), "0",
                        str(c["n_reads"]), fa], check=True, timeout=600)
        lib = str(tmp_path / "c2.lib"); open(lib, "w").write(fa + "\n")
        import json
        import time
        outs, wall = {}, {}
        for tag, exe in (("b200", B200_CONTIG), ("ref", ref)):
            pre = str(tmp_path / tag)
            t0 = time.perf_counter()
            rr = subprocess.run([exe, "-k", "31", "-r", "150", "-f", "2", "-t", "1", "-i", "0.2", "-M", "100", "-o", pre, lib],
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1500)
            wall[tag] = round(time.perf_counter() - t0, 2)
            assert rr.returncode == 0, rr.stderr.decode()[-2000:]
            outs[tag] = {suf: open(pre + suf, "rb").read() for suf in OUT_SUFFIXES}
            if tag == "b200":
                assert "count:\t100611066" in rr.stderr.decode()
                wall["b200_phases"] = [l for l in rr.stderr.decode().splitlines() if l.startswith("libdbgb200 wall clock")]
        out_dir = os.path.join(REPO, "gpurun_out")        # whole-program wall clocks, for profiles/ (scratch dir of the GPU runs)
        if os.path.isdir(out_dir):
            json.dump(wall, open(os.path.join(out_dir, "dropin_full_C2_wall.json"), "w"))
        for suf in OUT_SUFFIXES:
            assert outs["ref"][suf] == outs["b200"][suf], suf
        assert len(outs["ref"][".contig.seq.fa"]) > 4_000_000
    finally:
        if os.path.exists(fa):
            os.unlink(fa)
