"""Contig seed index of link_scaffold: map_pair / map_reads' k-mer -> {contig id, position, unique?, strand}
hash and get_align_seed.

CPU (not gpu): the oracle (oracle/seed_oracle.c) against golden fixtures produced by the reference's own code
(tests/golden/make_seed_golden.py -> oracle/_ref/ref_seed_driver) and, where the reference is present, against a live run.
GPU: the product (seedidx_* C ABI through dbg_assembly_b200.SeedIndex) against the oracle and the same fixtures: table size,
count, every node AND its slot, and the seeds -- bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

SEED_GOLDEN = ["seed_mixed_k31", "seed_polyT_k21", "seed_long_k27"]
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
RC = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def load_seed_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("K", "min_ctg_len", "seed_kmer_num", "size", "count", "max", "conflict"):
        d[k] = int(d[k])
    cb, co, rb, ro = d["contig_bases"].tobytes(), d["contig_offs"], d["read_bases"].tobytes(), d["read_offs"]
    d["contigs"] = [cb[int(co[i]):int(co[i + 1])] for i in range(len(co) - 1)]
    d["reads"] = [rb[int(ro[i]):int(ro[i + 1])] for i in range(len(ro) - 1)]
    return d


def kept_contigs(contigs, min_ctg_len):
    return [c if len(c) >= min_ctg_len else b"" for c in contigs]          # map_pair.cpp:100-110


def oracle_index(orc, contigs, K, min_ctg_len):
    kept = kept_contigs(contigs, min_ctg_len)
    o = orc.SeedOracle(K, 3 * sum(len(c) for c in kept), 0.5)
    assert o.add_contigs(kept) == 0
    return o


def filled(arr, nul, size):
    occ = np.unpackbits(nul)[:size].astype(bool)
    return np.nonzero(occ)[0].astype(np.uint64), arr["kmer"][occ], arr["value"][occ]


def rnd(rng, n):
    return bytes(rng.choice(ACGT, n))


def rc(s):
    return s[::-1].translate(RC)


def random_case(seed, n_contigs=12, K=31):
    rng = np.random.default_rng(seed)
    g = rnd(rng, 40_000)
    contigs = []
    for i in range(n_contigs):
        a = int(rng.integers(0, len(g) - 3000)); ln = int(rng.integers(K + 5, 3000))
        c = g[a:a + ln]
        if rng.random() < 0.3:
            c = rc(c)
        if rng.random() < 0.3:
            p = int(rng.integers(K, max(K + 1, len(c) - K)))
            c = c[:p] + b"N" * int(rng.integers(1, 20)) + c[p:] + rnd(rng, K + 3)
        contigs.append(c)
    # blocks shorter than K crash the reference (undefined loop bound): make every N-free block >= K
    fixed = []
    for c in contigs:
        parts = [b for b in c.split(b"N") if len(b) >= K]
        fixed.append(b"NNN".join(parts))
    reads = []
    for _ in range(150):
        c = fixed[int(rng.integers(len(fixed)))]
        L = int(rng.integers(K + 5, 200))
        if len(c) <= L:
            continue
        p = int(rng.integers(0, len(c) - L))
        r = bytearray(c[p:p + L])
        if rng.random() < 0.5:
            r[int(rng.integers(L))] = b"ACGT"[int(rng.integers(4))]
        r = bytes(r)
        reads.append(rc(r) if rng.random() < 0.5 else r)
    return fixed, reads


# ---------------------------------------------------------------------------------------------------
# CPU: the oracle is pinned against the reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SEED_GOLDEN)
def test_seed_oracle_matches_golden_reference_tables(oracle_mod, name):
    g = load_seed_golden(name)
    o = oracle_index(oracle_mod, g["contigs"], g["K"], g["min_ctg_len"])
    assert (o.size, o.count, o.max, o.conflict) == (g["size"], g["count"], g["max"], g["conflict"])
    slot, kmer, value = filled(o.array(), o.nul_flag(), o.size)
    assert np.array_equal(slot, g["slot"]) and np.array_equal(kmer, g["kmer"]) and np.array_equal(value, g["value"])
    assert np.array_equal(o.align(g["reads"], g["seed_kmer_num"]), g["seeds"])
    o.close()


@pytest.mark.parametrize("seed,K", [(1, 31), (2, 17), (3, 24)])
def test_seed_oracle_matches_live_reference(oracle_mod, seed, K):
    if not oracle_mod.have_seed_reference():
        pytest.skip("oracle/_ref/ref_seed_driver not built (no /root/reference here)")
    contigs, reads = random_case(seed, K=K)
    names = [b"c%d" % i for i in range(len(contigs))]
    ref = oracle_mod.run_ref_seed(names, contigs, reads, K, 60, 4)
    o = oracle_index(oracle_mod, contigs, K, 60)
    assert (o.size, o.count, o.max, o.conflict) == (ref["size"], ref["count"], ref["max"], ref["conflict"])
    assert np.array_equal(o.array(), ref["array"]) and np.array_equal(o.nul_flag(), ref["nul"])
    assert np.array_equal(o.align(reads, 4), ref["seeds"])
    o.close()


def test_seed_oracle_reports_a_table_the_reference_would_enlarge(oracle_mod):
    rng = np.random.default_rng(9)
    o = oracle_mod.SeedOracle(21, 300, 0.5)
    assert o.add_contigs([rnd(rng, 1000)]) == -1
    o.close()


# ---------------------------------------------------------------------------------------------------
# GPU: the product against the oracle / the reference's tables
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dbg():
    import dbg_assembly_b200 as d
    if d.capi.device_count() == 0:
        pytest.skip("no CUDA device")
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("name", SEED_GOLDEN)
def test_gpu_seed_index_is_the_reference_table(dbg, name):
    g = load_seed_golden(name)
    idx = dbg.SeedIndex.from_contigs(g["contigs"], K=g["K"], min_ctg_len=g["min_ctg_len"])
    assert (idx.size, idx.count, idx.max) == (g["size"], g["count"], g["max"])
    arr, nul = idx.export()
    slot, kmer, value = filled(arr, nul, idx.size)
    assert np.array_equal(slot, g["slot"]), "slot layout differs from the reference"
    assert np.array_equal(kmer, g["kmer"]) and np.array_equal(value, g["value"])
    assert np.array_equal(idx.align(g["reads"], seed_kmer_num=g["seed_kmer_num"]), g["seeds"])
    assert idx.launches > 0
    idx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,K", [(4, 31), (5, 15), (6, 22), (7, 31)])
def test_gpu_seed_index_matches_oracle_on_random_contigs(dbg, oracle_mod, seed, K):
    contigs, reads = random_case(seed, n_contigs=40, K=K)
    o = oracle_index(oracle_mod, contigs, K, 60)
    kept = kept_contigs(contigs, 60)
    with dbg.SeedIndex(K=K, init_slots=3 * sum(len(c) for c in kept)) as idx:
        half = len(kept) // 2
        idx.add_contigs(kept[:half])                     # ids continue across calls
        idx.add_contigs(kept[half:])
        st = idx.finalize()
        assert (st["size"], st["count"], st["max"]) == (o.size, o.count, o.max)
        arr, nul = idx.export()
        assert np.array_equal(nul, o.nul_flag())
        occ = np.unpackbits(nul)[:o.size].astype(bool)
        assert np.array_equal(arr[occ], o.array()[occ])
        assert np.array_equal(idx.align(reads, seed_kmer_num=4), o.align(reads, 4))
    o.close()


@pytest.mark.gpu
def test_gpu_seed_index_medium_genome_through_the_partitioned_path(dbg, oracle_mod, monkeypatch):
    """2 Mb of contigs (forced through the radix-partitioned build), 20 000 reads: same table, same seeds"""
    monkeypatch.setenv("DBG_B200_PARTITION", "1")
    rng = np.random.default_rng(21)
    g = rnd(rng, 2_000_000)
    cuts = np.sort(rng.integers(0, len(g), 60))
    contigs = [g[int(a):int(b)] for a, b in zip(np.r_[0, cuts], np.r_[cuts, len(g)]) if b - a >= 200]
    contigs += [rc(contigs[3][:5000]), contigs[5][1000:9000]]
    reads = []
    for _ in range(20_000):
        c = contigs[int(rng.integers(len(contigs)))]
        p = int(rng.integers(0, len(c) - 150))
        r = c[p:p + 150]
        reads.append(rc(r) if rng.random() < 0.5 else r)
    o = oracle_index(oracle_mod, contigs, 31, 125)
    idx = dbg.SeedIndex.from_contigs(contigs, K=31, min_ctg_len=125)
    assert (idx.size, idx.count) == (o.size, o.count)
    arr, nul = idx.export()
    assert np.array_equal(nul, o.nul_flag())
    occ = np.unpackbits(nul)[:o.size].astype(bool)
    assert np.array_equal(arr[occ], o.array()[occ])
    assert np.array_equal(idx.align(reads), o.align(reads, 5))
    idx.close(); o.close()


@pytest.mark.gpu
def test_gpu_seed_lookup_from_a_later_read_position(dbg, oracle_mod):
    """map_reads searches a second seed behind the first alignment (map_reads.cpp:484): search_start per read"""
    contigs, reads = random_case(8, n_contigs=30, K=31)
    o = oracle_index(oracle_mod, contigs, 31, 60)
    idx = dbg.SeedIndex.from_contigs(contigs, K=31, min_ctg_len=60)
    rng = np.random.default_rng(3)
    starts = np.array([int(rng.integers(1, max(2, len(r) - 40))) for r in reads], dtype=np.int32)
    want = np.full((len(reads), 6), -1, dtype=np.int32); want[:, 5] = ord("N")
    rec = np.zeros(6, dtype=np.int32)
    for i, r in enumerate(reads):
        o.L.orc_seed_align(o.h, r, len(r), int(starts[i]), len(r), 5, rec.ctypes.data)
        want[i] = rec
    assert np.array_equal(idx.align(reads, search_start=starts), want)
    idx.close(); o.close()


@pytest.mark.gpu
def test_gpu_seed_index_refuses_a_table_the_reference_would_enlarge(dbg):
    rng = np.random.default_rng(9)
    with dbg.SeedIndex(K=21, init_slots=300, load_factor=0.5) as idx:
        idx.add_contigs([rnd(rng, 200)])                 # 180 k-mers >= max = 153 of 307 slots
        with pytest.raises(dbg.capi.DbgError) as e:
            idx.finalize()
        assert e.value.code == dbg.capi.DBG_ERR_TABLE_FULL


@pytest.mark.gpu
def test_gpu_seed_index_of_nothing(dbg):
    with dbg.SeedIndex(K=31, init_slots=1000) as idx:
        idx.add_contigs([b"", b"ACGT"])
        st = idx.finalize()
        assert st["count"] == 0
        arr, nul = idx.export()
        assert not nul.any()
        assert np.array_equal(idx.align([b"ACGT" * 20])[0], [-1, -1, -1, -1, -1, ord("N")])
