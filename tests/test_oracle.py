"""CPU: pin the oracle (oracle/dbg_oracle.c) against the reference's own outputs.

Golden fixtures in tests/golden/*.npz were produced by running the reference itself
(tests/golden/make_golden.py -> oracle/_ref/ref_build_driver, -t 1).  They pin node contents, slot
layout, table size / max / count, and -- for the oracle only -- the enlarge and "memory reached the
maximum" paths (DBGgraph.cpp:337-351, kmerSet.cpp:132-189).
"""
import numpy as np
import pytest

from conftest import GOLDEN_ALL, load_golden, random_reads, reads_to_arrays


def build_oracle(orc, g, wide=False):
    o = orc.OracleGraph(g["K"], g["R"], g["init_slots"], g["load"], g["max_double"], g["B"], wide=wide)
    n = 0
    for bases, offs in g["files"]:
        n += o.add_file(bases, offs)
    o.finish()
    return o, n


@pytest.mark.parametrize("name", GOLDEN_ALL)
@pytest.mark.parametrize("wide", [False, True])
def test_oracle_matches_reference_golden(oracle_mod, name, wide):
    g = load_golden(name)
    o, n_reads = build_oracle(oracle_mod, g, wide=wide)
    assert o.size == g["size"]
    assert o.count == g["count"]
    assert o.max == g["max"]
    assert o.doublings == g["doublings"]
    assert n_reads == g["reads"]
    assert o.kmers_logged == g["kmers_logged"]
    assert o.conflict == g["conflict"]          # same probing scheme, same sequential order
    d = o.dump()
    assert np.array_equal(d["slot"], g["slot"])  # slot layout
    assert np.array_equal(d["kmer"], g["kmer"])
    assert np.array_equal(d["l"], g["l"])
    assert np.array_equal(d["r"], g["r"])
    if wide:
        assert not d["kmer_hi"].any()
    # array/nul_flag image agrees with the dump
    arr, nul = o.array(), o.nul_flag()
    bits = np.unpackbits(nul)[: o.size]
    assert np.array_equal(np.nonzero(bits)[0].astype(np.uint64), g["slot"])
    assert np.array_equal(arr["kmer"][g["slot"].astype(np.int64)], g["kmer"])
    o.close()


def test_scalar_known_answers(oracle_mod):
    L = oracle_mod.lib()
    # generated from the reference during the survey
    assert L.orc_hash_code(0) == 7654268697807496793
    assert L.orc_hash_code(1) == 2320827452992767577
    for n, p in ((10**7, 10000019), (10**8, 100000007), (10**9, 1000000007), (200000014, 200000033), (1000, 1009)):
        assert L.orc_find_next_prime(n) == p
    assert L.orc_hash_code_wide(12345, 0) == L.orc_hash_code(12345)
    assert L.orc_hash_code_wide(12345, 1) != L.orc_hash_code(12345)
    for c, v in zip(b"ACGTNacgtn", (0, 1, 2, 3, 0, 0, 1, 2, 3, 0)):
        assert L.orc_base_code(c) == v
    assert L.orc_base_code(ord("X")) == 4
    assert L.orc_seq2bit(b"ACGTT", 5) == 0b0001101111
    assert L.orc_rev_com_kbit(L.orc_seq2bit(b"ACGTT", 5), 5) == L.orc_seq2bit(b"AACGT", 5)


def test_parse_read_matches_definition(oracle_mod):
    """orc64 vs orc128 vs a direct python statement of Appendix A items 3-6"""
    comp = {0: 3, 1: 2, 2: 1, 3: 0}
    code = {ord(c): v for c, v in zip("ACGTNacgtn", (0, 1, 2, 3, 0, 0, 1, 2, 3, 0))}
    for seed, K in ((1, 5), (2, 16), (3, 31), (4, 21)):
        for read in random_reads(seed, 20, 1, 90):
            R = 70
            lo, hi, lb, rb = oracle_mod.parse_read(read, K, R, wide=False)
            lo2, hi2, lb2, rb2 = oracle_mod.parse_read(read, K, R, wide=True)
            assert np.array_equal(lo, lo2) and not hi2.any() and np.array_equal(lb, lb2) and np.array_equal(rb, rb2)
            n = min(len(read), R)
            exp = []
            for j in range(0, n - K + 1):
                f = 0
                for c in read[j:j + K]:
                    f = (f << 2) | code[c]
                r = 0
                for c in reversed(read[j:j + K]):
                    r = (r << 2) | comp[code[c]]
                left = code[read[j - 1]] if j > 0 else 4
                right = code[read[j + K]] if j < n - K else 4
                if f <= r:
                    exp.append((f, left, right))
                else:
                    exp.append((r, 3 - right if right != 4 else 4, 3 - left if left != 4 else 4))
            got = list(zip(lo.tolist(), lb.tolist(), rb.tolist()))
            assert got == exp


def test_wide_parse_k63_against_python(oracle_mod):
    code = {ord(c): v for c, v in zip("ACGTNacgtn", (0, 1, 2, 3, 0, 0, 1, 2, 3, 0))}
    for K in (32, 33, 47, 63):
        for read in random_reads(K, 10, K - 2, 120, n_rate=0.02):
            lo, hi, lb, rb = oracle_mod.parse_read(read, K, 100, wide=True)
            n = min(len(read), 100)
            exp = []
            for j in range(0, n - K + 1):
                f = 0
                for c in read[j:j + K]:
                    f = (f << 2) | code[c]
                r = 0
                for c in reversed(read[j:j + K]):
                    r = (r << 2) | (3 - code[c])
                left = code[read[j - 1]] if j > 0 else 4
                right = code[read[j + K]] if j < n - K else 4
                if f <= r:
                    exp.append((f, left, right))
                else:
                    exp.append((r, 3 - right if right != 4 else 4, 3 - left if left != 4 else 4))
            got = [(int(a) | (int(b) << 64), int(c), int(d)) for a, b, c, d in zip(lo, hi, lb, rb)]
            assert got == exp


def test_oracle_links_pass_small(oracle_mod):
    """calculate_kmer_links restatement: consistency with a direct numpy evaluation of contig.cpp:119-181"""
    g = load_golden("contig_k31")
    o, _ = build_oracle(oracle_mod, g)
    lk = o.kmer_links(2)
    lanes = np.stack([(g["l"] >> s) & 0xFF for s in (24, 16, 8, 0)] + [(g["r"] >> s) & 0xFF for s in (24, 16, 8, 0)], axis=1)
    hist = np.bincount(lanes.ravel(), minlength=256)
    assert np.array_equal(hist, lk["depth_stat"])
    ln = np.minimum((lanes[:, :4] > 2).sum(1), 3)
    rn = np.minimum((lanes[:, 4:] > 2).sum(1), 3)
    assert lk["total"] == g["count"]
    assert lk["deleted"] == int(((ln == 0) & (rn == 0)).sum())
    assert lk["linear"] == int(((ln == 1) & (rn == 1)).sum())
    assert np.array_equal(lk["tips"], g["slot"][(ln + rn) == 1])
    assert np.array_equal(lk["branches"], g["slot"][(ln > 1) | (rn > 1)])
    # the reference's own .contig.kmer.freq file (rows 1..255) equals the histogram
    txt = g["file_contig_kmer_freq"].tobytes().decode().splitlines()
    assert txt[0] == "Kmer_depth\tAppear_times"
    rows = [tuple(map(int, line.split("\t"))) for line in txt[1:]]
    assert rows == [(d, int(lk["depth_stat"][d])) for d in range(1, 256)]
    o.close()


@pytest.mark.skipif(not __import__("oracle.oracle", fromlist=["x"]).have_reference(), reason="oracle/_ref not built here")
def test_oracle_matches_live_reference(oracle_mod, tmp_path):
    """fresh random input through the compiled reference (this container / any box the binaries reached)"""
    reads = random_reads(77, 300, 10, 140, genome_len=5000)
    bases, offs = reads_to_arrays(reads)
    p = str(tmp_path / "r.fa")
    oracle_mod.write_fasta(p, bases, offs)
    init_g = 3e-5
    stats, d = oracle_mod.run_ref_build([p], 27, 120, init_g, threads=1)
    o = oracle_mod.OracleGraph(27, 120, int(init_g * 1e9), 0.7)
    o.add_file(bases, offs)
    o.finish()
    dd = o.dump()
    assert o.size == d["size"] and o.count == d["count"] and o.conflict == stats["conflict"]
    for k in ("slot", "kmer", "l", "r"):
        assert np.array_equal(dd[k], d[k])
    # threads only change the cluster-internal order, never node contents
    stats4, d4 = oracle_mod.run_ref_build([p], 27, 120, init_g, threads=4)
    a = np.lexsort((d["r"], d["l"], d["kmer"])); b = np.lexsort((d4["r"], d4["l"], d4["kmer"]))
    for k in ("kmer", "l", "r"):
        assert np.array_equal(d[k][a], d4[k][b])
    o.close()
