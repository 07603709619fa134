"""K-mer frequency table for correct_error.  PARITY UNPINNED: the producer `kmerfreq` is
external to the reference; the contract is what correct_error's loaders read (restated in oracle/oracle.py:
load_cz_1bit <- main_parallel_senior.cpp:334-408, load_cz_8bit <- main.cpp:161-220) and the .stat artefacts."""
import os
import zlib

import numpy as np
import pytest

from conftest import REPO, random_reads, reads_to_arrays


def both_strand_bits(counts, K, cutoff, orc):
    L = orc.lib()
    bits = np.zeros(len(counts), dtype=np.uint8)
    hi = np.nonzero(counts > cutoff)[0]
    bits[hi] = 1
    for i in hi.tolist():
        bits[L.orc_rev_com_kbit(int(i), K)] = 1
    return bits


def test_cz_loader_restatement_roundtrip(oracle_mod, tmp_path):
    """CPU: a .cz written from oracle counts with plain zlib and read back through the restated loaders gives the
    both-strand high-frequency table (pins the test's own reader of the format)"""
    K = 9
    reads = random_reads(5, 400, 5, 80, genome_len=600, err=0.01)
    bases, offs = reads_to_arrays(reads)
    counts = oracle_mod.kfreq_count(bases, offs, K)
    assert counts.sum() == sum(max(0, len(r) - K + 1) for r in reads)
    # canonical: an entry and its reverse complement never both counted unless equal
    L = oracle_mod.lib()
    nz = np.nonzero(counts)[0]
    assert all(i <= L.orc_rev_com_kbit(int(i), K) for i in nz.tolist())
    prefix = str(tmp_path / "t")
    cutoff = 2
    raw = np.packbits((counts > cutoff).astype(np.uint8)).tobytes()
    comp = zlib.compress(raw)
    open(prefix + ".kmer.freq.cz", "wb").write(comp)
    open(prefix + ".kmer.freq.cz.len", "w").write(f"{len(comp)}\n")
    both, canon = oracle_mod.load_cz_1bit(prefix, K)
    assert np.array_equal(canon, (counts > cutoff).astype(np.uint8))
    assert np.array_equal(both, both_strand_bits(counts, K, cutoff, oracle_mod))


@pytest.mark.gpu
@pytest.mark.parametrize("K", [5, 9, 13])
def test_kfreq_counts_match_oracle(oracle_mod, tmp_path, K):
    from dbg_assembly_b200.kfreq import KmerFreq
    reads = random_reads(100 + K, 3000, 1, 160, genome_len=20000, err=0.01, n_rate=0.01) + [b"A" * 100, b"", b"ACGT" * 30] + [b"GGGGGGGGGGGGGGGGGGGGGGGG"] * 400
    bases, offs = reads_to_arrays(reads)
    counts = oracle_mod.kfreq_count(bases, offs, K)
    with KmerFreq(K=K) as kf:
        half = len(reads) // 2
        kf.submit(bases, offs[: half + 1])
        kf.submit(bases, offs[half:])
        st = kf.finalize()
        assert st["occurrences"] == int(counts.sum()) and st["reads"] == len(reads)
        assert np.array_equal(kf.export(bits=8), np.minimum(counts, 255).astype(np.uint8))
        for cutoff in (0, 1, 10):
            got = np.unpackbits(kf.export(bits=1, cutoff=cutoff))[: len(counts)]
            assert np.array_equal(got, (counts > cutoff).astype(np.uint8)), cutoff
        h = kf.histogram()
        exp = np.bincount(np.minimum(counts[counts > 0], 65535), minlength=65536)
        assert np.array_equal(h[1:], exp[1:].astype(np.uint64))
        # files, read back the way correct_error reads them
        prefix = str(tmp_path / f"k{K}")
        kf.write_cz(prefix, bits=1, cutoff=1)
        both, canon = oracle_mod.load_cz_1bit(prefix, K)
        assert np.array_equal(canon, (counts > 1).astype(np.uint8))
        assert np.array_equal(both, both_strand_bits(counts, K, 1, oracle_mod))
        n_blocks = max(1, (4 ** K + (8 << 20) - 1) // (8 << 20))
        assert len(open(prefix + ".kmer.freq.cz.len").read().split()) == n_blocks
        # spectrum file: same layout as test/01.clean_correct/*.kmer.freq.stat
        lines = open(prefix + ".kmer.freq.stat").read().split("\n")
        assert lines[0] == f"#Kmer size: {K}" and lines[1] == "#Maximum Kmer frequency: 65535"
        assert lines[2] == f"#Kmer indivdual number: {int(counts.sum())}"
        assert lines[3] == f"#Kmer species number: {int((counts > 0).sum())}"
        assert lines[5] == "" and lines[6].startswith("#Kmer_Frequency\tKmer_Species_Number")
        rows = [l.split("\t") for l in lines[7:] if l]
        assert len(rows) == 65535 and [int(r[0]) for r in rows[:3]] == [1, 2, 3]
        assert [int(r[1]) for r in rows[:300]] == exp[1:301].tolist()
        prefix8 = str(tmp_path / f"k{K}_8")
        kf.write_cz(prefix8, bits=8, cutoff=0)
        bits8, vals = oracle_mod.load_cz_8bit(prefix8, K, 3)
        assert np.array_equal(vals, np.minimum(counts, 255).astype(np.uint8))
        assert np.array_equal(bits8, both_strand_bits(counts, K, 3, oracle_mod))


@pytest.mark.gpu
def test_kfreq_k17_shape_and_sharding(oracle_mod, tmp_path):
    """K=17 (the configuration correct_error is run with, test/01.clean_correct/work.sh:19): 2048 blocks like the
    reference's .cz.len artefact; two block-range shards together equal the single table"""
    from dbg_assembly_b200.kfreq import KmerFreq
    from dbg_assembly_b200 import synth
    K = 17
    p = synth.make_params(seed=4, genome_len=200_000, read_len=150, insert=400, err=0.01, n_rate=0.0)
    bases, offs = synth.reads_host(p, 0, 20_000)
    # sparse oracle: canonical 17-mers via the parse restatement (a dense table would be 68 GB on the host)
    L = oracle_mod.lib()
    import ctypes as C
    lo = np.zeros(150, np.uint64); lb = np.zeros(150, np.uint8); rb = np.zeros(150, np.uint8)
    allk = []
    u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    for i in range(2000):
        rd = bases[int(offs[i]):int(offs[i + 1])].tobytes()
        m = L.orc64_parse_read(rd, len(rd), K, 65535, lo.ctypes.data_as(u64p), lb.ctypes.data_as(u8p), rb.ctypes.data_as(u8p))
        allk.append(lo[:m].copy())
    keys, cnt = np.unique(np.concatenate(allk), return_counts=True)
    with KmerFreq(K=K) as kf:
        kf.submit(bases, offs[:2001])
        st = kf.finalize()
        assert st["occurrences"] == 2000 * 134
        bits = kf.export(bits=1, cutoff=0)
        assert len(bits) == 4 ** 17 // 8
        got = np.nonzero(np.unpackbits(bits[: (int(keys.max()) >> 3) + 1]))[0]
        assert np.array_equal(got.astype(np.uint64), keys)
        assert int(np.unpackbits(bits).sum()) == len(keys) if len(bits) < (1 << 28) else True
        h = kf.histogram()
        assert np.array_equal(h[1:20], np.bincount(cnt, minlength=20)[1:20].astype(np.uint64))
        prefix = str(tmp_path / "k17")
        kf.write_cz(prefix, bits=1, cutoff=0)
        lens = open(prefix + ".kmer.freq.cz.len").read().split()
        assert len(lens) == 2048                      # == test/01.clean_correct/clean_reads.lib.kmer.freq.cz.len
        assert os.path.getsize(prefix + ".kmer.freq.cz") == sum(int(x) for x in lens)
    # two shards (contiguous runs of whole blocks)
    parts = []
    for r in range(2):
        with KmerFreq(K=K, block_rank=r, block_count=2) as kf:
            kf.submit(bases, offs[:2001])
            kf.finalize()
            a, b = kf.index_range()
            assert (a, b) == (r * 4 ** 17 // 2, (r + 1) * 4 ** 17 // 2)
            parts.append(kf.export(bits=1, cutoff=0))
    assert np.array_equal(np.concatenate(parts), bits)
