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


def _sim_reads(seed, genome_len, n_reads, L=100, err=0.01):
    rng = np.random.default_rng(seed)
    genome = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=genome_len).tolist())
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for _ in range(n_reads):
        p = int(rng.integers(0, genome_len - L))
        r = bytearray(genome[p:p + L])
        for j in np.nonzero(rng.random(L) < err)[0].tolist():
            r[j] = b"ACGT"[(b"ACGT".index(r[j]) + int(rng.integers(1, 4))) % 4]
        r = bytes(r)
        reads.append(r.translate(comp)[::-1] if rng.random() < 0.5 else r)
    return reads


def _write_cz_1bit(prefix, counts, cutoff):
    """the format correct_error loads (main_parallel_senior.cpp:334-408): zlib streams of 1 MiB of bits (8 Mi k-mers)"""
    bits = np.packbits((counts > cutoff).astype(np.uint8))
    lens = []
    with open(prefix + ".kmer.freq.cz", "wb") as f:
        for i in range(0, len(bits), 1 << 20):
            c = zlib.compress(bits[i:i + (1 << 20)].tobytes())
            f.write(c); lens.append(len(c))
    with open(prefix + ".kmer.freq.cz.len", "w") as f:
        f.write("".join(f"{n}\n" for n in lens))


def _run_correct_error(orc, workdir, prefix, reads, K, tag):
    """the reference's own consumer (shipped binary, oracle/_ref/correct_error_reads_elf): loads <prefix>.kmer.freq.cz
    and corrects the reads; returns (number of high-frequency k-mers it found in the table, corrected FASTA bytes)"""
    import gzip
    import subprocess
    fa = os.path.join(workdir, f"r_{tag}.fa.gz")
    with gzip.open(fa, "wb", compresslevel=1) as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n%s\n" % (i, r))
    lib = os.path.join(workdir, f"r_{tag}.lib")
    with open(lib, "w") as f:
        f.write(fa + "\n")
    p = subprocess.run([orc.CORRECT_ELF, "-k", str(K), "-c", "2", "-f", "2", "-t", "1", "-r", "50", prefix + ".kmer.freq.cz", lib],
                       cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    log = p.stdout.decode()
    hifreq = int(log.split("Kmer_hifreq_num")[1].split()[0])
    return hifreq, gzip.open(fa + ".correct.fa.gz").read()


def test_cz_format_is_what_the_reference_consumer_loads(oracle_mod, tmp_path):
    """CPU pin of the table format against the REAL loader: the shipped correct_error_reads reads a .cz/.cz.len written
    by this test's plain-zlib writer, reports exactly the number of set canonical entries, and corrects reads with it
    (most erroneous reads come back as their error-free originals' length, almost nothing is deleted)"""
    if not os.path.exists(oracle_mod.CORRECT_ELF):
        pytest.skip("oracle/_ref/correct_error_reads_elf not present (make -C oracle ref)")
    K, cutoff = 13, 3
    reads = _sim_reads(3, 30_000, 12_000)
    bases, offs = reads_to_arrays(reads)
    counts = oracle_mod.kfreq_count(bases, offs, K)
    prefix = str(tmp_path / "tab")
    _write_cz_1bit(prefix, counts, cutoff)
    both, canon = oracle_mod.load_cz_1bit(prefix, K)
    hifreq, corrected = _run_correct_error(oracle_mod, str(tmp_path), prefix, reads, K, "cpu")
    assert hifreq == int(canon.sum()) == int((counts > cutoff).sum())
    heads = [l for l in corrected.split(b"\n") if l.startswith(b">")]
    assert len(heads) > 0.95 * len(reads)                      # few reads deleted: the table marks the genome's k-mers
    modified = sum(1 for h in heads if b"ModifiedBaseNum: 0" not in h)
    assert modified > 0.3 * len(reads)                         # ~63 % of the reads carry an error; most are corrected
    # a table with the bit order reversed (LSB first) is a different table for the consumer
    bad = str(tmp_path / "bad")
    bits = np.packbits((counts > cutoff).astype(np.uint8), bitorder="little")
    c = [zlib.compress(bits[i:i + (1 << 20)].tobytes()) for i in range(0, len(bits), 1 << 20)]
    open(bad + ".kmer.freq.cz", "wb").write(b"".join(c))
    open(bad + ".kmer.freq.cz.len", "w").write("".join(f"{len(x)}\n" for x in c))
    _, corrected_bad = _run_correct_error(oracle_mod, str(tmp_path), bad, reads, K, "bad")
    assert corrected_bad != corrected


def _genome_fasta(path, seed, n_chr, lo, hi, wrap=60):
    """multi-line FASTA, a few chromosomes (ACGT only: construct_ref_kmer_table has no N handling, seqKmer.cpp:34-41)"""
    rng = np.random.default_rng(seed)
    chrs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(lo, hi)))) for _ in range(n_chr)]
    with open(path, "wb") as f:
        for i, c in enumerate(chrs):
            f.write(b">chr%d some text\n" % i)
            for j in range(0, len(c), wrap):
                f.write(c[j:j + wrap] + b"\n")
    return chrs


@pytest.mark.parametrize("K", [5, 9, 12])
def test_oracle_bit_table_equals_the_reference_table_builder(oracle_mod, tmp_path, K):
    """SURVEY 8 a-15, CPU: the reference's own in-tree builder (construct_ref_kmer_table,
    correct_error/simulate_lowfreq_kmer.cpp:189-260, compiled in place) on a genome FASTA vs the oracle's canonical
    counter + the loader's reverse-complement rule: same bits, MSB-first layout included"""
    import os
    if not os.access(oracle_mod.KMER_TABLE_DRIVER, os.X_OK):
        pytest.skip("oracle/_ref/ref_kmer_table_driver not built (needs /root/reference)")
    fa = str(tmp_path / "genome.fa")
    chrs = _genome_fasta(fa, 40 + K, 4, 300, 5000)
    ref_bits = oracle_mod.ref_kmer_table(fa, K)
    bases, offs = reads_to_arrays(chrs)
    counts = oracle_mod.kfreq_count(bases, offs, K)
    assert np.array_equal(both_strand_bits(counts, K, 0, oracle_mod), ref_bits)
    assert int(ref_bits.sum()) > 100


@pytest.mark.gpu
@pytest.mark.parametrize("K", [9, 13])
def test_gpu_bit_table_equals_the_reference_table_builder(oracle_mod, tmp_path, K):
    """SURVEY 8 a-15, GPU: kfreq_export(bits=1, cutoff 0) of the chromosomes + the loader's reverse-complement OR ==
    the table construct_ref_kmer_table builds from the same genome FASTA (the reference's own code, compiled in place)"""
    import os
    from dbg_assembly_b200.kfreq import KmerFreq
    if not os.access(oracle_mod.KMER_TABLE_DRIVER, os.X_OK):
        pytest.skip("oracle/_ref/ref_kmer_table_driver not built (needs /root/reference)")
    fa = str(tmp_path / "genome.fa")
    chrs = _genome_fasta(fa, 70 + K, 5, 2000, 60000)
    ref_bits = oracle_mod.ref_kmer_table(fa, K)
    bases, offs = reads_to_arrays(chrs)
    with KmerFreq(K=K) as kf:
        kf.submit(bases, offs)
        kf.finalize()
        canon = np.unpackbits(kf.export(bits=1, cutoff=0))[: 1 << (2 * K)]
    both = canon.copy()
    L = oracle_mod.lib()
    for i in np.nonzero(canon)[0].tolist():
        both[L.orc_rev_com_kbit(int(i), K)] = 1
    assert np.array_equal(both, ref_bits)


@pytest.mark.gpu
def test_gpu_table_drives_the_reference_consumer_identically(oracle_mod, tmp_path):
    """SURVEY 8c round trip: the table written by the GPU library is loaded by the shipped correct_error_reads and
    corrects the reads exactly as the table written from the CPU oracle's counts does"""
    from dbg_assembly_b200.kfreq import KmerFreq
    if not os.path.exists(oracle_mod.CORRECT_ELF):
        pytest.skip("oracle/_ref/correct_error_reads_elf not present")
    K, cutoff = 13, 3
    reads = _sim_reads(4, 30_000, 12_000)
    bases, offs = reads_to_arrays(reads)
    counts = oracle_mod.kfreq_count(bases, offs, K)
    cpu_prefix = str(tmp_path / "cpu")
    _write_cz_1bit(cpu_prefix, counts, cutoff)
    gpu_prefix = str(tmp_path / "gpu")
    with KmerFreq(K=K) as kf:
        kf.submit(bases, offs)
        kf.finalize()
        kf.write_cz(gpu_prefix, bits=1, cutoff=cutoff)
    assert open(gpu_prefix + ".kmer.freq.cz.len").read().split() and len(open(gpu_prefix + ".kmer.freq.cz.len").read().split()) == 8
    h_cpu, out_cpu = _run_correct_error(oracle_mod, str(tmp_path), cpu_prefix, reads, K, "cpu")
    h_gpu, out_gpu = _run_correct_error(oracle_mod, str(tmp_path), gpu_prefix, reads, K, "gpu")
    assert h_gpu == h_cpu == int((counts > cutoff).sum())
    assert out_gpu == out_cpu


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


@pytest.mark.gpu
def test_kmerfreq_front_end_writes_the_table_files(oracle_mod, tmp_path):
    """integration/kmerfreq_b200.cpp: same command line as the pipeline's `kmerfreq -k K -m 1 -q Q reads.lib`
    (test/01.clean_correct/work.sh:18); two gzip FASTA files through the threaded reader; the three output files land
    next to the library file and hold the oracle's counts"""
    import gzip
    import subprocess
    exe = os.path.join(REPO, "integration", "_bin", "kmerfreq_b200")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(REPO, "integration"), "all"], check=True, timeout=300)
    K, cutoff = 11, 2
    reads = _sim_reads(9, 20_000, 6_000) + [b"ACGT", b"", b"N" * 40]
    half = len(reads) // 2
    paths = []
    for t, part in enumerate((reads[:half], reads[half:])):
        p = str(tmp_path / f"part{t}.fa.gz")
        with gzip.open(p, "wb", compresslevel=1) as f:
            for i, r in enumerate(part):
                f.write(b">r%d\n%s\n" % (i, r))
        paths.append(p)
    lib = str(tmp_path / "reads.lib")
    with open(lib, "w") as f:
        f.write("\n".join(paths) + "\n")
    p = subprocess.run([exe, "-k", str(K), "-f", "2", "-m", "1", "-q", str(cutoff), lib], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    bases, offs = reads_to_arrays(reads)
    counts = oracle_mod.kfreq_count(bases, offs, K)
    both, canon = oracle_mod.load_cz_1bit(lib, K)
    assert np.array_equal(canon, (counts > cutoff).astype(np.uint8))
    lines = open(lib + ".kmer.freq.stat").read().split("\n")
    assert lines[0] == f"#Kmer size: {K}" and lines[2] == f"#Kmer indivdual number: {int(counts.sum())}"
    assert lines[3] == f"#Kmer species number: {int((counts > 0).sum())}"
