"""CPU tests of the drop-in front end's reader (integration/fast_reader.h): the framing rules of
parse_one_reads_file (DBG_contig/DBGgraph.cpp:244-272) restated in Python, against the C++ reader run through
integration/reader_check.cpp -- plain and gzip input, FASTA and FASTQ, CRLF, junk lines, a header on the last line,
a missing final newline, lines longer than the read buffer, blocks smaller than the file, several files at once."""
import gzip
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def reader_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("reader") / "reader_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(REPO, "integration", "reader_check.cpp"), "-lz", "-lpthread"],
                   check=True, timeout=300)
    return exe


def frame(data: bytes, fmt: int):
    """std::getline framing of DBGgraph.cpp:244-272 (a failed getline yields an empty read)"""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                      # no line after the final '\n'
    hdr = b"@" if fmt == 1 else b">"
    reads, i = [], 0
    while i < len(lines):
        line = lines[i]; i += 1
        if line[:1] != hdr:
            continue
        reads.append(lines[i] if i < len(lines) else b"")
        i += 1
        if fmt == 1:
            i += 2
    return reads


def fnv(reads, trim, block_bases):
    h = 1469598103934665603
    M = (1 << 64) - 1
    n_bases = 0
    for r in reads:
        if len(r) > block_bases:
            r = r[:trim]
        for k in range(8):
            h ^= (len(r) >> (8 * k)) & 0xFF; h = (h * 1099511628211) & M
        for c in r:
            h ^= c; h = (h * 1099511628211) & M
        n_bases += len(r)
    return len(reads), n_bases, h


def run(exe, fmt, block_bases, block_reads, trim, paths):
    out = subprocess.run([exe, str(fmt), str(block_bases), str(block_reads), str(trim)] + paths, check=True, timeout=120,
                         stdout=subprocess.PIPE).stdout.decode().split("\n")
    return [tuple(int(x) for x in line.split()) for line in out if line]


def rand_seq(rng, n):
    return bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tolist())


CASES = {
    "fa_plain": (2, lambda r: b"".join(b">r%d\n%s\n" % (i, rand_seq(r, int(r.integers(1, 300)))) for i in range(3000))),
    "fq_plain": (1, lambda r: b"".join(b"@r%d\n%s\n+\n%s\n" % (i, rand_seq(r, 100), b"I" * 100) for i in range(2000))),
    "fa_crlf": (2, lambda r: b"".join(b">r%d\r\n%s\r\n" % (i, rand_seq(r, 50)) for i in range(500))),
    "fa_no_final_newline": (2, lambda r: b">a\nACGT\n>b\nGGCC"),
    "fa_header_last": (2, lambda r: b">a\nACGT\n>b"),
    "fa_header_last_nl": (2, lambda r: b">a\nACGT\n>b\n"),
    "fq_truncated": (1, lambda r: b"@a\nACGT\n+\nIIII\n@b\nGGGG\n+"),
    "fa_junk_and_blank": (2, lambda r: b"\n\njunk\n>a\nAC\nignored\n\n>b\n\n>c\n>d\nTT\n"),
    "fq_quality_at": (1, lambda r: b"@a\nACGT\n+\n@@@@\n@b\nGG\n+\n@I\n"),     # '@' quality lines are skipped, not headers
    "fa_multiline": (2, lambda r: b">a\nACGT\nTTTT\n>b\nGG\n"),                # only the first line after a header is the read
    "empty": (2, lambda r: b""),
    "fa_long_line": (2, lambda r: b">big\n" + rand_seq(r, 9_500_000) + b"\n>s\nACGT\n"),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("gz", [False, True])
def test_reader_framing(reader_check, tmp_path, name, gz):
    fmt, make = CASES[name]
    data = make(np.random.default_rng(abs(hash(name)) % 1000))
    path = str(tmp_path / (name + (".gz" if gz else ".txt")))
    with (gzip.open(path, "wb", compresslevel=1) if gz else open(path, "wb")) as fh:
        fh.write(data)
    reads = frame(data, fmt)
    for block_bases, block_reads in ((1 << 24, 1 << 16), (4096, 7)):
        if name == "fa_long_line" and block_bases < (1 << 24):
            continue
        got = run(reader_check, fmt, block_bases, block_reads, 150, [path])
        assert got == [fnv(reads, 150, block_bases)], (name, block_bases)


def test_reader_oversize_read_is_trimmed(reader_check, tmp_path):
    data = b">big\n" + b"ACGT" * 5000 + b"\n>s\nACGT\n"
    path = str(tmp_path / "big.fa")
    with open(path, "wb") as fh:
        fh.write(data)
    got = run(reader_check, 2, 4096, 16, 150, [path])
    assert got == [fnv(frame(data, 2), 150, 4096)] and got[0][1] == 150 + 4


def test_reader_reports_damaged_and_missing_input(reader_check, tmp_path):
    """gzread() <= 0 is a clean EOF only if zlib agrees: a truncated .gz and a missing file still deliver what the
    reference's reader would have used (the decodable prefix / nothing), but the producer says so (io_error)"""
    rng = np.random.default_rng(5)
    body = b"".join(b">r%d\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=90)) + b"\n" for i in range(4000))
    good = str(tmp_path / "good.fa.gz"); open(good, "wb").write(gzip.compress(body))
    bad = str(tmp_path / "cut.fa.gz"); open(bad, "wb").write(gzip.compress(body)[:20000])
    missing = str(tmp_path / "nope.fa")
    rr = subprocess.run([reader_check, "2", "65536", "512", "150", good, bad, missing], check=True, timeout=120,
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    lines = [tuple(int(x) for x in l.split()) for l in rr.stdout.decode().split("\n") if l]
    assert lines[0] == fnv(frame(body, 2), 150, 65536)
    assert 0 < lines[1][0] < 4000 and lines[2][0] == 0
    err = rr.stderr.decode()
    assert "cut.fa.gz" in err and "cannot open" in err and "good.fa.gz" not in err


def test_reader_many_files_in_order(reader_check, tmp_path):
    rng = np.random.default_rng(5)
    paths, exp = [], []
    for f in range(6):
        data = b"".join(b">f%d_%d\n%s\n" % (f, i, rand_seq(rng, int(rng.integers(20, 200)))) for i in range(400 + 50 * f))
        p = str(tmp_path / f"f{f}.fa.gz")
        with gzip.open(p, "wb", compresslevel=1) as fh:
            fh.write(data)
        paths.append(p); exp.append(fnv(frame(data, 2), 150, 8192))
    assert run(reader_check, 2, 8192, 64, 150, paths) == exp
    missing = str(tmp_path / "nope.fa")
    assert run(reader_check, 2, 8192, 64, 150, [missing, paths[0]]) == [(0, 0, 1469598103934665603), exp[0]]


def test_framing_rules_are_the_reference_readers(tmp_path):
    """the framing restated above (and implemented by fast_reader.h) against the REAL reader: the reference's build
    driver (oracle/_ref/ref_build_driver: parse_one_reads_file with igzstream + getline) reads awkward FASTQ / FASTA
    files -- quality lines starting with '@', junk between records, a missing final newline -- and must end with exactly
    the table the oracle builds from the reads this framing extracts"""
    from oracle import oracle as orc
    if not orc.have_reference():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    recs_fq, recs_fa = [], []
    for i in range(400):
        s = rand_seq(rng, int(rng.integers(30, 90))).replace(b"N", b"A")
        q = (b"@" if i % 3 == 0 else b"I") + b"I" * (len(s) - 1)          # '@' first: must be skipped, not taken as a header
        recs_fq.append(b"@r%d\n%s\n+\n%s\n" % (i, s, q))
        recs_fa.append((b"junk line\n" if i % 7 == 0 else b"") + b">r%d\n%s\n" % (i, s) + (b"\n" if i % 5 == 0 else b""))
    cases = {1: b"".join(recs_fq)[:-1], 2: b"".join(recs_fa)}                   # FASTQ without the final newline
    for fmt, data in cases.items():
        path = str(tmp_path / f"in{fmt}.txt")
        with open(path, "wb") as fh:
            fh.write(data)
        reads = frame(data, fmt)
        assert len(reads) == 400
        stats, ref = orc.run_ref_build([path], 21, 100, 0.0001, threads=1, fmt=fmt)
        lens = np.array([len(r) for r in reads], dtype=np.uint64)
        offs = np.zeros(len(reads) + 1, dtype=np.uint64); np.cumsum(lens, out=offs[1:])
        bases = np.frombuffer(b"".join(reads), dtype=np.uint8)
        o = orc.OracleGraph(21, 100, 100000, 0.7, 10, 10000)
        o.add_file(bases, offs); o.finish()
        e = o.dump()
        assert ref["size"] == o.size and ref["count"] == o.count
        for k in ("slot", "kmer", "l", "r"):
            assert np.array_equal(ref[k], e[k]), (fmt, k)
        o.close()
