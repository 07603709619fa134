"""CPU end-to-end test of the drop-in front end's HOST logic (integration/DBGgraph_b200.cpp): the reference's own
main.cpp / contig.cpp plus the binding are linked against tests/mock/dbg_mock.cpp -- a stand-in for the device-side
entry points built on the oracle -- in front of the real libdbgb200 (which still provides dbg_replay_growth,
dbg_find_next_prime ...).  The program then runs without a GPU, and its eight output files are compared byte for byte
with the reference program's: normal run, a run where the reference enlarges its hash three times (device-table retry
+ host-side growth replay), and a run that exhausts -e (refused loudly).  Needs the reference sources to compile the
front end (this container); skipped elsewhere.  The same flows run against the real library in the GPU tests."""
import os
import subprocess

import pytest

from conftest import REPO, random_reads, reads_to_arrays

REFSRC = "/root/reference/DBG_contig"
REF_BIN = os.path.join(REPO, "oracle", "_ref", "debruijn_contig_ref")
SUF = (".contig.seq.fa", ".contig.small.fa", ".contig.kmer.freq", ".contig.tip.fa", ".contig.bubble.fa", ".contig.lowedge.fa",
       ".contig.seq.depth", ".contig.small.depth")


@pytest.fixture(scope="module")
def mock_front_end(tmp_path_factory, oracle_mod):
    gen = os.path.join(REPO, "oracle", "_ref", "gen", "kmerSet.cpp")
    if not (os.path.exists(os.path.join(REFSRC, "main.cpp")) and os.path.exists(gen) and os.access(REF_BIN, os.X_OK)):
        pytest.skip("reference sources / oracle/_ref not present")
    from dbg_assembly_b200 import capi
    capi.load()                                                       # makes sure libdbgb200.so is built
    d = str(tmp_path_factory.mktemp("mockfe"))
    inc = ["-I" + os.path.join(REPO, "include"), "-I" + os.path.join(REPO, "oracle"), "-I" + os.path.join(REPO, "integration")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + inc + ["-o", os.path.join(d, "libdbgmock.so"),
                    os.path.join(REPO, "tests", "mock", "dbg_mock.cpp"), "-L" + os.path.join(REPO, "oracle"), "-loracle",
                    "-Wl,-rpath," + os.path.join(REPO, "oracle")], check=True, timeout=300)
    exe = os.path.join(d, "debruijn_contig_mock")
    srcs = [os.path.join(REPO, "integration", "DBGgraph_b200.cpp"), gen] + \
           [os.path.join(REFSRC, f) for f in ("seqKmer.cpp", "gzstream.cpp", "contig.cpp", "global_aligning.cpp", "main.cpp")]
    subprocess.run(["g++", "-O2", "-w", "-I" + REFSRC, "-I" + os.path.join(REPO, "oracle", "shim")] + inc + ["-o", exe] + srcs +
                   ["-L" + d, "-ldbgmock", "-L" + os.path.join(REPO, "dbg_assembly_b200"), "-ldbgb200", "-L" + os.path.join(REPO, "oracle"), "-loracle",
                    "-Wl,-rpath," + d, "-Wl,-rpath," + os.path.join(REPO, "dbg_assembly_b200"), "-Wl,-rpath," + os.path.join(REPO, "oracle"),
                    "-lz", "-lpthread"], check=True, timeout=600)
    return exe


@pytest.fixture(scope="module")
def reads_lib(tmp_path_factory, oracle_mod):
    d = tmp_path_factory.mktemp("fe_reads")
    reads = random_reads(98, 3000, 100, 100, genome_len=15000, err=0.004, n_rate=0.0, lower=0.0)
    paths = []
    for i, part in enumerate((reads[:1800], reads[1800:])):
        bases, offs = reads_to_arrays(part)
        p = str(d / f"g{i}.fa"); oracle_mod.write_fasta(p, bases, offs); paths.append(p)
    lib = str(d / "reads.lib")
    with open(lib, "w") as f:
        f.write("\n".join(paths) + "\n")
    return lib


def run(exe, lib, pre, extra):
    r = subprocess.run([exe, "-k", "25", "-r", "100", "-f", "2", "-t", "1", "-M", "100", "-o", pre, lib] + extra,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    files = {s: open(pre + s, "rb").read() for s in SUF if os.path.exists(pre + s)}
    return r.returncode, files, r.stderr.decode()


@pytest.mark.parametrize("label,extra,size,grows", [
    ("normal", ["-i", "0.0005"], 500009, 0),
    ("enlarge3", ["-i", "0.000008", "-b", "100", "-e", "10"], 64151, 3),
    ("count_over_max_in_final_block", ["-i", "0.00004", "-b", "100000"], 40009, 0),   # count > max but no grow check: no growth
])
def test_front_end_files_identical_to_the_reference(mock_front_end, reads_lib, tmp_path, label, extra, size, grows):
    rc_r, files_r, log_r = run(REF_BIN, reads_lib, str(tmp_path / "ref"), extra)
    rc_m, files_m, log_m = run(mock_front_end, reads_lib, str(tmp_path / "b200"), extra)
    assert rc_r == 0 and rc_m == 0, log_m[-1500:]
    assert f"array_size:\t{size}" in log_r and f"array_size:\t{size}" in log_m
    assert log_r.count("Enlarge hash array size") == grows and log_m.count("Enlarge hash array size") == (1 if grows else 0)
    assert (f"Hash enlarged {grows} time(s)" in log_m) == (grows > 0)
    assert len(files_r) == len(SUF) and files_r == files_m
    assert len(files_r[".contig.seq.fa"]) > 1000


@pytest.mark.parametrize("extra,loaded", [
    (["-i", "0.000008", "-b", "100", "-e", "1"], (300, 400)),      # -e exhausted in file 0; file 1 contributes one block
    (["-i", "0.00002", "-b", "150", "-e", "0"], (450, 600)),       # no doubling allowed at all
    (["-i", "0.00001", "-b", "50", "-e", "1"], (350, 400)),
    (["-i", "0.00003", "-b", "400", "-e", "0"], (1200, 1600)),
])
def test_front_end_reproduces_the_reference_when_it_drops_reads(mock_front_end, reads_lib, tmp_path, extra, loaded):
    """-e exhausted (DBGgraph.cpp:346-350): the reference stops reading the current file and uses only the first block of
    every later one.  The binding finds that point with the growth replay of a full build and rebuilds on exactly the
    reads the reference used: same eight files, same alert lines."""
    rc_r, files_r, log_r = run(REF_BIN, reads_lib, str(tmp_path / "ref"), extra)
    rc_m, files_m, log_m = run(mock_front_end, reads_lib, str(tmp_path / "b200"), extra)
    assert rc_r == 0 and rc_m == 0, log_m[-1500:]
    assert "Memory reach the maximum allowed" in log_r
    if loaded:
        for n in loaded:
            assert f"program have loaded {n} reads" in log_r and f"program have loaded {n} reads" in log_m
    size = [l for l in log_r.split("\n") if l.startswith("array_size:")]
    assert size and size == [l for l in log_m.split("\n") if l.startswith("array_size:")]
    assert len(files_r) == len(SUF) and files_r == files_m


def test_front_end_gzip_fastq_input(mock_front_end, oracle_mod, tmp_path):
    """-f 1 on .fq.gz files (quality lines starting with '@' included): threaded zlib reader + binding vs the reference
    program reading the same files through gzstream"""
    import gzip
    reads = random_reads(77, 2500, 100, 100, genome_len=12000, err=0.004, n_rate=0.0, lower=0.0)
    paths = []
    for i, part in enumerate((reads[:900], reads[900:])):
        p = str(tmp_path / f"r{i}.fq.gz")
        with gzip.open(p, "wb", compresslevel=1) as f:
            for j, r in enumerate(part):
                q = (b"@" if j % 4 == 0 else b"I") + b"I" * (len(r) - 1)
                f.write(b"@read%d\n%s\n+\n%s\n" % (j, r, q))
        paths.append(p)
    lib = str(tmp_path / "fq.lib")
    with open(lib, "w") as f:
        f.write("\n".join(paths) + "\n")
    outs = {}
    for tag, exe in (("ref", REF_BIN), ("b200", mock_front_end)):
        pre = str(tmp_path / tag)
        r = subprocess.run([exe, "-k", "25", "-r", "100", "-f", "1", "-t", "1", "-i", "0.0005", "-M", "100", "-o", pre, lib],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-1500:]
        outs[tag] = {s: open(pre + s, "rb").read() for s in SUF}
        assert "Total number of reads loaded into memory: 2500" in r.stderr.decode()
    assert outs["ref"] == outs["b200"] and len(outs["ref"][".contig.seq.fa"]) > 1000


def test_front_end_checkpoint_restart(mock_front_end, reads_lib, tmp_path):
    """DBG_B200_CHECKPOINT: the first run writes the finished KmerSet, the second one loads it, parses no reads, and still
    writes the reference's files -- also with another traversal cut-off (-D), where it must equal the reference run with
    that cut-off"""
    ck = str(tmp_path / "graph.ckpt")
    env = dict(os.environ, DBG_B200_CHECKPOINT=ck)
    extra = ["-i", "0.0005"]

    def run_env(pre, more, e):
        r = subprocess.run([mock_front_end, "-k", "25", "-r", "100", "-f", "2", "-t", "1", "-M", "100", "-o", pre, reads_lib] + extra + more,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, env=e)
        return r.returncode, {s: open(pre + s, "rb").read() for s in SUF if os.path.exists(pre + s)}, r.stderr.decode()
    rc1, f1, log1 = run_env(str(tmp_path / "a"), [], env)
    assert rc1 == 0 and "checkpoint written" in log1 and os.path.getsize(ck) > 1000
    rc2, f2, log2 = run_env(str(tmp_path / "b"), [], env)
    assert rc2 == 0 and "loaded from checkpoint" in log2 and "Start to parse reads file" not in log2
    assert f1 == f2 and len(f1) == len(SUF)
    rc3, f3, log3 = run_env(str(tmp_path / "c"), ["-D", "4"], env)
    rc_r, f_r, _ = run(REF_BIN, reads_lib, str(tmp_path / "ref"), extra + ["-D", "4"])
    assert rc3 == 0 and rc_r == 0 and "loaded from checkpoint" in log3
    assert f3 == f_r and f3 != f1
