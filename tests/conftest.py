import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _fresh_reference_binaries():
    """oracle/_ref/ is a git-ignored build product: where the reference sources exist (the build container) bring the
    reference binaries and the relinked front end up to date with the tree (make is incremental), so that no test
    ever runs a stale binary; on the GPU box (no /root/reference) the prebuilt files that travelled are used as is."""
    import subprocess
    if os.path.exists("/root/reference/DBG_contig/DBGgraph.cpp"):
        targets = ["all"]
        if os.path.exists(os.path.join(REPO, "dbg_assembly_b200", "libdbgb200.so")):
            targets.append("b200")
        subprocess.run(["make", "-C", os.path.join(REPO, "oracle")] + targets, check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    yield


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("K", "R", "init_slots", "max_double", "B", "n_files", "size", "count", "conflict", "max", "doublings",
              "reads", "kmers_logged"):
        d[k] = int(d[k])
    d["load"] = float(d["load"]); d["init_g"] = float(d["init_g"])
    d["files"] = [(d[f"bases{i}"], d[f"offs{i}"]) for i in range(d["n_files"])]
    return d


GOLDEN_NO_ENLARGE = ["kat_k5", "ragged_k31", "even_k16_two_files", "saturate_k21", "tiny_k3", "contig_k31"]
GOLDEN_ALL = GOLDEN_NO_ENLARGE + ["enlarge_k25", "maxmem_k25"]


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as orc
    orc.lib()
    return orc


def reads_to_arrays(reads):
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy()
    return bases, offs


def random_reads(seed, n_reads, len_lo, len_hi, genome_len=2000, err=0.02, n_rate=0.01, lower=0.2):
    """small ragged read sets for parity tests (numpy; the generator is not under test)"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    genome = rng.choice(acgt, size=genome_len)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
        comp[a] = b
    reads = []
    for _ in range(n_reads):
        L = int(rng.integers(len_lo, len_hi + 1))
        s = int(rng.integers(0, max(1, genome_len - L + 1)))
        r = genome[s:s + L].copy()
        if rng.random() < 0.5:
            r = comp[r[::-1]]
        m = rng.random(len(r)) < err
        r[m] = rng.choice(acgt, size=int(m.sum()))
        m = rng.random(len(r)) < n_rate
        r[m] = ord("N")
        m = rng.random(len(r)) < lower
        r[m] = r[m] | 0x20
        reads.append(r.tobytes())
    return reads
