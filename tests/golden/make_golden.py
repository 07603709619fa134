"""Regenerates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref/ref_build_driver, built by
oracle/Makefile from /root/reference/DBG_contig in place; the shipped ELF cross-checks the contig files).

    python tests/golden/make_golden.py

Only runs where /root/reference exists (the build container).  The .npz files are committed; tests on
the GPU box read them and never touch /root/reference.  Each fixture holds the input reads, the
parameters, and the reference's table: filled slots in slot order with (kmer, l_link, r_link).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as orc  # noqa: E402


def reads_to_arrays(reads):
    lens = np.array([len(r) for r in reads], dtype=np.uint64)
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy()
    return bases, offs


def random_genome(rng, n):
    return rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)


def sample_reads(rng, genome, n_reads, len_lo, len_hi, err=0.01, n_rate=0.0, lower=0.0, rc_frac=0.5):
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
        comp[a] = b
    reads = []
    G = len(genome)
    for _ in range(n_reads):
        L = int(rng.integers(len_lo, len_hi + 1))
        s = int(rng.integers(0, max(1, G - L + 1)))
        r = genome[s:s + L].copy()
        if rng.random() < rc_frac:
            r = comp[r[::-1]]
        m = rng.random(len(r)) < err
        r[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
        m = rng.random(len(r)) < n_rate
        r[m] = ord("N")
        if lower > 0:
            m = rng.random(len(r)) < lower
            r[m] = r[m] | 0x20
        reads.append(r.tobytes())
    return reads


def distinct_nodes(files_reads, K, R):
    """distinct canonical k-mers (incl. the k-mer-0 node), counted with the oracle on a roomy table"""
    occ = sum(max(0, min(len(r), R) - K + 1) for reads in files_reads for r in reads)
    o = orc.OracleGraph(K, R, max(1000, 4 * occ), 0.7, 10, 1 << 40)
    for reads in files_reads:
        o.add_file(*reads_to_arrays(reads))
    o.finish()
    n = o.count
    o.close()
    return n


def g_for_load(files_reads, K, R, load):
    """-i value (in G slots) that puts the table at about `load` occupancy"""
    return (distinct_nodes(files_reads, K, R) / load) / 1e9


def run_case(name, files_reads, K, R, init_g, load=0.7, max_double=10, B=10000, threads=1, contig=False):
    """files_reads: list of read lists (one per input file).  NB: a single-block file never triggers the
    reference's grow check (DBGgraph.cpp:329-331), so init_g must leave room or the reference spins forever."""
    with tempfile.TemporaryDirectory() as td:
        paths = []
        arrays = []
        for i, reads in enumerate(files_reads):
            bases, offs = reads_to_arrays(reads)
            p = os.path.join(td, f"f{i}.fa")
            orc.write_fasta(p, bases, offs)
            paths.append(p)
            arrays.append((bases, offs))
        stats, d = orc.run_ref_build(paths, K, R, init_g, threads=threads, load=load, max_double=max_double, buffer_reads=B)
        out = dict(K=K, R=R, init_g=init_g, init_slots=int(float(init_g) * 1000000000), load=load, max_double=max_double, B=B,
                   n_files=len(files_reads), size=d["size"], count=d["count"], conflict=stats["conflict"],
                   max=stats["max"], doublings=stats["doublings"], reads=stats["reads"], kmers_logged=stats["kmers_logged"],
                   slot=d["slot"], kmer=d["kmer"], l=d["l"], r=d["r"])
        for i, (bases, offs) in enumerate(arrays):
            out[f"bases{i}"] = bases
            out[f"offs{i}"] = offs
        if contig:
            # full program: reference rebuilt from source AND the shipped ELF must agree byte for byte
            lib = os.path.join(td, "reads.lib")
            with open(lib, "w") as f:
                f.write("\n".join(paths) + "\n")
            outs = {}
            for tag, exe in (("ref", orc.REF_CONTIG), ("elf", orc.REF_ELF)):
                pre = os.path.join(td, tag)
                subprocess.run([exe, "-k", str(K), "-r", str(R), "-f", "2", "-t", "1", "-i", repr(float(init_g)), "-l", repr(float(load)),
                                "-M", "100", "-o", pre, lib], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
                outs[tag] = {}
                for suf in (".contig.seq.fa", ".contig.small.fa", ".contig.kmer.freq", ".contig.tip.fa", ".contig.bubble.fa",
                            ".contig.lowedge.fa", ".contig.seq.depth", ".contig.small.depth"):
                    with open(pre + suf, "rb") as f:
                        outs[tag][suf] = f.read()
            assert outs["ref"] == outs["elf"], "rebuilt reference and shipped ELF disagree"
            for suf, data in outs["ref"].items():
                out["file" + suf.replace(".", "_")] = np.frombuffer(data, dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: P={d['size']} count={d['count']} conflict={stats['conflict']} doublings={stats['doublings']}")


def main():
    assert orc.have_reference(), "needs oracle/_ref (run `make -C oracle` in the build container)"
    rng = np.random.default_rng(20261018)

    # 1. the survey's known-answer test
    run_case("kat_k5", [[b"ACGTTGCAAN", b"TTGCAACGT", b"AAAAAAA", b"ACG", b"acgttgcaatt"]], K=5, R=10, init_g=1e-6)

    # 2. ragged reads, N and lower case, short reads, high load -> long probe clusters and wrap-around
    g = random_genome(rng, 3000)
    reads = sample_reads(rng, g, 400, 20, 120, err=0.02, n_rate=0.01, lower=0.3)
    reads += [b"", b"A", b"ACGT" * 7, b"N" * 40, b"T" * 50]
    run_case("ragged_k31", [reads], K=31, R=100, init_g=g_for_load([reads], 31, 100, 0.85), load=0.7)

    # 3. even K (palindromes tie -> forward), several files, trimming to -r
    g = random_genome(rng, 1500)
    f0 = sample_reads(rng, g, 150, 30, 90, err=0.01)
    f1 = sample_reads(rng, g, 150, 30, 90, err=0.01) + [b"ACGT" * 10, b"AATT" * 10, b"GC" * 30]
    run_case("even_k16_two_files", [f0, f1], K=16, R=60, init_g=g_for_load([f0, f1], 16, 60, 0.6))

    # 4. saturation at 255 and the poly-A side node
    g = random_genome(rng, 200)
    reads = sample_reads(rng, g, 1500, 40, 60, err=0.0) + [b"A" * 60] * 300 + [b"T" * 45] * 10 + [b"AAAAAAAAAAAAAAAAAAAAAAAAAC"]
    run_case("saturate_k21", [reads], K=21, R=60, init_g=g_for_load([reads], 21, 60, 0.5))

    # 5. the enlarge path (oracle only: the GPU library sizes the table up front), blocks of 50 reads
    g = random_genome(rng, 4000)
    reads = sample_reads(rng, g, 600, 40, 80, err=0.02)
    # (free space after an enlarge, 0.3 * size, must exceed the <= 20 * 56 new k-mers one block can add,
    #  or the reference itself spins forever inside the block)
    run_case("enlarge_k25", [reads], K=25, R=80, init_g=4e-6, B=20, max_double=10)

    # 6. "memory reach the maximum": -e 1, the rest of the file is ignored, the second file still starts
    run_case("maxmem_k25", [reads, reads[:120]], K=25, R=80, init_g=4e-6, B=20, max_double=1)

    # 7. small K
    g = random_genome(rng, 300)
    run_case("tiny_k3", [sample_reads(rng, g, 60, 1, 20, err=0.0, n_rate=0.05)], K=3, R=15, init_g=1e-7)   # <= 32 distinct 3-mers in 101 slots

    # 8. a small assembly with the full contig outputs (byte parity of the host traversal on our table)
    g = random_genome(rng, 20000)
    reads = sample_reads(rng, g, 6000, 100, 100, err=0.005)
    run_case("contig_k31", [reads[:3000], reads[3000:]], K=31, R=100, init_g=1e-3, contig=True)   # ~0.2 M nodes in 1 M slots


if __name__ == "__main__":
    main()
