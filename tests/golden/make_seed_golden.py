"""Regenerates tests/golden/seed_*.npz by RUNNING THE REFERENCE ITSELF: link_scaffold's read_contig_file, init_kmerset,
chop_contig_to_kmerset and get_align_seed compiled in place (oracle/_ref/ref_seed_driver, oracle/Makefile).

    python tests/golden/make_seed_golden.py

Only runs where /root/reference exists.  Each fixture holds the contigs (as given to map_pair, before its -l filter), the
reads, K / min_ctg_len / seed_kmer_num, and the reference's outputs: table size, count, max, conflict, the filled slots in
slot order with their nodes, and the six numbers get_align_seed returns per read.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as orc  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
RC = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def rnd(rng, n):
    return bytes(rng.choice(ACGT, n))


def rc(s):
    return s[::-1].translate(RC)


def sample_reads(rng, contigs, n, L, max_err=3):
    reads = []
    long_enough = [c for c in contigs if len(c) >= L + 5]
    for _ in range(n):
        c = long_enough[rng.integers(len(long_enough))]
        p = int(rng.integers(0, len(c) - L))
        r = bytearray(c[p:p + L].upper())
        for _e in range(int(rng.integers(0, max_err + 1))):
            r[int(rng.integers(L))] = b"ACGT"[int(rng.integers(4))]
        r = bytes(r)
        reads.append(rc(r) if rng.random() < 0.5 else r)
    return reads


def case_mixed(rng):
    """N gaps, a contig below -l, lower case, a reverse-complemented repeat, poly-A / poly-T runs (k-mer 0 on both strands)"""
    g = rnd(rng, 5000)
    contigs = [g[:1500] + b"NNNNN" + g[1500:2500] + b"N" + rnd(rng, 40) + b"NN" + g[2400:3000], rnd(rng, 90), g[2900:4000].lower(),
               rc(g[3500:4200]), b"A" * 50 + rnd(rng, 200) + b"T" * 40 + rnd(rng, 100) + b"A" * 35, rnd(rng, 300)]
    reads = sample_reads(rng, contigs, 200, 100) + [rnd(rng, 100), b"ACGT" * 5, b"A" * 100, b"T" * 60, rnd(rng, 35), rnd(rng, 36)]
    return dict(K=31, min_ctg_len=125, seed_kmer_num=5, contigs=contigs, reads=reads)


def case_polyT_first(rng):
    """the all-A k-mer is first seen on the reverse strand, then again forward; K = 21, seeds 3 apart"""
    contigs = [rnd(rng, 150) + b"T" * 30 + rnd(rng, 100), b"a" * 25 + rnd(rng, 200), rnd(rng, 400)]
    contigs.append(contigs[2][100:300])
    reads = sample_reads(rng, contigs, 80, 60, max_err=1) + [b"A" * 60, b"T" * 24]
    return dict(K=21, min_ctg_len=100, seed_kmer_num=3, contigs=contigs, reads=reads)


def case_long_blocks(rng):
    """blocks longer than one 32768-k-mer piece, duplicated segments across piece borders, K = 27"""
    g = rnd(rng, 150_000)
    contigs = [g[:70_000], g[65_000:66_000] + b"N" * 10 + g[69_900:150_000], rnd(rng, 500), rc(g[32_700:32_900])]
    reads = sample_reads(rng, contigs, 300, 150)
    return dict(K=27, min_ctg_len=125, seed_kmer_num=5, contigs=contigs, reads=reads)


CASES = {"seed_mixed_k31": (case_mixed, 11), "seed_polyT_k21": (case_polyT_first, 12), "seed_long_k27": (case_long_blocks, 13)}


def main():
    assert orc.have_seed_reference(), "oracle/_ref/ref_seed_driver missing: make -C oracle ref"
    for name, (fn, seed) in CASES.items():
        c = fn(np.random.default_rng(seed))
        names = [b"ctg%d len=%d" % (i, len(s)) for i, s in enumerate(c["contigs"])]
        ref = orc.run_ref_seed(names, c["contigs"], c["reads"], c["K"], c["min_ctg_len"], c["seed_kmer_num"])
        occ = np.unpackbits(ref["nul"])[: ref["size"]].astype(bool)
        slot = np.nonzero(occ)[0].astype(np.uint64)
        cb, co = orc.seqs_to_arrays(c["contigs"])
        rb, ro = orc.seqs_to_arrays(c["reads"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), K=c["K"], min_ctg_len=c["min_ctg_len"], seed_kmer_num=c["seed_kmer_num"],
                            contig_bases=cb, contig_offs=co, read_bases=rb, read_offs=ro,
                            size=ref["size"], count=ref["count"], max=ref["max"], conflict=ref["conflict"],
                            slot=slot, kmer=ref["array"]["kmer"][occ], value=ref["array"]["value"][occ], seeds=ref["seeds"])
        print(name, "size", ref["size"], "count", ref["count"], "seeds found", int((ref["seeds"][:, 0] >= 0).sum()), "of", len(c["reads"]))


if __name__ == "__main__":
    main()
