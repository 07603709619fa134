"""End-to-end drop-in check on real files: the reference program (oracle/_ref/debruijn_contig_ref) against the same
front end relinked to libdbgb200 (oracle/_ref/debruijn_contig_b200), same command line, same FASTA on disk.
Prints one JSON line with wall times and whether all eight output files are byte-identical.

    python scripts/dropin_compare.py [--reads 1000000] [--genome 1500000]
"""
import argparse, json, os, subprocess, sys, tempfile, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dbg_assembly_b200 import synth          # noqa: E402
from oracle import oracle as orc             # noqa: E402  (checker side: only writes the FASTA and runs the reference)

SUF = (".contig.seq.fa", ".contig.small.fa", ".contig.kmer.freq", ".contig.tip.fa", ".contig.bubble.fa", ".contig.lowedge.fa",
       ".contig.seq.depth", ".contig.small.depth")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--genome", type=int, default=1_500_000)
    ap.add_argument("--init-g", type=float, default=0.08)
    a = ap.parse_args()
    p = synth.make_params(seed=21, genome_len=a.genome, read_len=150, insert=500, err=0.005, n_rate=0.0005)
    bases, offs = synth.reads_host(p, 0, a.reads)
    out = {"reads": a.reads, "occurrences": a.reads * 120, "init_g": a.init_g}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "reads.fa")
        orc.write_fasta(fa, bases, offs)
        lib = os.path.join(td, "reads.lib")
        open(lib, "w").write(fa + "\n")
        runs = {}
        for tag, exe, t in (("ref_t1", orc.REF_CONTIG, 1), ("ref_tN", orc.REF_CONTIG, os.cpu_count() or 1), ("b200", orc.B200_CONTIG, 1)):
            pre = os.path.join(td, tag)
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-k", "31", "-r", "150", "-f", "2", "-t", str(t), "-i", repr(a.init_g), "-M", "100", "-o", pre, lib],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                print(r.stderr.decode()[-1500:], file=sys.stderr)
                raise SystemExit(f"{tag} failed")
            log = r.stderr.decode()
            phase = [l for l in log.splitlines() if l.startswith("libdbgb200 wall clock")]
            if phase:
                out[tag + "_phases"] = phase[0]
            runs[tag] = {s: open(pre + s, "rb").read() for s in SUF}
            out[tag + "_wall_s"] = round(dt, 3)
            out[tag + "_threads"] = t
        out["identical_to_ref_t1"] = all(runs["b200"][s] == runs["ref_t1"][s] for s in SUF)
        out["contig_fa_bytes"] = len(runs["ref_t1"][".contig.seq.fa"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
