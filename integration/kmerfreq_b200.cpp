// kmerfreq_b200.cpp -- producer of the K-mer frequency table that correct_error loads.
//
// In the reference pipeline this table comes from the EXTERNAL program `kmerfreq`
// (test/01.clean_correct/work.sh:18  `kmerfreq -k 17 -m 1 -q 10 ./clean_reads.lib`), which is not in the reference
// tree; `correct_error_reads -k 17 <lib>.kmer.freq.cz <lib>` then consumes <lib>.kmer.freq.cz + .cz.len.  This front
// end takes the same arguments, reads the same library file (one reads file per line, fastq or one-line fasta, plain
// or .gz, same framing rules as the assembler: fast_reader.h) and writes the same three files next to the library
// file, with the counting, thresholding and bit-packing done on the GPU through the kfreq_* calls of libdbgb200:
//
//     <lib>.kmer.freq.stat     spectrum: species per frequency 1..65535 (layout of test/01.clean_correct/*.stat)
//     <lib>.kmer.freq.cz       zlib blocks of 8 Mi k-mers each: 1 bit (count > -q) or 1 byte (min(255,count)) per k-mer
//     <lib>.kmer.freq.cz.len   compressed length of every block, one per line
//
// PARITY UNPINNED for the option semantics (kmerfreq is external: -q is taken as the low-frequency cutoff, a k-mer being
// high-frequency iff count > q, the rule of correct_error/main.cpp:202; -m is accepted and ignored); the file
// format is pinned against the shipped correct_error_reads (tests/test_kfreq.py).  No CPU fallback: without a CUDA
// device kfreq_create fails and the program exits 1.
//
// Build: make -C integration   (-> integration/_bin/kmerfreq_b200)
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "dbg_b200.h"
#include "fast_reader.h"

using std::cerr;
using std::endl;
using std::string;

static void usage()
{
    cerr << "kmerfreq_b200 [options] <reads_files.lib>\n"
            "   -k <int>   kmer size (<= 17: direct-index table on the GPU), default=17\n"
            "   -f <int>   input file format: 1: fq, 2: fa (one-line), default=1\n"
            "   -m <int>   accepted for command-line compatibility with kmerfreq (the table is always written)\n"
            "   -q <int>   low-frequency cutoff of the 1-bit table: a k-mer is high-frequency iff count > q, default=10\n"
            "   -b <int>   bits per k-mer in the table: 1 (correct_error_reads) or 8 (correct_error), default=1\n"
            "   -r <int>   use only the first <int> bases of every read, default=all\n"
            "   -D <int>   CUDA device, default=0\n"
            "   -h         this help\n";
    exit(0);
}

static void die(const char *where, int rc)
{
    cerr << "libdbgb200: " << where << " failed: " << dbg_strerror(rc) << " -- " << kfreq_last_error() << endl;
    exit(1);
}

int main(int argc, char **argv)
{
    int K = 17, format = 1, cutoff = 10, bits = 1, device = 0;
    long trim = 0;
    int c;
    while ((c = getopt(argc, argv, "k:f:m:q:b:r:D:h")) != -1) {
        switch (c) {
            case 'k': K = atoi(optarg); break;
            case 'f': format = atoi(optarg); break;
            case 'm': break;
            case 'q': cutoff = atoi(optarg); break;
            case 'b': bits = atoi(optarg); break;
            case 'r': trim = atol(optarg); break;
            case 'D': device = atoi(optarg); break;
            default: usage();
        }
    }
    if (optind + 1 != argc) usage();
    if (K < 1 || K > 17) { cerr << "kmerfreq_b200: -k must be 1..17" << endl; return 1; }
    if (format != 1 && format != 2) { cerr << "kmerfreq_b200: -f must be 1 or 2" << endl; return 1; }
    if (bits != 1 && bits != 8) { cerr << "kmerfreq_b200: -b must be 1 or 8" << endl; return 1; }
    const string lib = argv[optind];
    std::vector<string> files;
    {
        std::ifstream in(lib.c_str());
        if (!in) { cerr << "kmerfreq_b200: cannot open " << lib << endl; return 1; }
        string line;
        while (std::getline(in, line)) {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
            if (!line.empty()) files.push_back(line);
        }
    }
    if (files.empty()) { cerr << "kmerfreq_b200: no reads files listed in " << lib << endl; return 1; }

    const auto t0 = std::chrono::steady_clock::now();
    kfreq_ctx *ctx = nullptr;
    int rc = kfreq_create(&ctx, K, device, 0, 1);
    if (rc) die("kfreq_create", rc);

    const uint64_t BLOCK_BASES = 128ull << 20, BLOCK_READS = 2ull << 20;
    const size_t MAX_AHEAD = 4;
    uint64_t total_reads = 0;
    {
        std::vector<std::unique_ptr<dbgio::FileProducer> > prod(files.size());
        size_t started = 0;
        for (size_t i = 0; i < files.size(); i++) {
            for (; started < files.size() && started < i + MAX_AHEAD; started++)
                prod[started].reset(new dbgio::FileProducer(files[started], format, (uint64_t)(trim > 0 ? trim : 1 << 20), BLOCK_BASES,
                                                            BLOCK_READS, dbg_host_alloc, dbg_host_free));
            cerr << "parse reads file: " << files[i] << endl;
            for (;;) {
                dbgio::ReadBlock *b = prod[i]->pop();
                if (prod[i]->failed()) { cerr << "kmerfreq_b200: out of page-locked host memory" << endl; return 1; }
                if (b->n_reads) {
                    if (trim > 0) {
                        // -r: only the first <trim> bases of a read count; the offsets stay, the tail is masked by
                        // shortening every read in place (reads are contiguous: compact them)
                        uint64_t w = 0;
                        for (uint64_t r = 0; r < b->n_reads; r++) {
                            const uint64_t s = b->offs[r], e = b->offs[r + 1];
                            const uint64_t n = (e - s) < (uint64_t)trim ? (e - s) : (uint64_t)trim;
                            if (w != s) memmove(b->bases + w, b->bases + s, n);
                            b->offs[r] = w; w += n;
                        }
                        b->offs[b->n_reads] = w;
                    }
                    rc = kfreq_submit_reads(ctx, b->bases, b->offs, b->n_reads);
                    if (rc) die("kfreq_submit_reads", rc);
                    total_reads += b->n_reads;
                }
                const bool last = b->last;
                prod[i]->recycle(b);
                if (last) break;
            }
            {
                const std::string io = prod[i]->io_error();
                if (!io.empty()) {
                    cerr << "kmerfreq_b200: WARNING: input problem: " << io << endl;
                    if (getenv("DBG_B200_STRICT_IO")) return 1;
                }
            }
            prod[i].reset();
        }
    }
    uint64_t n_occ = 0, n_reads = 0;
    if ((rc = kfreq_finalize(ctx, &n_occ, &n_reads))) die("kfreq_finalize", rc);
    cerr << "reads: " << n_reads << "  kmers: " << n_occ << endl;
    if ((rc = kfreq_write_cz(ctx, lib.c_str(), bits, cutoff))) die("kfreq_write_cz", rc);
    cerr << "wrote " << lib << ".kmer.freq.cz, .kmer.freq.cz.len, .kmer.freq.stat" << endl;
    kfreq_destroy(ctx);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cerr << "Finished! wall clock: " << dt << " s" << endl;
    (void)total_reads;
    return 0;
}
