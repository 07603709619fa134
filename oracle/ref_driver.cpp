// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Harness around the UNMODIFIED reference build phase.  oracle/Makefile compiles this file together
// with /root/reference/DBG_contig/{DBGgraph,kmerSet,seqKmer,gzstream}.cpp (in place, no copies in
// the repo) into oracle/_ref/ref_build_driver.  It sets the reference's globals exactly like its
// main() does (/root/reference/DBG_contig/main.cpp:166-193), calls build_debruijn_graph()
// (DBGgraph.cpp:364) and then
//   * prints one JSON line with the table statistics and steady_clock wall time of the build, and
//   * optionally dumps every filled slot of the global `kset` in SLOT ORDER
//     (u64 slot, u64 kmer, u32 l_link, u32 r_link) so tests can compare node contents *and* the
//     reference's slot layout.
//
// Used by: tests/ (golden-vector generation, GPU parity at small sizes), bench.py cpu_baseline leg
// and `bench.py --impl reference`.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "DBGgraph.h"

static void die_usage()
{
    fprintf(stderr,
            "ref_build_driver [-k K] [-r maxReadLen] [-f 1|2] [-t threads] [-i initG] [-l load]\n"
            "                 [-e maxDoubles] [-b bufferReads] [-d dump.bin] reads_file...\n");
    exit(2);
}

int main(int argc, char **argv)
{
    std::string dump_path;
    int c;
    while ((c = getopt(argc, argv, "k:r:f:t:i:l:e:b:d:h")) != -1) {
        switch (c) {
        case 'k': KmerSize = atoi(optarg); break;
        case 'r': maxReadLen = atoi(optarg); break;
        case 'f': Input_file_format = atoi(optarg); break;
        case 't': threadNum = atoi(optarg); break;
        case 'i': initHashSize = atof(optarg); break;
        case 'l': hashLoadFactor = atof(optarg); break;
        case 'e': maxDoubleHashTimes = atoi(optarg); break;
        case 'b': BufferNum = atoi(optarg); break;
        case 'd': dump_path = optarg; break;
        default: die_usage();
        }
    }
    if (optind >= argc) die_usage();
    std::vector<std::string> files;
    for (int i = optind; i < argc; i++) files.push_back(argv[i]);

    auto t0 = std::chrono::steady_clock::now();
    build_debruijn_graph(files);
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();

    // occurrences actually inserted = sum over filled link lanes is lossy (saturation), so report
    // the reference's own counters; Kmer_total_num uses the UNTRIMMED read length (DBGgraph.cpp:101).
    printf("{\"reads\": %llu, \"kmers_logged\": %llu, \"array_size\": %llu, \"count\": %llu, "
           "\"conflict\": %llu, \"max\": %llu, \"doublings\": %llu, \"threads\": %d, \"wall_s\": %.6f}\n",
           (unsigned long long)Total_reads_num, (unsigned long long)Kmer_total_num,
           (unsigned long long)kset->size, (unsigned long long)kset->count,
           (unsigned long long)kset->count_conflict, (unsigned long long)kset->max,
           (unsigned long long)doubleHashTimes, threadNum, wall);

    if (!dump_path.empty()) {
        FILE *fp = fopen(dump_path.c_str(), "wb");
        if (!fp) { perror("dump"); return 1; }
        uint64_t hdr[3] = {0x4442474b53455431ULL /* "DBGKSET1" */, kset->size, kset->count};
        fwrite(hdr, sizeof(hdr), 1, fp);
        uint64_t written = 0;
        for (uint64_t i = 0; i < kset->size; i++) {
            if (is_entity_null(kset->nul_flag, i)) continue;
            uint64_t rec[3];
            rec[0] = i;
            rec[1] = kset->array[i].kmer;
            rec[2] = (uint64_t)kset->array[i].l_link | ((uint64_t)kset->array[i].r_link << 32);
            fwrite(rec, sizeof(rec), 1, fp);
            written++;
        }
        fclose(fp);
        if (written != kset->count) {
            fprintf(stderr, "ref_build_driver: nul_flag population %llu != count %llu\n",
                    (unsigned long long)written, (unsigned long long)kset->count);
            return 3;
        }
    }
    return 0;
}
