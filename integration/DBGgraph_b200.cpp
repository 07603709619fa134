// DBGgraph_b200.cpp -- the reference-side binding: a drop-in replacement for DBG_contig/DBGgraph.cpp.
//
// A maintainer of fanagislab/DBG_assembly swaps this file for DBGgraph.cpp in DBG_contig/Makefile and
// links libdbgb200.so; main.cpp (getopt, parameter echo), contig.cpp (traversal), kmerSet.cpp (lookups
// and flag helpers used by the traversal), seqKmer.cpp and gzstream.cpp stay untouched.  Command line,
// parameter meaning, stderr log lines and all nine output files are the reference's.
//
// It defines the globals DBGgraph.h declares (DBGgraph.h:25-49) and build_debruijn_graph()
// (DBGgraph.cpp:364): the host keeps the file reader (same framing as DBGgraph.cpp:244-272; fast_reader.h: zlib
// with large buffers, one decoding thread per file, submission in file order) and hands blocks of reads to the GPU
// library; hot loops #1 and #2
// (thread_parseBlock / thread_updatekmers) run as CUDA kernels; the finished table comes back in the
// reference's slot layout (== `debruijn_contig -t 1`) as the global `KmerSet *kset`.
//
// Differences a user can observe: the "conflict:" statistic counts GPU probe steps, "-t" only affects the
// host traversal, and "-e" (enlarge) is not emulated: if the node count passes max_cutoff a warning is
// printed (the reference would have enlarged and produced a different slot order; contents are the same).
#include <chrono>
#include <memory>
#include <thread>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "DBGgraph.h"     // the reference's header (globals + prototypes), found via -I<reference>/DBG_contig
#include "dbg_b200.h"
#include "fast_reader.h"

// ---- globals of DBGgraph.cpp:10-34 (same names, same defaults) -----------------------------------
int KmerSize = 31;
int maxReadLen = 250;
int KmerNumInRead = 0;
int Input_file_format = 1;
string Output_prefix = "output";
int threadNum = 10;
KmerSet *kset;
double initHashSize = 1.0;
uint64_t maxDoubleHashTimes = 10;
uint64_t doubleHashTimes = 0;
float hashLoadFactor = 0.7;
int BufferNum = 10000;
string *RawReads;
uint64_t *StoreKmer;
uint8_t *StoreLeftBase;
uint8_t *StoreRightBase;
uint8_t *Signal;
uint64_t Kmer_total_num = 0;
uint64_t Total_reads_num = 0;
uint64_t KmerHeadMaskVal = 0;
uint64_t KmerRCOrVal[4];
KmerNode *PolyA;
clock_t time_start;
clock_t time_end;

static void die(const char *where, int rc)
{
    cerr << "libdbgb200: " << where << " failed: " << dbg_strerror(rc) << " -- " << dbg_last_error() << endl;
    exit(1);
}

// Reader of parse_one_reads_file (DBGgraph.cpp:244-272).  The framing rules are the reference's; the mechanics are
// fast_reader.h: every input file is decoded and framed by its own thread into page-locked blocks, and this (main)
// thread submits the blocks in file order, so read ordinals -- and with them the slot layout -- are those of a
// sequential reader.  Blocks never span files.
static const uint64_t BLOCK_BASES = 128ull << 20;
static const uint64_t BLOCK_READS = 2ull << 20;
static const size_t MAX_AHEAD = 4;          // files being decoded at the same time

static void consume_file(dbg_ctx *ctx, dbgio::FileProducer &prod)
{
    uint64_t in_block = 0;
    for (;;) {
        dbgio::ReadBlock *b = prod.pop();
        if (prod.failed()) { cerr << "libdbgb200: out of page-locked host memory for the read blocks" << endl; exit(1); }
        if (b->n_reads) {
            int rc = dbg_submit_reads(ctx, b->bases, b->offs, b->n_reads);
            if (rc) die("dbg_submit_reads", rc);
        }
        // the "Load reads block" lines of the reference: one per BufferNum (-b) reads and one at the end of the file
        uint64_t left = b->n_reads;
        while (in_block + left >= (uint64_t)BufferNum) {
            const uint64_t take = (uint64_t)BufferNum - in_block;
            Total_reads_num += take; left -= take; in_block = 0;
            cerr << "Load reads block " << Total_reads_num << endl;
        }
        Total_reads_num += left; in_block += left;
        const bool last = b->last;
        prod.recycle(b);
        if (last) break;
    }
    cerr << "Load reads block " << Total_reads_num << endl;
    cerr << "this block has reach the end of file " << endl;
}

static double wall_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void build_debruijn_graph(vector<string> &reads_files)
{
    time_start = clock();
    const double w0 = wall_now();

    // contig.h:127-130 needs the head mask (DBGgraph.cpp:371)
    KmerHeadMaskVal = pow_integer(2, KmerSize * 2) - 1;
    KmerRCOrVal[3] = 0;
    KmerRCOrVal[1] = pow_integer(2, KmerSize * 2 - 1);
    KmerRCOrVal[2] = pow_integer(2, KmerSize * 2 - 1 - 1);
    KmerRCOrVal[0] = KmerRCOrVal[1] + KmerRCOrVal[2];
    KmerNumInRead = maxReadLen - KmerSize + 1;

    if (KmerSize > 31) { cerr << "debruijn_contig: -k max 31 for the 64-bit host traversal" << endl; exit(1); }

    cerr << "Start to initialize the kmerset hash" << endl;
    dbg_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.K = KmerSize;
    prm.max_read_len = maxReadLen;
    prm.init_slots = (uint64_t)(initHashSize * 1000000000);   // DBGgraph.cpp:381
    prm.load_factor = hashLoadFactor;
    prm.device = getenv("DBG_B200_DEVICE") ? atoi(getenv("DBG_B200_DEVICE")) : 0;
    prm.track_order = 1;                                        // reproduce the -t 1 slot layout
    dbg_ctx *ctx = NULL;
    int rc = dbg_create(&ctx, &prm);
    if (rc) die("dbg_create", rc);
    // the KmerSet the traversal consumes (kmerSet.h:88-99): allocated page-locked, in the background while the
    // reads are being parsed, so that the export is one PCIe-rate copy instead of a page-faulting pageable one
    const uint64_t P_slots = prm.init_slots < 3 ? 3 : dbg_find_next_prime(prm.init_slots);   // kmerSet.cpp:102-103
    KmerNode *pinned_array = NULL;
    std::thread alloc_thread([&]() {
        void *q = NULL;
        if (dbg_host_alloc(&q, P_slots * sizeof(KmerNode)) == DBG_OK) pinned_array = (KmerNode *)q;
    });
    cerr << "Hash initialization array size:  " << initHashSize << " G" << endl;
    cerr << "The initialization memory used:  " << initHashSize * 16 << " G" << endl;
    time_end = clock();
    cerr << "Finished! Run time: " << double(time_end - time_start) / CLOCKS_PER_SEC << endl;

    const double w1 = wall_now();
    cerr << "\nparse input reads files: " << endl;
    {
        std::vector<std::unique_ptr<dbgio::FileProducer> > prod(reads_files.size());
        size_t started = 0;
        for (size_t i = 0; i < reads_files.size(); i++) {
            for (; started < reads_files.size() && started < i + MAX_AHEAD; started++)
                prod[started].reset(new dbgio::FileProducer(reads_files[started], Input_file_format, (uint64_t)maxReadLen, BLOCK_BASES,
                                                            BLOCK_READS, dbg_host_alloc, dbg_host_free));
            cerr << "\nStart to parse reads file: " << reads_files[i] << endl;
            consume_file(ctx, *prod[i]);
            prod[i].reset();
            dbg_stats st;
            if ((rc = dbg_get_stats(ctx, &st))) die("dbg_get_stats", rc);
            Kmer_total_num = st.kmers_logged;
            cerr << "\nTotal number of reads loaded into memory: " << Total_reads_num << endl;
            cerr << "Total number of kmers loaded into memory: " << Kmer_total_num << endl;
            time_end = clock();
            cerr << "Finished! Run time: " << double(time_end - time_start) / CLOCKS_PER_SEC << endl;
        }
    }

    // add polyA and polyT [kmer: 0] to the kmerset (DBGgraph.cpp:418) happens inside dbg_finalize
    const double w2 = wall_now();
    dbg_stats st;
    if ((rc = dbg_finalize(ctx, &st))) die("dbg_finalize", rc);
    if (st.count - 1 > st.max_cutoff)
        cerr << "\nAlert message: " << st.count << " kmer nodes exceed max_cutoff " << st.max_cutoff
             << "; the CPU program would have enlarged its hash (-e). Node contents are unaffected, slot order may differ: raise -i\n" << endl;

    // the KmerSet the traversal consumes (kmerSet.h:88-99, kmerSet.cpp:98-127)
    kset = new KmerSet;
    kset->e_size = sizeof(KmerNode);
    kset->size = st.array_size;
    kset->count = st.count;
    kset->count_conflict = st.conflict;
    kset->load_factor = st.load_factor;
    kset->max = st.max_cutoff;
    kset->iter_ptr = 0;
    alloc_thread.join();
    kset->array = (pinned_array && P_slots == kset->size) ? pinned_array : (KmerNode *)malloc(kset->size * kset->e_size);
    kset->nul_flag = (uint8_t *)malloc(kset->size / 8 + 1);
    kset->del_flag = (uint8_t *)calloc(kset->size / 8 + 1, 1);
    if (!kset->array || !kset->nul_flag || !kset->del_flag) { cerr << "out of host memory for the kmerset" << endl; exit(1); }
    const double w3 = wall_now();
    if ((rc = dbg_export_kmerset(ctx, kset->array, kset->nul_flag))) die("dbg_export_kmerset", rc);
    const double w4 = wall_now();
    cerr << "libdbgb200 wall clock (s): init " << w1 - w0 << ", read files + submit " << w2 - w1 << ", finalize (GPU build + layout) "
         << w3 - w2 << ", export kmerset " << w4 - w3 << endl;

    dbg_destroy(ctx);

    print_kmerset_parameter(kset);
}
