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
// Differences a user can observe: the "conflict:" statistic counts GPU probe steps, "-t" only affects the host
// traversal.  Table growth ("-e", enlarge) never happens on the GPU; when the reference would have grown its table
// the post-growth slot layout is replayed on the host (dbg_replay_growth).  Only when "-e" would have been exhausted
// (the reference then ignores the rest of a file) does the result differ: all reads are used, and an alert says so.
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

// returns false when the device table turned out to be too small for the input (the caller rebuilds with a larger one)
static bool consume_file(dbg_ctx *ctx, dbgio::FileProducer &prod, bool quiet)
{
    uint64_t in_block = 0;
    bool full = false;
    for (;;) {
        dbgio::ReadBlock *b = prod.pop();
        if (prod.failed()) { cerr << "libdbgb200: out of page-locked host memory for the read blocks" << endl; exit(1); }
        if (b->n_reads) {
            int rc = dbg_submit_reads(ctx, b->bases, b->offs, b->n_reads);
            if (rc == DBG_ERR_TABLE_FULL) full = true;
            else if (rc) die("dbg_submit_reads", rc);
        }
        if (full) { prod.recycle(b); return false; }      // the producer is cancelled by its owner
        // the "Load reads block" lines of the reference: one per BufferNum (-b) reads and one at the end of the file
        uint64_t left = b->n_reads;
        while (in_block + left >= (uint64_t)BufferNum) {
            const uint64_t take = (uint64_t)BufferNum - in_block;
            Total_reads_num += take; left -= take; in_block = 0;
            if (!quiet) cerr << "Load reads block " << Total_reads_num << endl;
        }
        Total_reads_num += left; in_block += left;
        const bool last = b->last;
        prod.recycle(b);
        if (last) break;
    }
    if (!quiet) {
        cerr << "Load reads block " << Total_reads_num << endl;
        cerr << "this block has reach the end of file " << endl;
    }
    return !full;
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
    const uint64_t ref_init_slots = prm.init_slots;              // what the reference sizes (and grows) its table from
    // the KmerSet the traversal consumes (kmerSet.h:88-99): allocated page-locked, in the background while the
    // reads are being parsed, so that the export is one PCIe-rate copy instead of a page-faulting pageable one
    const uint64_t P_slots = ref_init_slots < 3 ? 3 : dbg_find_next_prime(ref_init_slots);   // kmerSet.cpp:102-103
    KmerNode *pinned_array = NULL;
    std::thread alloc_thread([&]() {
        void *q = NULL;
        if (dbg_host_alloc(&q, P_slots * sizeof(KmerNode)) == DBG_OK) pinned_array = (KmerNode *)q;
    });
    cerr << "Hash initialization array size:  " << initHashSize << " G" << endl;
    cerr << "The initialization memory used:  " << initHashSize * 16 << " G" << endl;

    // The device table never grows.  If -i turns out too small to even HOLD the nodes (the reference would have
    // enlarged its table, -e), the build is simply redone with a larger device table: the reads are streamed again, a
    // full build takes milliseconds, and the table the traversal gets is laid out by the growth replay further down from
    // the reference's own -i, so the size of the device table never shows.
    dbg_ctx *ctx = NULL;
    int rc = 0;
    dbg_stats st;
    std::vector<uint64_t> reads_per_file;
    double w1 = 0, w2 = 0;
    for (int attempt = 0;; attempt++) {
        rc = dbg_create(&ctx, &prm);
        if (rc) die("dbg_create", rc);
        if (attempt == 0) {
            time_end = clock();
            cerr << "Finished! Run time: " << double(time_end - time_start) / CLOCKS_PER_SEC << endl;
            w1 = wall_now();
            cerr << "\nparse input reads files: " << endl;
        }
        bool fits = true;
        Total_reads_num = 0;
        reads_per_file.clear();
        {
            std::vector<std::unique_ptr<dbgio::FileProducer> > prod(reads_files.size());
            size_t started = 0;
            for (size_t i = 0; i < reads_files.size() && fits; i++) {
                for (; started < reads_files.size() && started < i + MAX_AHEAD; started++)
                    prod[started].reset(new dbgio::FileProducer(reads_files[started], Input_file_format, (uint64_t)maxReadLen, BLOCK_BASES,
                                                                BLOCK_READS, dbg_host_alloc, dbg_host_free));
                if (attempt == 0) cerr << "\nStart to parse reads file: " << reads_files[i] << endl;
                const uint64_t reads_before = Total_reads_num;
                fits = consume_file(ctx, *prod[i], attempt > 0);
                reads_per_file.push_back(Total_reads_num - reads_before);
                prod[i].reset();
                if (fits) {
                    rc = dbg_get_stats(ctx, &st);
                    if (rc == DBG_ERR_TABLE_FULL) fits = false;
                    else if (rc) die("dbg_get_stats", rc);
                }
                if (fits) {
                    Kmer_total_num = st.kmers_logged;
                    cerr << "\nTotal number of reads loaded into memory: " << Total_reads_num << endl;
                    cerr << "Total number of kmers loaded into memory: " << Kmer_total_num << endl;
                    time_end = clock();
                    cerr << "Finished! Run time: " << double(time_end - time_start) / CLOCKS_PER_SEC << endl;
                }
            }
        }
        // add polyA and polyT [kmer: 0] to the kmerset (DBGgraph.cpp:418) happens inside dbg_finalize
        w2 = wall_now();
        if (fits) {
            rc = dbg_finalize(ctx, &st);
            if (rc == DBG_ERR_TABLE_FULL) fits = false;
            else if (rc) die("dbg_finalize", rc);
        }
        if (fits) break;
        dbg_destroy(ctx);
        ctx = NULL;
        if (attempt >= 12) { cerr << "libdbgb200: the input does not fit a device table of " << prm.init_slots << " slots" << endl; exit(1); }
        prm.init_slots = prm.init_slots < 1024 ? 2048 : prm.init_slots * 2;
        cerr << "libdbgb200: -i " << initHashSize << " cannot hold this input; rebuilding with a device table of " << prm.init_slots
             << " slots (the CPU program would have enlarged its hash)" << endl;
    }
    // Did the reference grow its table on this input (-e / enlarge, DBGgraph.cpp:337-351)?  Only possible if the final
    // node count passed max_cutoff.  The GPU table never grows; the reference's post-growth slot layout is replayed on the
    // host from the nodes' first-occurrence ordinals (dbg_replay_growth) so that the traversal sees the table it expects.
    dbg_growth_result grow;
    memset(&grow, 0, sizeof(grow));
    std::vector<uint64_t> g_kmer, g_ord;
    std::vector<uint32_t> g_l, g_r;
    float ref_lf = hashLoadFactor;
    if (ref_lf <= 0) ref_lf = 0.25f; else if (ref_lf >= 1) ref_lf = 0.75f;      // kmerSet.cpp:110-111
    const uint64_t ref_max = (uint64_t)(P_slots * ref_lf);
    if (st.count - 1 > ref_max) {
        uint64_t n = 0;
        if ((rc = dbg_dump_shard(ctx, NULL, NULL, NULL, NULL, NULL, &n))) die("dbg_dump_shard", rc);
        g_kmer.resize(n + 1); g_ord.resize(n + 1); g_l.resize(n + 1); g_r.resize(n + 1);
        uint64_t cap = n + 1;
        if ((rc = dbg_dump_shard(ctx, g_kmer.data(), NULL, g_l.data(), g_r.data(), g_ord.data(), &cap))) die("dbg_dump_shard", rc);
        dbg_growth_params gp;
        memset(&gp, 0, sizeof(gp));
        gp.init_slots = ref_init_slots; gp.load_factor = hashLoadFactor; gp.wide = 0;
        gp.max_double_times = maxDoubleHashTimes; gp.buffer_reads = (uint64_t)BufferNum;
        rc = dbg_replay_growth(&gp, reads_per_file.data(), (uint32_t)reads_per_file.size(), g_kmer.data(), NULL, g_l.data(), g_r.data(),
                               g_ord.data(), n, (uint32_t)st.polyA_l, (uint32_t)st.polyA_r, &grow, NULL, NULL);
        if (rc) die("dbg_replay_growth", rc);
        if (grow.truncated)
            cerr << "\nAlert message: Memory reach the maximum allowed in the CPU program (-e " << maxDoubleHashTimes
                 << "): it would have ignored the reads of file " << grow.truncated_file << " from read " << grow.truncated_first_read
                 << " on; this build used all reads. Raise -i or -e.\n" << endl;
        else if (grow.doublings)
            cerr << "\nHash enlarged " << grow.doublings << " time(s) by the CPU program's rule: array size " << grow.final_size << endl;
    }
    const bool grown = grow.doublings > 0 && !grow.truncated;
    doubleHashTimes = grown ? grow.doublings : 0;

    // the KmerSet the traversal consumes (kmerSet.h:88-99, kmerSet.cpp:98-127)
    kset = new KmerSet;
    kset->e_size = sizeof(KmerNode);
    if (!grown && st.array_size != P_slots) {
        // only reachable when the replay refused (-e exhausted) after the device table had to be enlarged
        cerr << "libdbgb200: this input needs more than -i " << initHashSize << " and -e " << maxDoubleHashTimes << " allow; raise -i" << endl;
        exit(1);
    }
    kset->size = grown ? grow.final_size : st.array_size;
    kset->count = st.count;
    kset->count_conflict = st.conflict;
    kset->load_factor = st.load_factor;
    kset->max = grown ? grow.final_max : st.max_cutoff;
    kset->iter_ptr = 0;
    alloc_thread.join();
    if (grown && pinned_array) { dbg_host_free(pinned_array); pinned_array = NULL; }
    kset->array = (pinned_array && P_slots == kset->size) ? pinned_array : (KmerNode *)malloc(kset->size * kset->e_size);
    kset->nul_flag = (uint8_t *)malloc(kset->size / 8 + 1);
    kset->del_flag = (uint8_t *)calloc(kset->size / 8 + 1, 1);
    if (!kset->array || !kset->nul_flag || !kset->del_flag) { cerr << "out of host memory for the kmerset" << endl; exit(1); }
    const double w3 = wall_now();
    if (grown) {
        const uint64_t n = g_kmer.size() - 1;
        dbg_growth_params gp;
        memset(&gp, 0, sizeof(gp));
        gp.init_slots = ref_init_slots; gp.load_factor = hashLoadFactor; gp.wide = 0;
        gp.max_double_times = maxDoubleHashTimes; gp.buffer_reads = (uint64_t)BufferNum;
        rc = dbg_replay_growth(&gp, reads_per_file.data(), (uint32_t)reads_per_file.size(), g_kmer.data(), NULL, g_l.data(), g_r.data(),
                               g_ord.data(), n, (uint32_t)st.polyA_l, (uint32_t)st.polyA_r, &grow, kset->array, kset->nul_flag);
        if (rc) die("dbg_replay_growth", rc);
    } else if ((rc = dbg_export_kmerset(ctx, kset->array, kset->nul_flag))) die("dbg_export_kmerset", rc);
    const double w4 = wall_now();
    cerr << "libdbgb200 wall clock (s): init " << w1 - w0 << ", read files + submit " << w2 - w1 << ", finalize (GPU build + layout) "
         << w3 - w2 << ", export kmerset " << w4 - w3 << endl;

    dbg_destroy(ctx);

    print_kmerset_parameter(kset);
}
