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
// Differences a user can observe (INTEGRATION.md lists them): the "conflict:" statistic counts GPU probe steps, "-t" only
// affects the host traversal, the per-block progress lines come in bursts (the reader runs ahead of them) and the
// "Enlarge hash array size" lines are printed after the file loop, once.  Table growth ("-e", enlarge) never happens on
// the GPU; when the reference would have grown its table the post-growth slot layout is replayed on the host
// (dbg_replay_growth).  When "-e" is exhausted the reference ignores the rest of a file (DBGgraph.cpp:346-350): the
// replay finds that point and the build is redone on exactly the reads the reference used, so the result is again
// the reference's, file by file.
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

// returns false when the device table turned out to be too small for the input (the caller rebuilds with a larger one).
// `limit`: use at most this many reads of the file (the reference stops reading a file when -e is exhausted,
// DBGgraph.cpp:346-350); UINT64_MAX = the whole file.  *cut is set when reads were left unread.
static bool consume_file(dbg_ctx *ctx, dbgio::FileProducer &prod, bool quiet, uint64_t limit, bool *cut)
{
    uint64_t in_block = 0, taken = 0;
    bool full = false, stop = false;
    *cut = false;
    if (!quiet) cerr << "\n" << threadNum << " children threads created!" << endl;       // DBGgraph.cpp:241
    for (;;) {
        dbgio::ReadBlock *b = prod.pop();
        if (prod.failed()) { cerr << "libdbgb200: out of page-locked host memory for the read blocks" << endl; exit(1); }
        uint64_t use = b->n_reads;
        if (taken + use >= limit) { if (taken + use > limit || !b->last) *cut = true; use = limit - taken; stop = true; }
        if (use) {
            int rc = dbg_submit_reads(ctx, b->bases, b->offs, use);
            if (rc == DBG_ERR_TABLE_FULL) full = true;
            else if (rc) die("dbg_submit_reads", rc);
        }
        taken += use;
        if (full) { prod.recycle(b); return false; }      // the producer is cancelled by its owner
        // the per-block lines of the reference (DBGgraph.cpp:241,276,301,323): one set per BufferNum (-b) reads and one
        // at the end of the file
        uint64_t left = use;
        while (in_block + left >= (uint64_t)BufferNum) {
            const uint64_t take = (uint64_t)BufferNum - in_block;
            Total_reads_num += take; left -= take; in_block = 0;
            if (!quiet) {
                cerr << "Load reads block " << Total_reads_num << endl;
                cerr << "chop reads to kmers done" << endl << "add kmers to hash done" << endl;
                if (!(stop && left == 0)) cerr << "\n" << threadNum << " children threads created!" << endl;
            }
        }
        Total_reads_num += left; in_block += left;
        const bool last = b->last;
        prod.recycle(b);
        if (last || stop) break;
    }
    if (!quiet && !*cut) {
        cerr << "Load reads block " << Total_reads_num << endl;
        cerr << "this block has reach the end of file " << endl;
        cerr << "chop reads to kmers done" << endl << "add kmers to hash done" << endl;
    }
    if (!quiet && *cut)                                                                 // DBGgraph.cpp:348
        cerr << "\nAlert message: Memory reach the maximum allowed, program have loaded " << Total_reads_num
             << " reads, the left others are ignored\n" << endl;
    return !full;
}

static double wall_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// the nodes of a finished build with their first-occurrence ordinals + the growth replay (plan only or layout)
struct GrowthInput {
    std::vector<uint64_t> kmer, ord;
    std::vector<uint32_t> l, r;
    uint64_t n = 0;
};

static void fetch_nodes(dbg_ctx *ctx, GrowthInput &g)
{
    int rc;
    uint64_t n = 0;
    if ((rc = dbg_dump_shard(ctx, NULL, NULL, NULL, NULL, NULL, &n))) die("dbg_dump_shard", rc);
    g.kmer.resize(n + 1); g.ord.resize(n + 1); g.l.resize(n + 1); g.r.resize(n + 1);
    uint64_t cap = n + 1;
    if ((rc = dbg_dump_shard(ctx, g.kmer.data(), NULL, g.l.data(), g.r.data(), g.ord.data(), &cap))) die("dbg_dump_shard", rc);
    g.n = n;
}

static void replay(const GrowthInput &g, uint64_t ref_init_slots, const std::vector<uint64_t> &reads_per_file, const dbg_stats &st,
                   dbg_growth_result *grow, void *array, uint8_t *nul_flag)
{
    dbg_growth_params gp;
    memset(&gp, 0, sizeof(gp));
    gp.init_slots = ref_init_slots; gp.load_factor = hashLoadFactor; gp.wide = 0;
    gp.max_double_times = maxDoubleHashTimes; gp.buffer_reads = (uint64_t)BufferNum;
    int rc = dbg_replay_growth(&gp, reads_per_file.data(), (uint32_t)reads_per_file.size(), g.kmer.data(), NULL, g.l.data(), g.r.data(),
                               g.ord.data(), g.n, (uint32_t)st.polyA_l, (uint32_t)st.polyA_r, grow, array, nul_flag);
    if (rc) die("dbg_replay_growth", rc);
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
        if (dbg_host_alloc(&q, (P_slots + 1) * sizeof(KmerNode)) == DBG_OK) pinned_array = (KmerNode *)q;
    });
    cerr << "Hash initialization array size:  " << initHashSize << " G" << endl;
    cerr << "The initialization memory used:  " << initHashSize * 16 << " G" << endl;
    float ref_lf = hashLoadFactor;
    if (ref_lf <= 0) ref_lf = 0.25f; else if (ref_lf >= 1) ref_lf = 0.75f;      // kmerSet.cpp:110-111
    const uint64_t ref_max = (uint64_t)(P_slots * ref_lf);

    // The device table never grows.  If -i turns out too small to even HOLD the nodes (the reference would have
    // enlarged its table, -e), the build is simply redone with a larger device table: the reads are streamed again, a
    // full build takes milliseconds, and the table the traversal gets is laid out by the growth replay further down from
    // the reference's own -i, so the size of the device table never shows.
    // The same loop reproduces "-e exhausted" (DBGgraph.cpp:346-350: the reference stops reading the current file, and
    // every later file after its first block): the replay of the full build tells where the reference stopped, and the
    // build is redone on exactly the reads the reference used (`limits`).
    dbg_ctx *ctx = NULL;
    int rc = 0;
    dbg_stats st;
    std::vector<uint64_t> reads_per_file, limits;
    dbg_growth_result grow;
    GrowthInput gin;
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
                bool cut = false;
                fits = consume_file(ctx, *prod[i], attempt > 0 && limits.empty(), limits.empty() ? UINT64_MAX : limits[i], &cut);
                reads_per_file.push_back(Total_reads_num - reads_before);
                if (!cut) {
                    const std::string io = prod[i]->io_error();
                    if (!io.empty() && (attempt == 0 || !limits.empty())) {
                        cerr << "libdbgb200: WARNING: input problem: " << io << endl;
                        if (getenv("DBG_B200_STRICT_IO")) exit(1);
                    }
                }
                prod[i].reset();
                if (fits) {
                    rc = dbg_get_stats(ctx, &st);
                    if (rc == DBG_ERR_TABLE_FULL) fits = false;
                    else if (rc) die("dbg_get_stats", rc);
                }
                if (fits && (attempt == 0 || !limits.empty())) {
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
        if (!fits) {
            dbg_destroy(ctx);
            ctx = NULL;
            if (attempt >= 24) { cerr << "libdbgb200: the input does not fit a device table of " << prm.init_slots << " slots" << endl; exit(1); }
            prm.init_slots = prm.init_slots < 1024 ? 2048 : prm.init_slots * 2;
            cerr << "libdbgb200: -i " << initHashSize << " cannot hold this input; rebuilding with a device table of " << prm.init_slots
                 << " slots (the CPU program would have enlarged its hash)" << endl;
            continue;
        }
        // Did the reference grow its table on this input (-e / enlarge, DBGgraph.cpp:337-351)?  Only possible if the final
        // node count passed max_cutoff.  The GPU table never grows; the reference's post-growth slot layout is replayed on
        // the host from the nodes' first-occurrence ordinals (dbg_replay_growth) so that the traversal sees the table it
        // expects.
        memset(&grow, 0, sizeof(grow));
        if (st.count - 1 > ref_max) {
            fetch_nodes(ctx, gin);
            replay(gin, ref_init_slots, reads_per_file, st, &grow, NULL, NULL);
            if (grow.truncated) {
                if (!limits.empty()) { cerr << "libdbgb200: internal error: the truncated read set is truncated again" << endl; exit(1); }
                // -e exhausted inside file `truncated_file`: the reference used its reads up to there, and of every later
                // file only the first block of -b reads (a later file's first full block ends with count > max again)
                uint64_t base = 0;
                for (size_t f = 0; f < reads_per_file.size(); f++) {
                    uint64_t lim = reads_per_file[f];
                    if (f == grow.truncated_file) lim = grow.truncated_first_read - base;
                    else if (f > grow.truncated_file && lim > (uint64_t)BufferNum) lim = (uint64_t)BufferNum;
                    limits.push_back(lim);
                    base += reads_per_file[f];
                }
                cerr << "\nlibdbgb200: -e " << maxDoubleHashTimes << " is exhausted inside file " << grow.truncated_file
                     << ": the CPU program ignores the rest of it; rebuilding on exactly the reads it used" << endl;
                dbg_destroy(ctx);
                ctx = NULL;
                prm.init_slots = ref_init_slots;       // the reduced read set starts over from the reference's -i
                continue;
            }
            if (grow.doublings) {
                cerr << "Enlarge hash array size to be: " << grow.final_size << endl;                    // DBGgraph.cpp:343-344
                cerr << "The expanded memory used now:  " << (double)grow.final_size / 1000000000 * 16 << " G" << endl;
                cerr << "\nHash enlarged " << grow.doublings << " time(s) by the CPU program's rule: array size " << grow.final_size << endl;
            }
        }
        break;
    }
    const bool grown = grow.doublings > 0;
    doubleHashTimes = grow.doublings;

    // the KmerSet the traversal consumes (kmerSet.h:88-99, kmerSet.cpp:98-127)
    kset = new KmerSet;
    kset->e_size = sizeof(KmerNode);
    if (!grown && st.array_size != P_slots) {
        // the device table had to be enlarged although the reference's rule never grows: only possible when a final
        // (unchecked) block overfills the table, where the reference itself would probe forever
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
    // one spare, zeroed node behind the table: the traversal reads array[kset->size] when a walk ends without a last node
    // (contig.cpp:320-338 with last_idx == size: out of bounds in the reference, where the bytes behind a freshly mapped
    // table are zero) -- keeps the printed "EndKmer: 0" independent of heap history
    kset->array = (pinned_array && P_slots == kset->size) ? pinned_array : (KmerNode *)malloc((kset->size + 1) * kset->e_size);
    if (kset->array) memset(kset->array + kset->size, 0, sizeof(KmerNode));
    kset->nul_flag = (uint8_t *)malloc(kset->size / 8 + 1);
    kset->del_flag = (uint8_t *)calloc(kset->size / 8 + 1, 1);
    if (!kset->array || !kset->nul_flag || !kset->del_flag) { cerr << "out of host memory for the kmerset" << endl; exit(1); }
    const double w3 = wall_now();
    if (grown) replay(gin, ref_init_slots, reads_per_file, st, &grow, kset->array, kset->nul_flag);
    else if ((rc = dbg_export_kmerset(ctx, kset->array, kset->nul_flag))) die("dbg_export_kmerset", rc);
    const double w4 = wall_now();
    cerr << "libdbgb200 wall clock (s): init " << w1 - w0 << ", read files + submit " << w2 - w1 << ", finalize (GPU build + layout) "
         << w3 - w2 << ", export kmerset " << w4 - w3 << endl;

    dbg_destroy(ctx);

    print_kmerset_parameter(kset);
}
