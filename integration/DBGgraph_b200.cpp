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

// The context outlives build_debruijn_graph() when the table the traversal got IS the device image (no growth replay):
// calculate_kmer_links (below) then takes the link pass from the GPU instead of scanning P slots.
static dbg_ctx *g_links_ctx = NULL;

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

// The build runs on one GPU (dbg_*) or, with DBG_B200_GPUS=n in the environment (DBG_B200_DEVICES=0,1,... picks them), on
// several from this one process (dbg_mg_*: reads dealt to the GPUs, k-mers exchanged over NVLink by owner slot range, the
// table slices merged back into the one KmerSet the traversal consumes).  Same calls, same results.
struct Build {
    dbg_ctx *one = NULL;
    dbg_mg *many = NULL;
    int submit(const char *bases, const uint64_t *offs, uint64_t n) { return many ? dbg_mg_submit_reads(many, bases, offs, n) : dbg_submit_reads(one, bases, offs, n); }
    int get_stats(dbg_stats *st) { return many ? dbg_mg_get_stats(many, st) : dbg_get_stats(one, st); }
    int finalize(dbg_stats *st) { return many ? dbg_mg_finalize(many, st) : dbg_finalize(one, st); }
    int dump(uint64_t *k, uint32_t *l, uint32_t *r, uint64_t *o, uint64_t *n) { return many ? dbg_mg_dump_nodes(many, k, NULL, l, r, o, n) : dbg_dump_shard(one, k, NULL, l, r, o, n); }
    int export_kmerset(void *array, uint8_t *nul) { return many ? dbg_mg_export_kmerset(many, array, nul) : dbg_export_kmerset(one, array, nul); }
    void destroy() { if (many) dbg_mg_destroy(many); if (one) dbg_destroy(one); many = NULL; one = NULL; }
    const char *err() { return many && dbg_mg_last_error()[0] ? dbg_mg_last_error() : dbg_last_error(); }
};

// returns false when the device table turned out to be too small for the input (the caller rebuilds with a larger one).
// `limit`: use at most this many reads of the file (the reference stops reading a file when -e is exhausted,
// DBGgraph.cpp:346-350); UINT64_MAX = the whole file.  *cut is set when reads were left unread.
static bool consume_file(Build &ctx, dbgio::FileProducer &prod, bool quiet, uint64_t limit, bool *cut)
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
            int rc = ctx.submit(b->bases, b->offs, use);
            if (rc == DBG_ERR_TABLE_FULL) full = true;
            else if (rc) { cerr << ctx.err() << endl; die("dbg_submit_reads", rc); }
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

static void fetch_nodes(Build &ctx, GrowthInput &g)
{
    int rc;
    uint64_t n = 0;
    if ((rc = ctx.dump(NULL, NULL, NULL, NULL, &n))) die("dbg_dump_shard", rc);
    g.kmer.resize(n + 1); g.ord.resize(n + 1); g.l.resize(n + 1); g.r.resize(n + 1);
    uint64_t cap = n + 1;
    if ((rc = ctx.dump(g.kmer.data(), g.l.data(), g.r.data(), g.ord.data(), &cap))) die("dbg_dump_shard", rc);
    g.n = cap;
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

// DBG_B200_CHECKPOINT=<file>: the finished KmerSet is kept on disk (dbg_checkpoint_write); a later run with the same -k
// finds it and goes straight to the traversal -- other cut-offs (-D -T -I -P -W -C -G -B -U -L -E -M) without reading
// the reads again.  The reference has no such restart point.
static bool load_checkpoint(const char *path)
{
    dbg_checkpoint_header h;
    if (dbg_checkpoint_read_header(path, &h) != DBG_OK) return false;
    if ((int)h.K != KmerSize || h.wide) { cerr << "libdbgb200: checkpoint " << path << " was built with -k " << h.K << ": ignored" << endl; return false; }
    kset = new KmerSet;
    kset->e_size = sizeof(KmerNode);
    kset->size = h.size; kset->count = h.count; kset->count_conflict = h.count_conflict; kset->load_factor = h.load_factor;
    kset->max = h.max_cutoff; kset->iter_ptr = 0;
    kset->array = (KmerNode *)malloc((kset->size + 1) * kset->e_size);
    kset->nul_flag = (uint8_t *)malloc(kset->size / 8 + 1);
    kset->del_flag = (uint8_t *)calloc(kset->size / 8 + 1, 1);
    if (!kset->array || !kset->nul_flag || !kset->del_flag) { cerr << "out of host memory for the kmerset" << endl; exit(1); }
    memset(kset->array + kset->size, 0, sizeof(KmerNode));
    if (dbg_checkpoint_read(path, kset->array, kset->nul_flag) != DBG_OK) { cerr << "libdbgb200: checkpoint " << path << " is damaged: rebuilding" << endl; return false; }
    Total_reads_num = h.reads; Kmer_total_num = h.kmers_logged;
    cerr << "libdbgb200: graph of " << h.count << " nodes loaded from checkpoint " << path << " (no reads were parsed)" << endl;
    return true;
}

static void save_checkpoint(const char *path)
{
    dbg_checkpoint_header h;
    memset(&h, 0, sizeof(h));
    h.K = (uint32_t)KmerSize; h.wide = 0; h.load_factor = kset->load_factor; h.size = kset->size; h.max_cutoff = kset->max; h.count = kset->count;
    h.count_conflict = kset->count_conflict; h.reads = Total_reads_num; h.kmers_logged = Kmer_total_num;
    if (dbg_checkpoint_write(path, &h, kset->array, kset->nul_flag) != DBG_OK) cerr << "libdbgb200: WARNING: could not write checkpoint " << path << endl;
    else cerr << "libdbgb200: checkpoint written to " << path << endl;
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

    const char *ckpt = getenv("DBG_B200_CHECKPOINT");
    if (ckpt && *ckpt && load_checkpoint(ckpt)) { print_kmerset_parameter(kset); return; }

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
    Build ctx;
    int n_gpus = getenv("DBG_B200_GPUS") ? atoi(getenv("DBG_B200_GPUS")) : 1;
    std::vector<int32_t> gpu_list;
    if (const char *e = getenv("DBG_B200_DEVICES")) {
        for (const char *q = e; *q;) { gpu_list.push_back((int32_t)strtol(q, (char **)&q, 10)); if (*q == ',') q++; else break; }
        if (!getenv("DBG_B200_GPUS")) n_gpus = (int)gpu_list.size();
    }
    if (n_gpus < 1) n_gpus = 1;
    if (!gpu_list.empty() && (int)gpu_list.size() != n_gpus) { cerr << "libdbgb200: DBG_B200_DEVICES lists " << gpu_list.size() << " devices, DBG_B200_GPUS=" << n_gpus << endl; exit(1); }
    if (n_gpus > 1) cerr << "libdbgb200: building on " << n_gpus << " GPUs" << endl;
    int rc = 0;
    dbg_stats st;
    std::vector<uint64_t> reads_per_file, limits;
    dbg_growth_result grow;
    GrowthInput gin;
    double w1 = 0, w2 = 0;
    for (int attempt = 0;; attempt++) {
        if (n_gpus > 1) { rc = dbg_mg_create(&ctx.many, &prm, n_gpus, gpu_list.empty() ? NULL : gpu_list.data()); if (rc) cerr << dbg_mg_last_error() << endl; }
        else rc = dbg_create(&ctx.one, &prm);
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
                    rc = ctx.get_stats(&st);
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
            rc = ctx.finalize(&st);
            if (rc == DBG_ERR_TABLE_FULL) fits = false;
            else if (rc) { cerr << ctx.err() << endl; die("dbg_finalize", rc); }
        }
        if (!fits) {
            ctx.destroy();
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
                ctx.destroy();
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
    else if ((rc = ctx.export_kmerset(kset->array, kset->nul_flag))) { cerr << ctx.err() << endl; die("dbg_export_kmerset", rc); }
    const double w4 = wall_now();
    if (ckpt && *ckpt) save_checkpoint(ckpt);
    cerr << "libdbgb200 wall clock (s): init " << w1 - w0 << ", read files + submit " << w2 - w1 << ", finalize (GPU build + layout) "
         << w3 - w2 << ", export kmerset " << w4 - w3 << endl;

#ifdef DBG_B200_GPU_LINKS
    if (!grown && ctx.one && !getenv("DBG_B200_CPU_LINKS")) { g_links_ctx = ctx.one; ctx.one = NULL; }   // calculate_kmer_links() below finishes with it
#endif
    ctx.destroy();

    print_kmerset_parameter(kset);
}

#ifdef DBG_B200_GPU_LINKS
// ---- first pass of the traversal on the GPU (contig.cpp:107-205) ----------------------------------------------------
// Compiled in when the build recipe (oracle/Makefile, target b200) renamed contig.cpp's own calculate_kmer_links to
// calculate_kmer_links_cpu in a generated copy: build_contig_sequence() (contig.cpp:62) then calls THIS function.  klink,
// del_flag, the depth histogram and the slot-ordered tip / branch lists come from dbg_export_links (k_links_classify +
// ordered compaction on the table image that is still resident on the device) instead of a host scan over P slots; the
// .contig.kmer.freq file and the log lines are written exactly like the original.  When the table was laid out by the
// host-side growth replay the device image is not the table the traversal holds: the original runs.
#include "contig.h"
#include <fstream>
void calculate_kmer_links_cpu(KmerSet *kset, KmerLink *klink, vector<uint64_t> &tip_nodes, vector<uint64_t> &branch_nodes);

void calculate_kmer_links(KmerSet *kset, KmerLink *klink, vector<uint64_t> &tip_nodes, vector<uint64_t> &branch_nodes)
{
    if (!g_links_ctx) { calculate_kmer_links_cpu(kset, klink, tip_nodes, branch_nodes); return; }
    static_assert(sizeof(KmerLink) == 2, "KmerLink is the 2-byte bit-field struct of contig.h:31-42");
    int64_t DepthStat[256], stats3[3] = {0, 0, 0};
    uint64_t nt = 0, nb = 0;
    int rc = dbg_export_links(g_links_ctx, KmerFreqCutoff, (uint8_t *)klink, kset->del_flag, DepthStat, NULL, &nt, NULL, &nb, stats3);
    if (rc) die("dbg_export_links", rc);
    tip_nodes.resize(nt); branch_nodes.resize(nb);
    uint64_t ct = nt ? nt : 1, cb = nb ? nb : 1;
    uint64_t dummy = 0;
    rc = dbg_export_links(g_links_ctx, KmerFreqCutoff, NULL, NULL, NULL, nt ? tip_nodes.data() : &dummy, &ct, nb ? branch_nodes.data() : &dummy, &cb, NULL);
    if (rc) die("dbg_export_links (lists)", rc);
    dbg_destroy(g_links_ctx);
    g_links_ctx = NULL;
    const int64_t total_kmer_speceis_num = stats3[0], deleted_lowFreq_kmer_num = stats3[1], linear_kmer_node_num = stats3[2];

    string kmer_depth_file = Output_prefix + ".contig.kmer.freq";                       // contig.cpp:186-203
    ofstream depthFile(kmer_depth_file.c_str());
    if (!depthFile) cerr << "fail to open file " << kmer_depth_file << endl;
    cerr << "\nTotal kmer nodes number:    " << total_kmer_speceis_num << endl;
    cerr << "Deleted lowfreq kmer nodes: " << deleted_lowFreq_kmer_num << "\t" << (double)deleted_lowFreq_kmer_num / total_kmer_speceis_num << endl;
    cerr << "Used linear kmer nodes:     " << linear_kmer_node_num << "\t" << (double)linear_kmer_node_num / total_kmer_speceis_num << endl;
    cerr << "Used tip kmer nodes:        " << tip_nodes.size() << "\t" << (double)tip_nodes.size() / total_kmer_speceis_num << endl;
    cerr << "Used branching kmer nodes:  " << branch_nodes.size() << "\t" << (double)branch_nodes.size() / total_kmer_speceis_num << endl;
    depthFile << "Kmer_depth\tAppear_times\n";
    for (int i = 1; i <= 255; i++) depthFile << i << "\t" << DepthStat[i] << endl;
    depthFile.close();
}
#endif
