// dbg_build.cu -- host side of libdbgb200.so: context, streams, staging and the extern "C" entry
// points declared in include/dbg_b200.h.  The kernels live in dbg_kernels.cuh.
//
// There is no CPU fallback in this file: every entry point that computes launches sm_100a kernels, and
// dbg_create fails with DBG_ERR_CUDA when no device is available.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dbg_b200.h"
#include "dbg_kernels.cuh"
#include "export_pipe.h"

using namespace dbg;

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? DBG_ERR_NOMEM : DBG_ERR_CUDA;          \
            return set_err(code_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
        }                                                                                         \
    } while (0)

extern "C" const char *dbg_strerror(int code)
{
    switch (code) {
    case DBG_OK: return "ok";
    case DBG_ERR_INVALID: return "invalid argument";
    case DBG_ERR_CUDA: return "CUDA error";
    case DBG_ERR_NOMEM: return "out of memory";
    case DBG_ERR_TABLE_FULL: return "k-mer table full (the reference would enlarge; raise -i)";
    case DBG_ERR_STATE: return "call order violated";
    case DBG_ERR_BUFFER: return "output buffer too small";
    default: return "unknown error";
    }
}

extern "C" const char *dbg_last_error(void) { return g_err; }

extern "C" int dbg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------------------------
// scalar helpers, bit-identical to the reference
// ---------------------------------------------------------------------------------------------------
// kmerSet.cpp:72-82: trial division up to (uint64)sqrt((float)num), EXCLUSIVE -- float rounding and the
// `<` make some composites pass; table sizes must match the reference, so the quirk is kept.
static int ref_is_prime(uint64_t num)
{
    if (num < 4) return 1;
    if (num % 2 == 0) return 0;
    uint64_t max = (uint64_t)sqrtf((float)num);
    for (uint64_t i = 3; i < max; i += 2)
        if (num % i == 0) return 0;
    return 1;
}

extern "C" uint64_t dbg_find_next_prime(uint64_t num)   // kmerSet.cpp:86-95
{
    if (num % 2 == 0) num++;
    while (!ref_is_prime(num)) num += 2;
    return num;
}

extern "C" uint64_t dbg_hash_code(uint64_t kmer) { return hash_code(kmer); }
extern "C" uint64_t dbg_hash_code_wide(uint64_t lo, uint64_t hi) { return hash_code_wide(lo, hi); }

extern "C" int dbg_host_alloc(void **p, uint64_t bytes)
{
    if (!p) return set_err(DBG_ERR_INVALID, "dbg_host_alloc: NULL");
    CU_TRY(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return DBG_OK;
}

// page-lock memory the caller allocated itself (e.g. a shared-memory table image several ranks export into)
extern "C" int dbg_host_register(void *p, uint64_t bytes)
{
    if (!p) return set_err(DBG_ERR_INVALID, "dbg_host_register: NULL");
    CU_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return DBG_OK;
}

extern "C" int dbg_host_unregister(void *p)
{
    if (p) CU_TRY(cudaHostUnregister(p));
    return DBG_OK;
}

extern "C" int dbg_host_free(void *p)
{
    if (p) CU_TRY(cudaFreeHost(p));
    return DBG_OK;
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
// host submit is cut into sub-blocks so copies overlap kernels (env DBG_B200_SUB_BASES / _SUB_READS override, for tests)
static const uint64_t SUB_BASES_DEFAULT = 64ull << 20;
static const uint64_t SUB_READS_DEFAULT = 2ull << 20;
static const uint64_t MARGIN_SLOTS = 1ull << 16; // overflow zone past a shard's home range (no wrap in the hot loop)

struct EvPair { cudaEvent_t a, b; int slot; };   // slot: which ms[] entry the elapsed time is added to

struct dbg_ctx {
    dbg_params prm;
    bool wide, track;
    int device, n_sms;
    uint64_t P, M, max_cutoff;
    float lf;
    uint64_t shard_lo, shard_hi, shard_size, n_local;
    int n_shards;
    cudaStream_t stream, copy_stream, own_stream;
    void *d_nodes;                 // n_local x NodeT<wide>: local slot 0 (behind the porch)
    void *d_nodes_alloc = nullptr; // the allocation: [porch | n_local]
    uint64_t porch = 0;            // sharded contexts: slots in front of local slot 0 that receive the left neighbour's tail cluster
    TailInfo *d_tail = nullptr;    // cross-shard hand-off scratch
    void *d_tail_nodes = nullptr;  // staging of imported margin nodes
    uint64_t tail_nodes_cap = 0;
    uint64_t tail_a_own = 0, tail_mt_own = 0, tail_a_in = 0;
    bool tail_exported = false, tail_imported = false, tail_moved = false;
    uint64_t undo_reads = 0;       // reads counted by the last optimistic exchange scatter (dbg_exchange_scatter_undo)
    uint64_t guard_total = 0;      // occurrences (upper bound) submitted to the inserts so far: sizes the long-probe budget
    int64_t nodes_delta = 0;       // nodes handed to / adopted from the neighbours (dump_shard counts what is physically here)
    u32 *d_nul_slice = nullptr;    // occupancy words of the laid-out slice (global word positions)
    uint64_t nul_slice_words = 0;
    u64 *d_counters, *d_polyA;
    // host-submit staging (double buffered)
    char *d_bases[2];
    u64 *d_offs[2];
    uint64_t cap_bases, cap_reads, sub_bases, sub_reads;
    cudaEvent_t ev_free[2];
    int cur;
    u64 *d_offs_stage;
    uint64_t batch_bases, batch_reads, batch_read_index0;
    u64 *d_chunk_first;
    uint64_t cap_chunks;
    // radix-partitioned build
    int part_mode;                 // 0 never, 1 always, 2 auto
    int part_shift;                // bucket = local slot >> part_shift
    uint32_t n_buckets;
    u64 *d_boffs, *d_roffs, *d_tuples, *d_tile_sums;   // d_roffs: owner-rank offsets of the exchange (65 entries)
    u32 *d_matrix;
    uint64_t cap_tuples, cap_matrix, part_blocks;
    // finalize / export
    bool finalized;
    u64 *d_owner;                  // layout scratch (regions) or owner[P] (global method)
    uint64_t owner_cap;
    LayoutInfo *d_layout_info;
    LayoutRegion *d_regions;
    int layout_v;                  // cluster-local pass: 2 (default) or 1 (env DBG_B200_LAYOUT_V)
    int layout_mode;               // 0 cluster-local (default), 1 global atomicMin method (env DBG_B200_LAYOUT=global)
    int pipeline;                  // env DBG_B200_PIPELINE (default 1): a large dbg_submit_reads call scatters sub-block i while sub-block i+1 is copied
    int optimistic;                // env DBG_B200_OPTIMISTIC (default 1): single-pass partition with fixed bucket regions first
    int opt_capb;                  // env DBG_B200_OPT_CAPB: force the region size (tests: provoke the overflow fallback)
    u32 *d_fill = nullptr;         // [n_buckets] tuples per bucket + [n_buckets] overflow flag
    u64 *d_snap = nullptr;         // counters + polyA before an optimistic scatter (restored on overflow)
    u32 *h_flag = nullptr;         // pinned
    u64 *h_cnt = nullptr;          // pinned copy of the counters taken after every host batch (prompt table-full report)
    cudaEvent_t ev_cnt = nullptr;
    bool cnt_pending = false;
    uint64_t path_counts[4] = {0, 0, 0, 0};   // blocks: direct, partitioned exact, partitioned optimistic, overflow fallbacks
    int stage_cap;                 // env DBG_B200_STAGE_CAP: batch size of the staged scatter (-1 default, 0 off)
    int peer_unstaged;             // env DBG_B200_PEER_UNSTAGED=1: fused exchange stores tuples one by one (experiments)
    uint32_t layout_regions;
    void *d_out;
    u32 *d_nul32;
    u64 polyA_links;
    // links scratch
    unsigned short *d_klink;
    u32 *d_del32, *d_tile_counts;
    u64 *d_tile_offs, *d_small;   // d_small: hist[256] + stats3[3] + totals[4]
    int links_cutoff;             // cutoff the scratch currently reflects (INT32_MIN = none)
    // bookkeeping
    uint64_t reads_total, next_read_index, launches;
    dbg_stats st;
    std::vector<EvPair> build_ev;
    std::vector<cudaEvent_t> ev_all, ev_free_list;   // every timing event this context ever created / the idle ones
    float ms[8];
    // dbg_reset clears the table on a side stream and returns; the first KERNEL of the next build waits for it (table_ready), so
    // the 6.4-GB memset of C2 runs beside the host->device copy of the first reads instead of in front of it
    cudaStream_t clear_stream = nullptr;
    cudaEvent_t ev_clear = nullptr;
    bool clear_pending = false;
    EvPair clear_ev{nullptr, nullptr, 0};
    bool clear_timed = false;
    LayoutInfo *d_layout_info2 = nullptr;      // dbg_finish_export: the wrap-around region found after the grouped layout
    LayoutRegion *d_regions2 = nullptr;
    LayoutRegion *h_regions = nullptr;         // pinned copy of the region list (patch copies)
    ExportPipe *pipe = nullptr;    // pipelined export (export_pipe.cu): pinned ring, compact buffer, host worker threads
    uint64_t export_info[4] = {0, 0, 0, 0};   // last dbg_export_kmerset: chunks sent compact, chunks sent plain, bytes over the link, nodes
};

static int node_bytes(const dbg_ctx *c) { return c->wide ? 32 : 16; }            // export image
static size_t build_node_bytes(const dbg_ctx *c) { return c->wide ? sizeof(NodeT<true>) : sizeof(NodeT<false>); }

static TableView view_of(dbg_ctx *c)
{
    TableView t;
    t.nodes = c->d_nodes; t.P = c->P; t.M = c->M; t.lo = c->shard_lo; t.n_local = c->n_local;
    t.counters = c->d_counters; t.polyA = c->d_polyA;
    t.guard_budget = (1ull << 12) + c->guard_total / 16;      // in units of 16384 probe steps (insert_probe)
    return t;
}

// timing events come from a per-context pool: an error return between ev_begin and ev_put loses nothing (the
// events stay owned by the context and die with it)
static int ev_get(dbg_ctx *c, cudaEvent_t *e)
{
    if (!c->ev_free_list.empty()) { *e = c->ev_free_list.back(); c->ev_free_list.pop_back(); return DBG_OK; }
    CU_TRY(cudaEventCreate(e));
    c->ev_all.push_back(*e);
    return DBG_OK;
}

static void ev_put(dbg_ctx *c, EvPair &p)
{
    c->ev_free_list.push_back(p.a); c->ev_free_list.push_back(p.b);
}

static int ev_begin(dbg_ctx *c, cudaStream_t s, EvPair *p)
{
    p->slot = 1;
    int rc = ev_get(c, &p->a);
    if (rc) return rc;
    if ((rc = ev_get(c, &p->b))) { c->ev_free_list.push_back(p->a); return rc; }
    CU_TRY(cudaEventRecord(p->a, s));
    return DBG_OK;
}

static void free_finalize_buffers(dbg_ctx *c)
{
    cudaFree(c->d_owner); cudaFree(c->d_out); cudaFree(c->d_nul32); cudaFree(c->d_layout_info); cudaFree(c->d_regions);
    cudaFree(c->d_layout_info2); cudaFree(c->d_regions2); if (c->h_regions) cudaFreeHost(c->h_regions);
    c->d_layout_info2 = nullptr; c->d_regions2 = nullptr; c->h_regions = nullptr;
    c->owner_cap = 0; c->d_layout_info = nullptr; c->d_regions = nullptr;
    cudaFree(c->d_klink); cudaFree(c->d_del32); cudaFree(c->d_tile_counts); cudaFree(c->d_tile_offs); cudaFree(c->d_small);
    c->d_owner = nullptr; c->d_out = nullptr; c->d_nul32 = nullptr;
    c->d_klink = nullptr; c->d_del32 = nullptr; c->d_tile_counts = nullptr; c->d_tile_offs = nullptr; c->d_small = nullptr;
    c->links_cutoff = INT32_MIN;
}

extern "C" void dbg_destroy(dbg_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    export_pipe_destroy(c->pipe);
    free_finalize_buffers(c);
    for (cudaEvent_t e : c->ev_all) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_bases[i]); cudaFree(c->d_offs[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    cudaFree(c->d_chunk_first); cudaFree(c->d_nodes_alloc); cudaFree(c->d_counters); cudaFree(c->d_polyA);
    cudaFree(c->d_tail); cudaFree(c->d_tail_nodes); cudaFree(c->d_nul_slice);
    cudaFree(c->d_offs_stage); cudaFree(c->d_boffs); cudaFree(c->d_roffs); cudaFree(c->d_tuples); cudaFree(c->d_matrix); cudaFree(c->d_tile_sums);
    if (c->clear_stream) cudaStreamDestroy(c->clear_stream);
    if (c->ev_clear) cudaEventDestroy(c->ev_clear);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaFree(c->d_fill); cudaFree(c->d_snap); if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    if (c->ev_cnt) cudaEventDestroy(c->ev_cnt);
    delete c;
}

// the table clear was started on the side stream by dbg_reset: `s` -- the stream about to touch the table -- waits for it
static int table_ready(dbg_ctx *c, cudaStream_t s)
{
    if (!c->clear_pending) return DBG_OK;
    CU_TRY(cudaStreamWaitEvent(s, c->ev_clear, 0));
    if (s != c->stream) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_clear, 0));      // later table work on the context's stream is ordered too
    c->clear_pending = false;
    return DBG_OK;
}

static int clear_table(dbg_ctx *c, bool async = false)
{
    if (async && getenv("DBG_B200_SYNC_CLEAR")) async = false;
    if (c->clear_timed) {          // the previous clear's time was never collected (reset twice in a row)
        if (cudaEventSynchronize(c->clear_ev.b) == cudaSuccess) ev_put(c, c->clear_ev);
        c->clear_timed = false;
    }
    if (!c->clear_stream) {
        CU_TRY(cudaStreamCreateWithFlags(&c->clear_stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_clear, cudaEventDisableTiming));
    }
    cudaStream_t cs = async ? c->clear_stream : c->stream;
    EvPair e;
    int rc = ev_begin(c, cs, &e);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(c->d_nodes_alloc, 0, (c->porch + c->n_local) * build_node_bytes(c), cs));
    CU_TRY(cudaEventRecord(e.b, cs));
    // the side counters are read and written by the extraction kernels: they are cleared on the context's own stream
    CU_TRY(cudaMemsetAsync(c->d_counters, 0, CNT_N * sizeof(u64), c->stream));
    CU_TRY(cudaMemsetAsync(c->d_polyA, 0, 8 * sizeof(u64), c->stream));
    if (async) {
        CU_TRY(cudaEventRecord(c->ev_clear, cs));
        c->clear_pending = true;
        c->clear_ev = e; c->clear_timed = true;      // elapsed time collected by collect_clear_time (dbg_finalize / dbg_get_timings)
        c->ms[0] = 0;
        return DBG_OK;
    }
    CU_TRY(cudaEventSynchronize(e.b));
    CU_TRY(cudaEventElapsedTime(&c->ms[0], e.a, e.b));
    ev_put(c, e);
    c->clear_pending = false;
    return DBG_OK;
}

static void collect_clear_time(dbg_ctx *c)
{
    if (!c->clear_timed) return;
    if (cudaEventSynchronize(c->clear_ev.b) == cudaSuccess) {
        float t = 0;
        if (cudaEventElapsedTime(&t, c->clear_ev.a, c->clear_ev.b) == cudaSuccess) c->ms[0] = t;
    }
    cudaGetLastError();
    ev_put(c, c->clear_ev);
    c->clear_timed = false;
}

extern "C" int dbg_create(dbg_ctx **out, const dbg_params *p)
{
    if (!out || !p) return set_err(DBG_ERR_INVALID, "dbg_create: NULL argument");
    *out = nullptr;
    if (p->K < 1 || p->K > 63) return set_err(DBG_ERR_INVALID, "K=%d outside 1..63", p->K);
    if (p->max_read_len < 1 || p->max_read_len > 65535) return set_err(DBG_ERR_INVALID, "max_read_len=%d outside 1..65535", p->max_read_len);
    int n_shards = p->shard_count > 1 ? p->shard_count : 1;
    if (p->payload_mode != 0 && p->payload_mode != 1) return set_err(DBG_ERR_INVALID, "payload_mode=%d", p->payload_mode);
    if (p->payload_mode == 1 && (p->K > 31 || p->force_wide || n_shards > 1 || !p->track_order))
        return set_err(DBG_ERR_INVALID, "the seed index needs K <= 31, one shard and track_order");
    if (p->shard_rank < 0 || p->shard_rank >= n_shards) return set_err(DBG_ERR_INVALID, "shard_rank %d / %d", p->shard_rank, n_shards);
    int ndev = dbg_device_count();
    if (ndev == 0) return set_err(DBG_ERR_CUDA, "no CUDA device visible: libdbgb200 has no CPU fallback");
    if (p->device < 0 || p->device >= ndev) return set_err(DBG_ERR_INVALID, "device %d of %d", p->device, ndev);
    CU_TRY(cudaSetDevice(p->device));
    int n_sms = 148;
    CU_TRY(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, p->device));

    dbg_ctx *c = new dbg_ctx();
    c->n_sms = n_sms;
    memset(&c->st, 0, sizeof(c->st));
    c->prm = *p;
    c->wide = p->K > 31 || p->force_wide;
    c->track = p->track_order != 0;
    c->device = p->device;
    // init_kmerset_parallel, kmerSet.cpp:98-115
    uint64_t size = p->init_slots < 3 ? 3 : dbg_find_next_prime(p->init_slots);
    float lf = p->load_factor;
    if (lf <= 0) lf = 0.25f; else if (lf >= 1) lf = 0.75f;
    c->P = size; c->lf = lf;
    c->max_cutoff = (uint64_t)(size * lf);            // uint64 * float -> float, as in the reference
    c->M = (uint64_t)((((unsigned __int128)1) << 64) / size);
    c->n_shards = n_shards;
    c->shard_size = (size + n_shards - 1) / n_shards;
    c->shard_lo = (uint64_t)p->shard_rank * c->shard_size;
    if (c->shard_lo > size) c->shard_lo = size;
    c->shard_hi = c->shard_lo + c->shard_size < size ? c->shard_lo + c->shard_size : size;
    c->n_local = (c->shard_hi - c->shard_lo) + MARGIN_SLOTS;
    // porch for the cross-shard hand-off of boundary clusters (see k_shard_tail_scan); window [porch | own range] must
    // stay shorter than the ring of P slots so that home -> virtual slot is unambiguous
    c->porch = n_shards > 1 ? (MARGIN_SLOTS < c->shard_size / 2 ? MARGIN_SLOTS : c->shard_size / 2) : 0;
    c->links_cutoff = INT32_MIN;
    c->sub_bases = SUB_BASES_DEFAULT; c->sub_reads = SUB_READS_DEFAULT;
    if (const char *e = getenv("DBG_B200_SUB_BASES")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 16) c->sub_bases = v; }
    if (const char *e = getenv("DBG_B200_SUB_READS")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1) c->sub_reads = v; }
    c->cap_bases = 1ull << 30; c->cap_reads = 16ull << 20;      // one host batch (x2 buffers)
    if (const char *e = getenv("DBG_B200_BATCH_BASES")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 16) c->cap_bases = v; }
    if (const char *e = getenv("DBG_B200_BATCH_READS")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1) c->cap_reads = v; }
    if (c->sub_bases > c->cap_bases) c->sub_bases = c->cap_bases;
    if (c->sub_reads > c->cap_reads) c->sub_reads = c->cap_reads;
    c->pipeline = getenv("DBG_B200_PIPELINE") ? atoi(getenv("DBG_B200_PIPELINE")) : 1;
    c->optimistic = getenv("DBG_B200_OPTIMISTIC") ? atoi(getenv("DBG_B200_OPTIMISTIC")) : 1;
    c->opt_capb = getenv("DBG_B200_OPT_CAPB") ? atoi(getenv("DBG_B200_OPT_CAPB")) : 0;
    c->stage_cap = getenv("DBG_B200_STAGE_CAP") ? atoi(getenv("DBG_B200_STAGE_CAP")) : -1;
    c->peer_unstaged = getenv("DBG_B200_PEER_UNSTAGED") ? atoi(getenv("DBG_B200_PEER_UNSTAGED")) : 0;
    // version 2 of the cluster-local layout pass for 32-B nodes (C2: 4.61 -> 3.67 ms); 64-B nodes keep version 1, which
    // measured faster there (C3, K=63, table load 0.37: 13.3 vs 17.5 ms -- three 16-B quarters per staged slot and fewer
    // occupied slots per warp-owned word pair)
    c->layout_v = c->wide ? 1 : 2;
    if (const char *e = getenv("DBG_B200_LAYOUT_V")) { int v = atoi(e); if (v == 1 || v == 2) c->layout_v = v; }
    c->layout_mode = 0;
    if (const char *e = getenv("DBG_B200_LAYOUT")) c->layout_mode = strcmp(e, "global") == 0 ? 1 : 0;
    c->part_mode = 2;
    if (const char *e = getenv("DBG_B200_PARTITION")) c->part_mode = atoi(e) == 0 ? 0 : (atoi(e) == 1 ? 1 : 2);
    // one bucket = a table slice of 16 MB (2^19 nodes of 32 B, 2^18 of 64 B): the slice in use, the one being
    // prefetched and the tuple stream fit the 126 MB L2 with room to spare
    c->part_shift = c->wide ? 18 : 19;
    // sharded tables: 32-MB slices (measured neutral for the insert: 7.62 vs 7.58 ms on C2, 7.82 with 64-MB slices), which
    // halves the (owner, slice) buckets of the pull exchange -- 382 on 2 GPUs, 764 on 4, 1528 on 8: up to 2048 the source-side
    // partition holds its reservations in registers and keeps three CTAs per SM, and its store runs get longer
    // (64-B nodes too: C3 on 2 GPUs has 2 x 2289 slices of 16 MB -- beyond the 4096 buckets of the pull exchange, i.e. the
    // push exchange with its owner-side partition pass, 68.3 ms per step -- and 2 x 1145 slices of 32 MB: pull, 59.3 ms)
    if (n_shards >= 2) c->part_shift++;
    if (const char *e = getenv("DBG_B200_PART_SHIFT")) { int v = atoi(e); if (v >= 4 && v <= 40) c->part_shift = v; }
    // slice geometry from the (rank-independent) shard size, so that every rank of a sharded build agrees on it
    // (the last shard may be smaller: its trailing slices just stay empty)
    {
        const uint64_t uni = c->shard_size + MARGIN_SLOTS;
        while (((uni + (1ull << c->part_shift) - 1) >> c->part_shift) > 4096) c->part_shift++;
        c->n_buckets = (uint32_t)((uni + (1ull << c->part_shift) - 1) >> c->part_shift);
    }
    *out = c;   // from here on the caller can dbg_destroy() after a failure

    CU_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU_TRY(cudaMalloc(&c->d_nodes_alloc, (c->porch + c->n_local) * build_node_bytes(c)));
    c->d_nodes = static_cast<char *>(c->d_nodes_alloc) + c->porch * build_node_bytes(c);
    CU_TRY(cudaMalloc(&c->d_counters, CNT_N * sizeof(u64)));
    CU_TRY(cudaMalloc(&c->d_polyA, 8 * sizeof(u64)));
    CU_TRY(cudaMalloc(&c->d_boffs, ((size_t)c->n_buckets + 1) * sizeof(u64)));
    CU_TRY(cudaMalloc(&c->d_roffs, (64 * 1024 + 1) * sizeof(u64)));     // owner (x slice) offsets of the exchange
    return clear_table(c);
}

extern "C" int dbg_set_stream(dbg_ctx *c, void *stream)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaDeviceSynchronize());
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return DBG_OK;
}

extern "C" int dbg_reset(dbg_ctx *c)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaDeviceSynchronize());
    for (auto &e : c->build_ev) ev_put(c, e);
    c->build_ev.clear();
    c->finalized = false; c->reads_total = 0; c->next_read_index = 0; c->polyA_links = 0;
    c->batch_reads = 0; c->batch_bases = 0; c->part_blocks = 0;
    c->cnt_pending = false; c->guard_total = 0;
    c->tail_exported = c->tail_imported = c->tail_moved = false; c->tail_a_own = c->tail_mt_own = c->tail_a_in = 0; c->nodes_delta = 0;
    for (int i = 0; i < 4; i++) c->path_counts[i] = 0;
    c->links_cutoff = INT32_MIN;
    for (int i = 1; i < 8; i++) c->ms[i] = 0;
    return clear_table(c, true);
}

// ---------------------------------------------------------------------------------------------------
// build: launch helpers
// ---------------------------------------------------------------------------------------------------
static uint32_t stage_words_for(int R) { return (uint32_t)(((CB + ((R + 15) / 16) * 16) / 16 + 8 + 1) & ~1); }   // even: keeps sink smem 8-B aligned
static size_t build_smem(uint32_t stage_words, uint32_t n_buckets, size_t extra_bytes = 0)
{
    size_t words = (size_t)stage_words + MAXR + MAXR + 2 + 3;   // + 16-byte alignment of the sink's part
    if (n_buckets) words += 2 * (size_t)n_buckets;   // hist + base (u32 each)
    return words * sizeof(u32) + extra_bytes;
}

static int ensure_chunks(dbg_ctx *c, uint64_t n_chunks)
{
    if (n_chunks + 1 <= c->cap_chunks) return DBG_OK;
    CU_TRY(cudaDeviceSynchronize());
    cudaFree(c->d_chunk_first);
    c->d_chunk_first = nullptr;
    c->cap_chunks = n_chunks + 1 + n_chunks / 4;
    CU_TRY(cudaMalloc(&c->d_chunk_first, c->cap_chunks * sizeof(u64)));
    return DBG_OK;
}

template <bool WIDE, class Sink>
static int launch_build(dbg_ctx *c, const BuildArgs &a, Sink sink, uint64_t n_chunks, cudaStream_t s, uint32_t n_buckets = 0,
                        size_t extra_bytes = 0)
{
    size_t smem = build_smem(a.stage_words, n_buckets, extra_bytes);
    if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute(k_build<WIDE, Sink>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_build<WIDE, Sink><<<(unsigned)n_chunks, BLOCK, smem, s>>>(a, sink);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return DBG_OK;
}

template <bool WIDE, bool TRACK>
static int launch_insert(dbg_ctx *c, const void *d_tuples, uint64_t n_upper, const u64 *d_n, cudaStream_t s, bool bucketed, const u32 *d_fill,
                         uint64_t t_begin = 0, uint32_t b_begin = 0)
{
    if (n_upper <= t_begin) return DBG_OK;
    uint64_t tiles = (n_upper - t_begin + INS_TILE - 1) / INS_TILE;
    uint64_t persistent = (uint64_t)c->n_sms * INS_CTAS;
    unsigned grid = (unsigned)(tiles < persistent ? tiles : persistent);
    CU_TRY(cudaMemsetAsync(c->d_counters + 7, 0, sizeof(u64), s));      // tile counter
    k_insert_tuples<WIDE, TRACK><<<grid, INS_BLOCK, 0, s>>>((const u64 *)d_tuples, n_upper, d_n, view_of(c),
                                                            bucketed ? c->d_boffs : nullptr, c->n_buckets, c->part_shift,
                                                            c->d_counters + 7, d_fill, t_begin, b_begin);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return DBG_OK;
}

static int insert_any(dbg_ctx *c, const void *d_tuples, uint64_t n_upper, const u64 *d_n, cudaStream_t s, bool bucketed = false,
                      const u32 *d_fill = nullptr, uint64_t t_begin = 0, uint32_t b_begin = 0)
{
    if (c->wide) return c->track ? launch_insert<true, true>(c, d_tuples, n_upper, d_n, s, bucketed, d_fill, t_begin, b_begin)
                                 : launch_insert<true, false>(c, d_tuples, n_upper, d_n, s, bucketed, d_fill, t_begin, b_begin);
    return c->track ? launch_insert<false, true>(c, d_tuples, n_upper, d_n, s, bucketed, d_fill, t_begin, b_begin)
                    : launch_insert<false, false>(c, d_tuples, n_upper, d_n, s, bucketed, d_fill, t_begin, b_begin);
}

// batch capacity (tuples) of the shared-memory staging of the scatter passes; 0 = store tuples one by one
// (env DBG_B200_STAGE_CAP, experiments).  Must hold one round of BLOCK*G occurrences and index with 16 bits.
static uint32_t stage_cap(const dbg_ctx *c, bool wide)
{
    uint32_t cap = c->stage_cap >= 0 ? (uint32_t)c->stage_cap : 2048u;
    if (cap == 0) return 0;
    if (cap < (uint32_t)BLOCK * G) cap = BLOCK * G;
    if (cap > 32768) cap = 32768;
    const size_t tb = wide ? 32 : 16;
    // keep two CTAs per SM resident next to the build kernel's own shared memory
    while (cap > (uint32_t)BLOCK * G && cap * tb + cap * 6 + (size_t)c->n_buckets * 12 > 98 * 1024) cap /= 2;
    if (cap * tb + cap * 6 + (size_t)c->n_buckets * 12 > 150 * 1024) return 0;
    return cap;
}

// radix-partitioned build of one device-resident block (see PartitionSink): count, scan, scatter, insert
template <bool WIDE>
static int run_partitioned(dbg_ctx *c, BuildArgs a, uint64_t n_chunks, uint64_t occ_upper, cudaStream_t s)
{
    const uint32_t nb = c->n_buckets;
    const uint64_t n_tiles = (n_chunks + PT_CHUNKS - 1) / PT_CHUNKS;
    int rc = DBG_OK;
    // ---- optimistic single pass: every bucket gets a fixed region of the tuple buffer (the buffer holds one tuple
    // per BASE, occurrences are fewer, and hash%P spreads them evenly); batches reserve their runs with atomics.
    // No counting pass, no scan.  Skewed input (a bucket region overflows) falls back to the exact partition.
    const uint32_t cap0 = stage_cap(c, WIDE);
    uint64_t capb64 = (c->opt_capb > 0 ? (uint64_t)c->opt_capb : c->cap_tuples / nb) / INS_TILE * INS_TILE;   // regions start on insert tiles
    if (c->optimistic && cap0 && capb64 >= INS_TILE && capb64 * nb <= c->cap_tuples && capb64 * nb < (1ull << 32)) {
        const uint32_t capb = (uint32_t)capb64;
        if (!c->d_fill) {
            CU_TRY(cudaMalloc(&c->d_fill, ((size_t)4096 + 1) * sizeof(u32)));
            CU_TRY(cudaMalloc(&c->d_snap, (CNT_N + 8) * sizeof(u64)));
            CU_TRY(cudaMallocHost(&c->h_flag, sizeof(u32)));
        }
        CU_TRY(cudaMemcpyAsync(c->d_snap, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CU_TRY(cudaMemcpyAsync(c->d_snap + CNT_N, c->d_polyA, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CU_TRY(cudaMemsetAsync(c->d_fill, 0, ((size_t)nb + 1) * sizeof(u32), s));
        StagedScatterSink<WIDE, true> st; st.t = view_of(c); st.shift = c->part_shift; st.n_buckets = nb; st.cap = cap0;
        st.matrix = nullptr; st.tuples = c->d_tuples; st.fill = c->d_fill; st.capb = capb; st.flag = c->d_fill + nb; st.filled = 0;
        a.count_stats = 1;
        rc = launch_build<WIDE>(c, a, st, n_chunks, s, 0, StageBuf<WIDE>::bytes(cap0, nb));
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(c->h_flag, c->d_fill + nb, sizeof(u32), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        if (*c->h_flag == 0) {
            k_opt_finish<WIDE><<<nb, 256, 0, s>>>(c->d_tuples, c->d_fill, capb, nb, c->d_boffs, (u32)INS_TILE);
            CU_TRY(cudaGetLastError());
            c->launches++;
            c->path_counts[2]++;
            if ((rc = table_ready(c, s))) return rc;       // the scatter above did not touch the table; the insert does
            EvPair ev;
            rc = ev_begin(c, s, &ev);
            if (rc) return rc;
            ev.slot = 6;
            rc = insert_any(c, c->d_tuples, (uint64_t)capb * nb, nullptr, s, true, c->d_fill);
            if (rc) return rc;
            CU_TRY(cudaEventRecord(ev.b, s));
            c->build_ev.push_back(ev);
            return DBG_OK;
        }
        // overflow: nothing was inserted; undo the side counters of the scatter and redo the block exactly
        CU_TRY(cudaMemcpyAsync(c->d_counters, c->d_snap, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CU_TRY(cudaMemcpyAsync(c->d_polyA, c->d_snap + CNT_N, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        c->path_counts[3]++;
    }
    c->path_counts[1]++;
    PartitionSink<WIDE, 0> cs; cs.t = view_of(c); cs.shift = c->part_shift; cs.div = 0; cs.div_M = 0; cs.nb_local = 0; cs.n_buckets = nb;
    cs.matrix = c->d_matrix; cs.tuples = nullptr; cs.dst_ptrs = nullptr; cs.dst_base = nullptr; cs.roffs = nullptr; cs.hist = nullptr; cs.base = nullptr;
    a.count_stats = 0;
    rc = launch_build<WIDE>(c, a, cs, n_chunks, s, nb);
    if (rc) return rc;
    dim3 g1((unsigned)n_tiles, (nb + 255) / 256);
    k_part_scan1<<<g1, 256, 0, s>>>(c->d_matrix, n_chunks, nb, c->d_tile_sums);
    CU_TRY(cudaGetLastError());
    k_part_scan2<<<1, 1024, 0, s>>>(c->d_tile_sums, n_tiles, nb, c->d_boffs);
    CU_TRY(cudaGetLastError());
    k_part_scan3<<<g1, 256, 0, s>>>(c->d_matrix, n_chunks, nb, c->d_tile_sums, c->d_boffs);
    CU_TRY(cudaGetLastError());
    c->launches += 3;
    a.count_stats = 1;
    const uint32_t cap = stage_cap(c, WIDE);
    if (cap) {
        // scatter through shared-memory batches copied out in bucket order (whole-sector runs)
        StagedScatterSink<WIDE> st; st.t = view_of(c); st.shift = c->part_shift; st.n_buckets = nb; st.cap = cap;
        st.matrix = c->d_matrix; st.tuples = c->d_tuples; st.fill = nullptr; st.capb = 0; st.flag = nullptr; st.filled = 0;
        rc = launch_build<WIDE>(c, a, st, n_chunks, s, 0, StageBuf<WIDE>::bytes(cap, nb));
    } else {
        PartitionSink<WIDE, 1> ss; ss.t = view_of(c); ss.shift = c->part_shift; ss.div = 0; ss.div_M = 0; ss.nb_local = 0; ss.n_buckets = nb;
        ss.matrix = c->d_matrix; ss.tuples = c->d_tuples; ss.dst_ptrs = nullptr; ss.dst_base = nullptr; ss.roffs = nullptr; ss.hist = nullptr; ss.base = nullptr;
        rc = launch_build<WIDE>(c, a, ss, n_chunks, s, nb);
    }
    if (rc) return rc;
    if ((rc = table_ready(c, s))) return rc;
    EvPair ev;                       // the insert kernel alone (ms[6]): the roofline's dominant kernel
    rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    ev.slot = 6;
    rc = insert_any(c, c->d_tuples, occ_upper, c->d_boffs + nb, s, true);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

// multi-GPU: extract this rank's occurrences and pack them by OWNER rank (exact, atomic-free, same machinery).
// phase 1 = count + scan (-> per-owner sizes in d_counts), phase 2 = scatter; phase 2 either packs locally
// (d_tuples) or stores straight into the owners' receive buffers over NVLink peer mappings (fused exchange).
template <bool WIDE>
static void fill_rank_sink(dbg_ctx *c, PartitionSink<WIDE, 0> &cs, int n_parts, uint32_t nb_local = 0)
{
    const uint64_t div = (c->P + n_parts - 1) / n_parts;
    cs.t = view_of(c); cs.shift = nb_local ? c->part_shift : 0; cs.div = div; cs.div_M = (uint64_t)((((unsigned __int128)1) << 64) / div);
    cs.nb_local = nb_local;
    cs.n_buckets = nb_local ? (uint32_t)n_parts * nb_local : (uint32_t)n_parts; cs.matrix = c->d_matrix; cs.tuples = nullptr;
    cs.dst_ptrs = nullptr; cs.dst_base = nullptr; cs.roffs = nullptr; cs.hist = nullptr; cs.base = nullptr;
}

template <bool WIDE>
static int run_rank_count(dbg_ctx *c, BuildArgs a, uint64_t n_chunks, int n_parts, u64 *d_counts, cudaStream_t s, uint32_t nb_local = 0)
{
    const uint32_t nb = nb_local ? (uint32_t)n_parts * nb_local : (uint32_t)n_parts;
    const uint64_t n_tiles = (n_chunks + PT_CHUNKS - 1) / PT_CHUNKS;
    PartitionSink<WIDE, 0> cs; fill_rank_sink<WIDE>(c, cs, n_parts, nb_local);
    a.count_stats = 0;
    int rc = launch_build<WIDE>(c, a, cs, n_chunks, s, nb);
    if (rc) return rc;
    dim3 g1((unsigned)n_tiles, (nb + 255) / 256);
    k_part_scan1<<<g1, 256, 0, s>>>(c->d_matrix, n_chunks, nb, c->d_tile_sums);
    CU_TRY(cudaGetLastError());
    k_part_scan2<<<1, 1024, 0, s>>>(c->d_tile_sums, n_tiles, nb, c->d_roffs);
    CU_TRY(cudaGetLastError());
    k_part_scan3<<<g1, 256, 0, s>>>(c->d_matrix, n_chunks, nb, c->d_tile_sums, c->d_roffs);
    CU_TRY(cudaGetLastError());
    k_offsets_to_counts<<<(nb + 255) / 256, 256, 0, s>>>(c->d_roffs, nb, d_counts);
    CU_TRY(cudaGetLastError());
    c->launches += 4;
    return DBG_OK;
}

template <bool WIDE>
static int run_rank_scatter(dbg_ctx *c, BuildArgs a, uint64_t n_chunks, int n_parts, void *d_tuples, u64 *const *d_dst_ptrs,
                            const u64 *d_dst_base, cudaStream_t s, uint32_t nb_local = 0)
{
    PartitionSink<WIDE, 0> cs; fill_rank_sink<WIDE>(c, cs, n_parts, nb_local);
    if (d_dst_ptrs && nb_local == 0 && n_parts <= 32 && !c->peer_unstaged) {
        // fused exchange, owner buckets: sort each round by owner in shared memory, copy out in contiguous runs
        PeerStagedSink<WIDE> ps; ps.t = cs.t; ps.div = cs.div; ps.div_M = cs.div_M; ps.n_buckets = cs.n_buckets; ps.matrix = c->d_matrix;
        ps.dst_ptrs = d_dst_ptrs; ps.dst_base = d_dst_base; ps.roffs = c->d_roffs; ps.cnt = nullptr; ps.base = nullptr; ps.stage = nullptr; ps.parity = 0;
        a.count_stats = 1;
        return launch_build<WIDE>(c, a, ps, n_chunks, s, 0, PeerStagedSink<WIDE>::smem_bytes());
    }
    PartitionSink<WIDE, 1> ss; ss.t = cs.t; ss.shift = cs.shift; ss.div = cs.div; ss.div_M = cs.div_M; ss.nb_local = cs.nb_local; ss.n_buckets = cs.n_buckets;
    ss.matrix = c->d_matrix; ss.tuples = (u64 *)d_tuples; ss.dst_ptrs = d_dst_ptrs; ss.dst_base = d_dst_base; ss.roffs = c->d_roffs;
    ss.hist = nullptr; ss.base = nullptr;
    a.count_stats = 1;
    return launch_build<WIDE>(c, a, ss, n_chunks, s, ss.n_buckets);
}

template <bool WIDE>
static int run_rank_partition(dbg_ctx *c, BuildArgs a, uint64_t n_chunks, int n_parts, void *d_tuples, u64 *d_counts, cudaStream_t s)
{
    int rc = run_rank_count<WIDE>(c, a, n_chunks, n_parts, d_counts, s);
    if (rc) return rc;
    return run_rank_scatter<WIDE>(c, a, n_chunks, n_parts, d_tuples, nullptr, nullptr, s);
}

static bool want_partition(dbg_ctx *c, uint64_t occ_upper)
{
    if (c->part_mode == 0) return false;
    if (c->part_mode == 1) return true;
    // auto (measured, profiles/README.md): the partitioned path costs ~34 ps per occurrence with 64-bit keys (C2) and
    // ~70 ps with 128-bit keys (C3) plus one streaming pass over the table per block (~0.17 ps per table byte); the
    // direct path ~55 ps (64-bit) / ~130 ps (128-bit) per occurrence whatever the table size.  occ_upper counts BASES
    // (about 0.6-0.8 occurrences each) -> partition when the block has more than table_bytes / 80 (/ 150) of them.
    const double per_base = c->wide ? 150.0 : 80.0;
    return c->n_buckets >= 2 && (double)occ_upper * per_base > (double)c->n_local * build_node_bytes(c);
}

static int ensure_matrix(dbg_ctx *c, uint64_t n_chunks, uint32_t nb = 0)
{
    if (nb == 0) nb = c->n_buckets;
    uint64_t cells = n_chunks * (uint64_t)(nb > c->n_buckets ? nb : c->n_buckets);
    if (cells > (1ull << 28)) return DBG_ERR_NOMEM;          // > 1 GiB of offsets: not worth it, use the direct path
    if (cells <= c->cap_matrix) return DBG_OK;
    CU_TRY(cudaDeviceSynchronize());
    cudaFree(c->d_matrix); cudaFree(c->d_tile_sums);
    c->d_matrix = nullptr; c->d_tile_sums = nullptr; c->cap_matrix = 0;
    uint64_t n_tiles = (n_chunks + PT_CHUNKS - 1) / PT_CHUNKS;
    if (cudaMalloc(&c->d_matrix, (cells + 16) * sizeof(u32)) != cudaSuccess) { cudaGetLastError(); return DBG_ERR_NOMEM; }
    if (cudaMalloc(&c->d_tile_sums, (n_tiles + 1) * (uint64_t)(nb > c->n_buckets ? nb : c->n_buckets) * sizeof(u64)) != cudaSuccess) { cudaGetLastError(); cudaFree(c->d_matrix); c->d_matrix = nullptr; return DBG_ERR_NOMEM; }
    c->cap_matrix = cells;
    return DBG_OK;
}

static int ensure_tuples(dbg_ctx *c, uint64_t need)
{
    if (need >= (1ull << 32)) return DBG_ERR_NOMEM;          // write offsets are 32-bit
    if (need <= c->cap_tuples) return DBG_OK;
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    size_t tb = c->wide ? 32 : 16;
    size_t have = c->cap_tuples * tb;
    if ((need + 1024) * tb > (free_b + have) * 6 / 10) return DBG_ERR_NOMEM;   // caller falls back to the direct path
    CU_TRY(cudaDeviceSynchronize());
    cudaFree(c->d_tuples); c->d_tuples = nullptr; c->cap_tuples = 0;
    cudaError_t e = cudaMalloc(&c->d_tuples, (need + 1024) * tb);
    if (e != cudaSuccess) { cudaGetLastError(); return DBG_ERR_NOMEM; }
    c->cap_tuples = need;
    return DBG_OK;
}

// d_bases: pointer such that d_bases[off] is base `off` of the global offset space
static int build_device(dbg_ctx *c, const char *d_bases, const u64 *d_offs, uint64_t n_reads, uint64_t first_base,
                        uint64_t total_bases, uint64_t read_index0, cudaStream_t s,
                        int n_parts, void *d_tuples, uint64_t bucket_stride, u64 *d_counts)
{
    if (n_reads == 0 || total_bases == 0) return DBG_OK;
    c->guard_total += total_bases;
    uint64_t abase = first_base & ~15ull;
    if (((uintptr_t)(d_bases + abase) & 15) != 0) return set_err(DBG_ERR_INVALID, "device base buffer must be 16-byte aligned");
    uint64_t n_chunks = (first_base + total_bases - abase + CB - 1) / CB;
    if (n_chunks > 0x7fffffffull) return set_err(DBG_ERR_INVALID, "block too large: %llu chunks", (unsigned long long)n_chunks);
    int rc = ensure_chunks(c, n_chunks);
    if (rc) return rc;
    bool part = n_parts == 0 && want_partition(c, total_bases);
    if (part && (ensure_tuples(c, total_bases) != DBG_OK || ensure_matrix(c, n_chunks) != DBG_OK)) part = false;
    // kernels wait for the table clear that dbg_reset started on the side stream.  (It may only overlap COPIES: running the
    // 6.4-GB memset next to the extraction / partition kernel was measured -- the memset takes the DRAM write bandwidth the
    // partition's scattered stores need and the step got slower, 17.5 vs 17.3 ms on C2.)
    if ((rc = table_ready(c, s))) return rc;
    if (n_parts > 0) {
        if (total_bases > bucket_stride) return set_err(DBG_ERR_BUFFER, "tuple capacity %llu < %llu (bases in the block)", (unsigned long long)bucket_stride, (unsigned long long)total_bases);
        if (total_bases >= (1ull << 32)) return set_err(DBG_ERR_INVALID, "block too large for the exchange: split it (< 2^32 bases)");
        if (ensure_matrix(c, n_chunks, (uint32_t)n_parts) != DBG_OK) return set_err(DBG_ERR_NOMEM, "partition offsets");
    }

    EvPair ev;
    rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    unsigned gb = (unsigned)((n_reads + 1 + 255) / 256);
    k_chunk_first<<<gb, 256, 0, s>>>(d_offs, n_reads, abase, n_chunks, c->d_chunk_first);
    CU_TRY(cudaGetLastError());
    c->launches++;

    BuildArgs a;
    a.bases = d_bases; a.offs = d_offs; a.n_reads = n_reads; a.abase = abase; a.end_base = first_base + total_bases;
    a.chunk_first = c->d_chunk_first; a.read_index0 = read_index0; a.K = c->prm.K; a.R = c->prm.max_read_len;
    a.stage_words = stage_words_for(c->prm.max_read_len); a.count_stats = 1; a.seed = c->prm.payload_mode == 1;

    if (part) {
        rc = c->wide ? run_partitioned<true>(c, a, n_chunks, total_bases, s) : run_partitioned<false>(c, a, n_chunks, total_bases, s);
        c->part_blocks++;
    } else if (n_parts > 0) {
        rc = c->wide ? run_rank_partition<true>(c, a, n_chunks, n_parts, d_tuples, d_counts, s)
                     : run_rank_partition<false>(c, a, n_chunks, n_parts, d_tuples, d_counts, s);
    } else if (c->wide) {
        c->path_counts[0]++;
        if (c->track) { InsertSink<true, true> sk; sk.t = view_of(c); rc = launch_build<true>(c, a, sk, n_chunks, s); }
        else { InsertSink<true, false> sk; sk.t = view_of(c); rc = launch_build<true>(c, a, sk, n_chunks, s); }
    } else {
        c->path_counts[0]++;
        if (c->track) { InsertSink<false, true> sk; sk.t = view_of(c); rc = launch_build<false>(c, a, sk, n_chunks, s); }
        else { InsertSink<false, false> sk; sk.t = view_of(c); rc = launch_build<false>(c, a, sk, n_chunks, s); }
    }
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

// ---- host submit: reads are appended to a device-resident batch (double buffered); a batch is built when
// ---- it is full or when results are requested, so that the partitioned path sees enough occurrences per
// ---- table slice.  Copies of batch n+1 overlap the kernels of batch n.
static int ensure_batch(dbg_ctx *c)
{
    if (c->d_bases[0]) return DBG_OK;
    for (int i = 0; i < 2; i++) {
        CU_TRY(cudaMalloc(&c->d_bases[i], c->cap_bases + 64));
        CU_TRY(cudaMalloc(&c->d_offs[i], (c->cap_reads + 2) * sizeof(u64)));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
    }
    CU_TRY(cudaMalloc(&c->d_offs_stage, (c->sub_reads + 2) * sizeof(u64)));
    return DBG_OK;
}

__global__ void k_append_offs(const u64 *__restrict__ stage, u64 n, u64 *dst, u64 base)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) dst[i] = stage[i] - stage[0] + base;
}

// A table that cannot hold the input is reported as soon as a batch has shown it (the counters travel to pinned memory
// behind every batch): DBG_ERR_TABLE_FULL from the next dbg_submit_reads / flush, so that a front end can rebuild with a
// larger table right away instead of streaming the rest of the input into a full one.
static int check_table_full(dbg_ctx *c, bool wait)
{
    if (!c->cnt_pending) return DBG_OK;
    if (wait) CU_TRY(cudaEventSynchronize(c->ev_cnt));
    else {
        cudaError_t q = cudaEventQuery(c->ev_cnt);
        if (q == cudaErrorNotReady) { cudaGetLastError(); return DBG_OK; }
        CU_TRY(q);
    }
    c->cnt_pending = false;
    if (c->h_cnt[CNT_ERROR] || (c->n_shards <= 1 && c->h_cnt[CNT_NEW] + 1 > c->P))
        return set_err(DBG_ERR_TABLE_FULL, "%llu nodes so far: the table of %llu slots cannot hold this input", (unsigned long long)c->h_cnt[CNT_NEW],
                       (unsigned long long)c->P);
    return DBG_OK;
}

static int flush_batch(dbg_ctx *c)
{
    if (c->batch_reads == 0) return DBG_OK;
    int b = c->cur;
    int rc = check_table_full(c, true);       // (the previous batch's kernels precede this batch's on the stream anyway)
    if (rc) return rc;
    rc = build_device(c, c->d_bases[b], c->d_offs[b], c->batch_reads, 0, c->batch_bases, c->batch_read_index0, c->stream,
                      0, nullptr, 0, nullptr);
    if (rc) return rc;
    if (!c->h_cnt) {
        CU_TRY(cudaMallocHost(&c->h_cnt, CNT_N * sizeof(u64)));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_cnt, cudaEventDisableTiming));
    }
    CU_TRY(cudaMemcpyAsync(c->h_cnt, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaEventRecord(c->ev_cnt, c->stream));
    c->cnt_pending = true;
    CU_TRY(cudaEventRecord(c->ev_free[b], c->stream));
    c->cur ^= 1;
    c->batch_reads = 0; c->batch_bases = 0;
    // the other buffer may still be read by the previous batch's kernels
    CU_TRY(cudaEventSynchronize(c->ev_free[c->cur]));
    return DBG_OK;
}

// dbg_finish_export: where the caller wants the table image (filled by the grouped insert / layout / copy pipeline)
struct FinishExport { void *array; uint8_t *nul_flag; bool done; };
template <bool WIDE, bool TRACK>
static int grouped_finish(dbg_ctx *c, uint32_t capb, FinishExport *fx);

// One dbg_submit_reads call that is a whole partitioned block by itself (front ends that hand over a file's worth of reads,
// bench e2e): pipeline it.  The optimistic scatter appends to fixed bucket regions, so it can run sub-block by sub-block --
// the extraction + partition of sub-block i overlaps the host->device copy of sub-block i+1 -- and ONE bucketed insert
// follows the last scatter.  Same table as the one-launch build (ordinals are global read indices); an overflowing region
// falls back to the exact partition of the whole (device-resident) batch.
template <bool WIDE>
static int submit_pipelined(dbg_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads, FinishExport *fx)
{
    const uint64_t call_bases = offs[n_reads] - offs[0];
    const int b = c->cur;
    const uint32_t nb = c->n_buckets;
    const uint32_t cap0 = stage_cap(c, WIDE);
    const uint64_t capb64 = (c->opt_capb > 0 ? (uint64_t)c->opt_capb : c->cap_tuples / nb) / INS_TILE * INS_TILE;
    const uint32_t capb = (uint32_t)capb64;
    cudaStream_t s = c->stream;
    int rc = DBG_OK;
    if (!c->d_fill) {
        CU_TRY(cudaMalloc(&c->d_fill, ((size_t)4096 + 1) * sizeof(u32)));
        CU_TRY(cudaMalloc(&c->d_snap, (CNT_N + 8) * sizeof(u64)));
        CU_TRY(cudaMallocHost(&c->h_flag, sizeof(u32)));
    }
    c->guard_total += call_bases;
    c->batch_read_index0 = c->next_read_index;
    CU_TRY(cudaMemcpyAsync(c->d_snap, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->d_snap + CNT_N, c->d_polyA, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaMemsetAsync(c->d_fill, 0, ((size_t)nb + 1) * sizeof(u32), s));
    EvPair ev_build, ev_copy;
    rc = ev_begin(c, s, &ev_build);
    if (rc) return rc;
    rc = ev_begin(c, c->copy_stream, &ev_copy);
    if (rc) return rc;
    std::vector<cudaEvent_t> sub_ev;
    const uint64_t SUB_BASES = c->sub_bases, SUB_READS = c->sub_reads;
    uint64_t r0 = 0;
    while (r0 < n_reads) {
        uint64_t lim = r0 + SUB_READS < n_reads ? r0 + SUB_READS : n_reads;
        uint64_t lo = r0 + 1, hi = lim;
        while (lo < hi) { uint64_t mid = (lo + hi + 1) / 2; if (offs[mid] - offs[r0] <= SUB_BASES) lo = mid; else hi = mid - 1; }
        const uint64_t r1 = lo;
        if (offs[r1] < offs[r0]) return set_err(DBG_ERR_INVALID, "offsets must be non-decreasing");
        const uint64_t nbases = offs[r1] - offs[r0], nr = r1 - r0;
        if (c->batch_bases + nbases > c->cap_bases || c->batch_reads + nr > c->cap_reads)
            return set_err(DBG_ERR_INVALID, "a single read longer than the sub-block size inside a pipelined submit");
        if (nbases) CU_TRY(cudaMemcpyAsync(c->d_bases[b] + c->batch_bases, bases + offs[r0], nbases, cudaMemcpyHostToDevice, c->copy_stream));
        // (one staging buffer for the offsets: the copy stream runs copy -> append in order, so it is free again when the
        // next sub-block's offsets arrive)
        CU_TRY(cudaMemcpyAsync(c->d_offs_stage, offs + r0, (nr + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->copy_stream));
        k_append_offs<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, c->copy_stream>>>(c->d_offs_stage, nr, c->d_offs[b] + c->batch_reads, c->batch_bases);
        CU_TRY(cudaGetLastError());
        c->launches++;
        cudaEvent_t e;
        rc = ev_get(c, &e);
        if (rc) return rc;
        sub_ev.push_back(e);
        CU_TRY(cudaEventRecord(e, c->copy_stream));
        CU_TRY(cudaStreamWaitEvent(s, e, 0));
        if ((rc = table_ready(c, s))) return rc;       // (first sub-block: the table clear of dbg_reset ran beside its copy)
        if (nbases) {
            const uint64_t first_base = c->batch_bases, abase = first_base & ~15ull;
            const uint64_t n_chunks = (first_base + nbases - abase + CB - 1) / CB;
            rc = ensure_chunks(c, n_chunks);     // (grows with a device-wide synchronize, so launches in flight are safe)
            if (rc) return rc;
            k_chunk_first<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, s>>>(c->d_offs[b] + c->batch_reads, nr, abase, n_chunks, c->d_chunk_first);
            CU_TRY(cudaGetLastError());
            c->launches++;
            BuildArgs a;
            a.bases = c->d_bases[b]; a.offs = c->d_offs[b] + c->batch_reads; a.n_reads = nr; a.abase = abase; a.end_base = first_base + nbases;
            a.chunk_first = c->d_chunk_first; a.read_index0 = c->batch_read_index0 + c->batch_reads; a.K = c->prm.K; a.R = c->prm.max_read_len;
            a.stage_words = stage_words_for(c->prm.max_read_len); a.count_stats = 1; a.seed = c->prm.payload_mode == 1;
            StagedScatterSink<WIDE, true> st; st.t = view_of(c); st.shift = c->part_shift; st.n_buckets = nb; st.cap = cap0;
            st.matrix = nullptr; st.tuples = c->d_tuples; st.fill = c->d_fill; st.capb = capb; st.flag = c->d_fill + nb; st.filled = 0;
            rc = launch_build<WIDE>(c, a, st, n_chunks, s, 0, StageBuf<WIDE>::bytes(cap0, nb));
            if (rc) return rc;
        }
        c->batch_bases += nbases; c->batch_reads += nr;
        r0 = r1;
    }
    CU_TRY(cudaEventRecord(ev_copy.b, c->copy_stream));
    CU_TRY(cudaMemcpyAsync(c->h_flag, c->d_fill + nb, sizeof(u32), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));            // every copy has landed too: the host buffers are reusable from here on
    { float t = 0; CU_TRY(cudaEventElapsedTime(&t, ev_copy.a, ev_copy.b)); c->ms[4] += t; ev_put(c, ev_copy); }
    for (cudaEvent_t e : sub_ev) c->ev_free_list.push_back(e);
    c->next_read_index += n_reads; c->reads_total += n_reads;
    if (*c->h_flag == 0) {
        k_opt_finish<WIDE><<<nb, 256, 0, s>>>(c->d_tuples, c->d_fill, capb, nb, c->d_boffs, (u32)INS_TILE);
        CU_TRY(cudaGetLastError());
        c->launches++;
        c->path_counts[2]++;
        c->part_blocks++;
        if ((rc = table_ready(c, s))) return rc;
        int grc = 1;
        if (fx) {
            // last block of the build and the caller waits for the image: insert, lay out and copy slice group by slice group
            CU_TRY(cudaEventRecord(ev_build.b, s));
            c->build_ev.push_back(ev_build);
            grc = c->track ? grouped_finish<WIDE, true>(c, capb, fx) : grouped_finish<WIDE, false>(c, capb, fx);
            if (grc < 0) return grc;
        }
        if (grc == 1) {            // (1 = the grouped pipeline does not apply here: nothing was inserted yet)
            EvPair ev;
            rc = ev_begin(c, s, &ev);
            if (rc) return rc;
            ev.slot = 6;
            rc = insert_any(c, c->d_tuples, (uint64_t)capb * nb, nullptr, s, true, c->d_fill);
            if (rc) return rc;
            CU_TRY(cudaEventRecord(ev.b, s));
            c->build_ev.push_back(ev);
            if (!fx) { CU_TRY(cudaEventRecord(ev_build.b, s)); c->build_ev.push_back(ev_build); }
        }
    } else {
        // a bucket region overflowed (skewed input): nothing was inserted; undo the side counters and build the whole
        // batch -- it is resident -- through the exact partition
        CU_TRY(cudaMemcpyAsync(c->d_counters, c->d_snap, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CU_TRY(cudaMemcpyAsync(c->d_polyA, c->d_snap + CNT_N, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CU_TRY(cudaEventRecord(ev_build.b, s));
        c->build_ev.push_back(ev_build);
        c->path_counts[3]++;
        c->guard_total -= call_bases;            // build_device counts the block again
        const int opt = c->optimistic;
        c->optimistic = 0;
        rc = build_device(c, c->d_bases[b], c->d_offs[b], c->batch_reads, 0, c->batch_bases, c->batch_read_index0, s, 0, nullptr, 0, nullptr);
        c->optimistic = opt;
        if (rc) return rc;
    }
    // the tail of flush_batch: counters to pinned memory (prompt table-full report), hand the buffer over
    if (!c->h_cnt) {
        CU_TRY(cudaMallocHost(&c->h_cnt, CNT_N * sizeof(u64)));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_cnt, cudaEventDisableTiming));
    }
    CU_TRY(cudaMemcpyAsync(c->h_cnt, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaEventRecord(c->ev_cnt, s));
    c->cnt_pending = true;
    CU_TRY(cudaEventRecord(c->ev_free[b], s));
    c->cur ^= 1;
    c->batch_reads = 0; c->batch_bases = 0;
    CU_TRY(cudaEventSynchronize(c->ev_free[c->cur]));
    return DBG_OK;
}

static int submit_impl(dbg_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads, FinishExport *fx)
{
    if (!c || (!bases && n_reads) || (!offs && n_reads)) return set_err(DBG_ERR_INVALID, "dbg_submit_reads: NULL argument");
    if (c->finalized) return set_err(DBG_ERR_STATE, "submit after finalize");
    if (c->n_shards > 1) return set_err(DBG_ERR_STATE, "sharded contexts take tuples (dbg_insert_tuples_device)");
    CU_TRY(cudaSetDevice(c->device));
    int rc = check_table_full(c, false);
    if (rc) return rc;
    rc = ensure_batch(c);
    if (rc) return rc;
    if (n_reads && c->pipeline && c->optimistic && c->part_mode != 0 && offs[n_reads] >= offs[0]) {
        // a call that is a partitioned block by itself and spans several sub-blocks: scatter while copying
        const uint64_t call_bases = offs[n_reads] - offs[0];
        if (call_bases > c->sub_bases && call_bases <= c->cap_bases && n_reads <= c->cap_reads && want_partition(c, call_bases) &&
            stage_cap(c, c->wide) != 0) {
            rc = flush_batch(c);                 // reads of earlier, smaller calls: their own block, in call order
            if (rc) return rc;
            const uint32_t nbk = c->n_buckets;
            if (ensure_tuples(c, call_bases) == DBG_OK) {
                const uint64_t capb64 = (c->opt_capb > 0 ? (uint64_t)c->opt_capb : c->cap_tuples / nbk) / INS_TILE * INS_TILE;
                if (capb64 >= INS_TILE && capb64 * nbk <= c->cap_tuples && capb64 * nbk < (1ull << 32))
                    return c->wide ? submit_pipelined<true>(c, bases, offs, n_reads, fx) : submit_pipelined<false>(c, bases, offs, n_reads, fx);
            }
        }
    }
    const uint64_t SUB_BASES = c->sub_bases, SUB_READS = c->sub_reads;
    uint64_t r0 = 0;
    while (r0 < n_reads) {
        // cut a sub-block: at most SUB_BASES bases / SUB_READS reads (at least one read)
        uint64_t lim = r0 + SUB_READS < n_reads ? r0 + SUB_READS : n_reads;
        uint64_t lo = r0 + 1, hi = lim;
        while (lo < hi) { uint64_t mid = (lo + hi + 1) / 2; if (offs[mid] - offs[r0] <= SUB_BASES) lo = mid; else hi = mid - 1; }
        uint64_t r1 = lo;
        if (offs[r1] < offs[r0]) return set_err(DBG_ERR_INVALID, "offsets must be non-decreasing");
        uint64_t nb = offs[r1] - offs[r0], nr = r1 - r0;
        uint64_t copy_nb = nb;
        if (nb > c->cap_bases) {
            // a single sequence longer than a whole batch: only its first max_read_len bases are ever used
            // (DBGgraph.cpp:63); the logged untrimmed k-mer count is the only thing that changes
            copy_nb = (uint64_t)c->prm.max_read_len < nb ? (uint64_t)c->prm.max_read_len : nb;
        }
        if (c->batch_bases + copy_nb > c->cap_bases || c->batch_reads + nr > c->cap_reads) {
            rc = flush_batch(c);
            if (rc) return rc;
        }
        int b = c->cur;
        if (c->batch_reads == 0) c->batch_read_index0 = c->next_read_index;
        EvPair ev;
        rc = ev_begin(c, c->copy_stream, &ev);
        if (rc) return rc;
        if (copy_nb) CU_TRY(cudaMemcpyAsync(c->d_bases[b] + c->batch_bases, bases + offs[r0], copy_nb, cudaMemcpyHostToDevice, c->copy_stream));
        if (copy_nb == nb) {
            CU_TRY(cudaMemcpyAsync(c->d_offs_stage, offs + r0, (nr + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->copy_stream));
        } else {
            uint64_t two[2] = {0, copy_nb};
            CU_TRY(cudaMemcpyAsync(c->d_offs_stage, two, sizeof(two), cudaMemcpyHostToDevice, c->copy_stream));
        }
        k_append_offs<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, c->copy_stream>>>(c->d_offs_stage, nr, c->d_offs[b] + c->batch_reads, c->batch_bases);
        CU_TRY(cudaGetLastError());
        c->launches++;
        CU_TRY(cudaEventRecord(ev.b, c->copy_stream));
        CU_TRY(cudaEventSynchronize(ev.b));          // host buffer is reusable from here on
        float t = 0; CU_TRY(cudaEventElapsedTime(&t, ev.a, ev.b)); c->ms[4] += t;
        ev_put(c, ev);
        c->batch_bases += copy_nb; c->batch_reads += nr;
        c->next_read_index += nr; c->reads_total += nr;
        r0 = r1;
    }
    return DBG_OK;
}

extern "C" int dbg_submit_reads(dbg_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads)
{
    return submit_impl(c, bases, offs, n_reads, nullptr);
}

extern "C" int dbg_submit_reads_device(dbg_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                       uint64_t first_base, uint64_t total_bases, uint64_t first_read_index, void *stream)
{
    if (!c || (!d_bases && n_reads) || (!d_offs && n_reads)) return set_err(DBG_ERR_INVALID, "dbg_submit_reads_device: NULL argument");
    if (c->finalized) return set_err(DBG_ERR_STATE, "submit after finalize");
    if (c->n_shards > 1) return set_err(DBG_ERR_STATE, "sharded contexts take tuples (dbg_insert_tuples_device)");
    CU_TRY(cudaSetDevice(c->device));
    int rc = flush_batch(c);      // keep blocks in call order
    if (rc) return rc;
    uint64_t idx0 = first_read_index == UINT64_MAX ? c->next_read_index : first_read_index;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    rc = build_device(c, d_bases, (const u64 *)d_offs, n_reads, first_base, total_bases, idx0, s, 0, nullptr, 0, nullptr);
    if (rc) return rc;
    c->next_read_index = idx0 + n_reads; c->reads_total += n_reads;
    return DBG_OK;
}

extern "C" int dbg_tuple_bytes(const dbg_ctx *c) { return c ? (c->wide ? 32 : 16) : 0; }

extern "C" int dbg_extract_tuples_device(dbg_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                         uint64_t first_base, uint64_t total_bases, uint64_t first_read_index,
                                         int32_t n_parts, void *d_tuples, uint64_t capacity, uint64_t *d_counts, void *stream)
{
    if (!c || !d_tuples || !d_counts || n_parts < 1 || n_parts > 64) return set_err(DBG_ERR_INVALID, "dbg_extract_tuples_device: bad argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    int rc = build_device(c, d_bases, (const u64 *)d_offs, n_reads, first_base, total_bases, first_read_index, s,
                          n_parts, d_tuples, capacity, (u64 *)d_counts);
    if (rc) return rc;
    c->reads_total += n_reads;
    return DBG_OK;
}

// ---- fused exchange over NVLink peer memory ---------------------------------------------------------
static int rank_phase(dbg_ctx *c, int phase, const char *d_bases, const u64 *d_offs, uint64_t n_reads, uint64_t first_base,
                      uint64_t total_bases, uint64_t read_index0, int n_parts, int by_slice, u64 *d_counts, u64 *const *d_dst_ptrs,
                      const u64 *d_dst_base, cudaStream_t s)
{
    if (n_parts < 1 || n_parts > 64) return set_err(DBG_ERR_INVALID, "n_parts outside 1..64");
    // by_slice: buckets = (owner, 16-MB table slice of the owner); every shard has the same slice count as this one
    const uint32_t nb_local = by_slice ? c->n_buckets : 0;
    const uint32_t nb_total = nb_local ? (uint32_t)n_parts * nb_local : (uint32_t)n_parts;
    if (nb_total > 64 * 1024 || 2 * (size_t)nb_total * sizeof(u32) > 160 * 1024) return set_err(DBG_ERR_INVALID, "too many exchange buckets (%u)", nb_total);
    if (n_reads == 0 || total_bases == 0) {
        if (phase == 1) CU_TRY(cudaMemsetAsync(d_counts, 0, nb_total * sizeof(u64), s));
        return DBG_OK;
    }
    uint64_t abase = first_base & ~15ull;
    if (((uintptr_t)(d_bases + abase) & 15) != 0) return set_err(DBG_ERR_INVALID, "device base buffer must be 16-byte aligned");
    uint64_t n_chunks = (first_base + total_bases - abase + CB - 1) / CB;
    if (n_chunks > 0x7fffffffull || total_bases >= (1ull << 32)) return set_err(DBG_ERR_INVALID, "block too large for the exchange: split it (< 2^32 bases)");
    int rc = ensure_chunks(c, n_chunks);
    if (rc) return rc;
    if (ensure_matrix(c, n_chunks, nb_total) != DBG_OK) return set_err(DBG_ERR_NOMEM, "partition offsets");
    EvPair ev;
    rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    BuildArgs a;
    a.bases = d_bases; a.offs = d_offs; a.n_reads = n_reads; a.abase = abase; a.end_base = first_base + total_bases;
    a.chunk_first = c->d_chunk_first; a.read_index0 = read_index0; a.K = c->prm.K; a.R = c->prm.max_read_len;
    a.stage_words = stage_words_for(c->prm.max_read_len); a.count_stats = 1; a.seed = c->prm.payload_mode == 1;
    if (phase == 1) {
        unsigned gb = (unsigned)((n_reads + 1 + 255) / 256);
        k_chunk_first<<<gb, 256, 0, s>>>(d_offs, n_reads, abase, n_chunks, c->d_chunk_first);
        CU_TRY(cudaGetLastError());
        c->launches++;
        rc = c->wide ? run_rank_count<true>(c, a, n_chunks, n_parts, d_counts, s, nb_local) : run_rank_count<false>(c, a, n_chunks, n_parts, d_counts, s, nb_local);
    } else {
        rc = c->wide ? run_rank_scatter<true>(c, a, n_chunks, n_parts, nullptr, d_dst_ptrs, d_dst_base, s, nb_local)
                     : run_rank_scatter<false>(c, a, n_chunks, n_parts, nullptr, d_dst_ptrs, d_dst_base, s, nb_local);
        c->reads_total += n_reads;
    }
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

extern "C" int dbg_exchange_count_device(dbg_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads, uint64_t first_base,
                                         uint64_t total_bases, int32_t n_parts, int32_t by_slice, uint64_t *d_counts, void *stream)
{
    if (!c || !d_counts) return set_err(DBG_ERR_INVALID, "dbg_exchange_count_device: NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    return rank_phase(c, 1, d_bases, (const u64 *)d_offs, n_reads, first_base, total_bases, 0, n_parts, by_slice, (u64 *)d_counts, nullptr, nullptr,
                      stream ? (cudaStream_t)stream : c->stream);
}

extern "C" int dbg_exchange_scatter_device(dbg_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads, uint64_t first_base,
                                           uint64_t total_bases, uint64_t first_read_index, int32_t n_parts, int32_t by_slice,
                                           void *const *d_dst_ptrs, const uint64_t *d_dst_base, void *stream)
{
    if (!c || !d_dst_ptrs || !d_dst_base) return set_err(DBG_ERR_INVALID, "dbg_exchange_scatter_device: NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    return rank_phase(c, 2, d_bases, (const u64 *)d_offs, n_reads, first_base, total_bases, first_read_index, n_parts, by_slice, nullptr,
                      (u64 *const *)d_dst_ptrs, (const u64 *)d_dst_base, stream ? (cudaStream_t)stream : c->stream);
}

extern "C" int dbg_peer_alloc(dbg_ctx *c, uint64_t bytes, void **d_ptr, uint8_t handle[64])
{
    if (!c || !d_ptr || !handle) return set_err(DBG_ERR_INVALID, "dbg_peer_alloc: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaMalloc(d_ptr, bytes ? bytes : 256));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, *d_ptr));
    memcpy(handle, &h, 64);
    return DBG_OK;
}

extern "C" int dbg_peer_open(dbg_ctx *c, const uint8_t handle[64], void **d_ptr)
{
    if (!c || !d_ptr || !handle) return set_err(DBG_ERR_INVALID, "dbg_peer_open: NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DBG_OK;
}

extern "C" int dbg_peer_close(dbg_ctx *c, void *d_ptr)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    CU_TRY(cudaSetDevice(c->device));
    if (d_ptr) CU_TRY(cudaIpcCloseMemHandle(d_ptr));
    return DBG_OK;
}

extern "C" int dbg_peer_free(dbg_ctx *c, void *d_ptr)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    CU_TRY(cudaSetDevice(c->device));
    if (d_ptr) CU_TRY(cudaFree(d_ptr));
    return DBG_OK;
}

template <bool WIDE>
static int partition_tuples(dbg_ctx *c, const void *d_src, uint64_t n, cudaStream_t s)
{
    const uint32_t nb = c->n_buckets;
    const uint64_t n_rows = (n + TP_TILE - 1) / TP_TILE;
    const uint64_t n_tiles = (n_rows + PT_CHUNKS - 1) / PT_CHUNKS;
    size_t smem = 2 * (size_t)nb * sizeof(u32);
    TableView t = view_of(c);
    k_tuple_partition<WIDE, 0><<<(unsigned)n_rows, 256, smem, s>>>((const u64 *)d_src, n, t, c->part_shift, nb, c->d_matrix, nullptr);
    CU_TRY(cudaGetLastError());
    dim3 g1((unsigned)n_tiles, (nb + 255) / 256);
    k_part_scan1<<<g1, 256, 0, s>>>(c->d_matrix, n_rows, nb, c->d_tile_sums);
    CU_TRY(cudaGetLastError());
    k_part_scan2<<<1, 1024, 0, s>>>(c->d_tile_sums, n_tiles, nb, c->d_boffs);
    CU_TRY(cudaGetLastError());
    k_part_scan3<<<g1, 256, 0, s>>>(c->d_matrix, n_rows, nb, c->d_tile_sums, c->d_boffs);
    CU_TRY(cudaGetLastError());
    const size_t st_bytes = StageBuf<WIDE>::bytes(TP_TILE, nb);
    if (c->stage_cap != 0 && st_bytes <= 200 * 1024) {
        if (st_bytes > 48 * 1024)
            CU_TRY(cudaFuncSetAttribute(k_tuple_scatter_staged<WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_bytes));
        k_tuple_scatter_staged<WIDE, false><<<(unsigned)n_rows, 256, st_bytes, s>>>((const u64 *)d_src, n, t, c->part_shift, nb, c->d_matrix, c->d_tuples,
                                                                                     nullptr, 0, nullptr);
    } else {
        k_tuple_partition<WIDE, 1><<<(unsigned)n_rows, 256, smem, s>>>((const u64 *)d_src, n, t, c->part_shift, nb, c->d_matrix, c->d_tuples);
    }
    CU_TRY(cudaGetLastError());
    c->launches += 5;
    return DBG_OK;
}

// optimistic single-pass partition of received tuples (see run_partitioned): returns 1 when it was used (the caller
// inserts with the fixed regions), 0 when it does not apply or overflowed (caller partitions exactly), < 0 on error
static int ensure_opt_buffers(dbg_ctx *c)
{
    if (!c->d_fill) {
        CU_TRY(cudaMalloc(&c->d_fill, ((size_t)4096 + 1) * sizeof(u32)));
        CU_TRY(cudaMalloc(&c->d_snap, (CNT_N + 8) * sizeof(u64)));
        CU_TRY(cudaMallocHost(&c->h_flag, sizeof(u32)));
    }
    return DBG_OK;
}

// the source tuples come as n_regions arrays: region r = counts[r] tuples at base + r * stride (tuples)
template <bool WIDE>
static int partition_tuples_optimistic(dbg_ctx *c, const void *d_base, uint32_t n_regions, uint64_t stride, const uint64_t *counts, cudaStream_t s,
                                       uint32_t *capb_out)
{
    const uint32_t nb = c->n_buckets;
    const size_t st_bytes = StageBuf<WIDE>::bytes(TP_TILE, nb);
    uint64_t capb64 = (c->opt_capb > 0 ? (uint64_t)c->opt_capb : c->cap_tuples / nb) / INS_TILE * INS_TILE;
    if (!c->optimistic || c->stage_cap == 0 || st_bytes > 200 * 1024 || capb64 < INS_TILE || capb64 * nb > c->cap_tuples ||
        capb64 * nb >= (1ull << 32)) return 0;
    const uint32_t capb = (uint32_t)capb64;
    if (ensure_opt_buffers(c) != DBG_OK) return DBG_ERR_CUDA;
    CU_TRY(cudaMemsetAsync(c->d_fill, 0, ((size_t)nb + 1) * sizeof(u32), s));
    if (st_bytes > 48 * 1024)
        CU_TRY(cudaFuncSetAttribute(k_tuple_scatter_staged<WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_bytes));
    const size_t tw = WIDE ? 4 : 2;
    for (uint32_t r = 0; r < n_regions; r++) {
        const uint64_t n = counts[r];
        if (n == 0) continue;
        const uint64_t n_rows = (n + TP_TILE - 1) / TP_TILE;
        const u64 *src = static_cast<const u64 *>(d_base) + (size_t)r * stride * tw;
        k_tuple_scatter_staged<WIDE, true><<<(unsigned)n_rows, 256, st_bytes, s>>>(src, n, view_of(c), c->part_shift, nb, nullptr, c->d_tuples,
                                                                                   c->d_fill, capb, c->d_fill + nb);
        CU_TRY(cudaGetLastError());
        c->launches++;
    }
    CU_TRY(cudaMemcpyAsync(c->h_flag, c->d_fill + nb, sizeof(u32), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    if (*c->h_flag != 0) { c->path_counts[3]++; return 0; }
    k_opt_finish<WIDE><<<nb, 256, 0, s>>>(c->d_tuples, c->d_fill, capb, nb, c->d_boffs, (u32)INS_TILE);
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->path_counts[2]++;
    *capb_out = capb;
    return 1;
}

extern "C" int dbg_insert_tuples_device(dbg_ctx *c, const void *d_tuples, uint64_t n, void *stream)
{
    if (!c || (!d_tuples && n)) return set_err(DBG_ERR_INVALID, "dbg_insert_tuples_device: NULL argument");
    if (c->finalized) return set_err(DBG_ERR_STATE, "insert after finalize");
    if (n == 0) return DBG_OK;
    c->guard_total += n;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    { int trc = table_ready(c, s); if (trc) return trc; }
    // enough tuples per table slice: put them in slice order first, then insert through L2-resident slices
    bool part = want_partition(c, n) && 2 * (size_t)c->n_buckets * sizeof(u32) <= 48 * 1024;
    // room for the fixed bucket regions of the optimistic partition: 1/8 slack plus a tile per bucket
    const uint64_t want_cap = c->optimistic ? n + n / 8 + (uint64_t)c->n_buckets * INS_TILE : n;
    if (part && ensure_tuples(c, want_cap) != DBG_OK && ensure_tuples(c, n) != DBG_OK) part = false;
    if (part && ensure_matrix(c, (n + TP_TILE - 1) / TP_TILE) != DBG_OK) part = false;
    EvPair ev;
    int rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    if (part) {
        uint32_t capb = 0;
        const uint64_t one[1] = {n};
        int opt = c->wide ? partition_tuples_optimistic<true>(c, d_tuples, 1, 0, one, s, &capb) : partition_tuples_optimistic<false>(c, d_tuples, 1, 0, one, s, &capb);
        if (opt < 0) return opt;
        if (!opt) {
            c->path_counts[1]++;
            rc = c->wide ? partition_tuples<true>(c, d_tuples, n, s) : partition_tuples<false>(c, d_tuples, n, s);
            if (rc) return rc;
        }
        EvPair ei;
        rc = ev_begin(c, s, &ei);
        if (rc) return rc;
        ei.slot = 6;
        rc = opt ? insert_any(c, c->d_tuples, (uint64_t)capb * c->n_buckets, nullptr, s, true, c->d_fill)
                 : insert_any(c, c->d_tuples, n, nullptr, s, true);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(ei.b, s));
        c->build_ev.push_back(ei);
        c->part_blocks++;
    } else {
        rc = insert_any(c, d_tuples, n, nullptr, s);
        if (rc) return rc;
    }
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

// tuples already in this shard's slice order (fused exchange with by_slice=1): d_slice_offs[n_slices+1] = start of
// every slice's tuples; goes straight to the bucketed insert through L2-resident slices
extern "C" int dbg_insert_sliced_device(dbg_ctx *c, const void *d_tuples, uint64_t n, const uint64_t *d_slice_offs, void *stream)
{
    if (!c || (!d_tuples && n) || !d_slice_offs) return set_err(DBG_ERR_INVALID, "dbg_insert_sliced_device: NULL argument");
    if (c->finalized) return set_err(DBG_ERR_STATE, "insert after finalize");
    if (n == 0) return DBG_OK;
    c->guard_total += n;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    { int trc = table_ready(c, s); if (trc) return trc; }
    CU_TRY(cudaMemcpyAsync(c->d_boffs, d_slice_offs, ((size_t)c->n_buckets + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    EvPair ev;
    int rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    ev.slot = 6;
    rc = insert_any(c, d_tuples, n, nullptr, s, true);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    c->part_blocks++;
    return DBG_OK;
}

// ---- optimistic peer exchange: ONE extraction pass, no counting pass, no offsets exchanged beforehand -----------------
// Every owner keeps a fixed region of `cap_pair` tuples for every source rank in its receive buffer.  The scatter pass of
// the source (k_build<StagedScatterSink<OPT>> with owner buckets) sorts every 2048-tuple batch by owner in shared memory
// and stores the runs straight into the owners' regions over NVLink peer mappings, reserving space with one atomicAdd per
// owner and batch on its LOCAL counters d_fill[0..n_parts); d_fill[n_parts] is raised when a region would overflow (the
// caller then redoes the block with the exact two-pass exchange: nothing of it has been inserted yet).
extern "C" int dbg_exchange_scatter_opt_device(dbg_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads, uint64_t first_base,
                                               uint64_t total_bases, uint64_t first_read_index, int32_t n_parts, void *const *d_dst_ptrs,
                                               uint64_t region_off, uint32_t cap_pair, uint32_t *d_fill, void *stream)
{
    if (!c || !d_dst_ptrs || !d_fill) return set_err(DBG_ERR_INVALID, "dbg_exchange_scatter_opt_device: NULL argument");
    if (n_parts < 1 || n_parts > 64) return set_err(DBG_ERR_INVALID, "n_parts outside 1..64");
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CU_TRY(cudaMemsetAsync(d_fill, 0, ((size_t)n_parts + 1) * sizeof(u32), s));
    if (n_reads == 0 || total_bases == 0) return DBG_OK;
    uint64_t abase = first_base & ~15ull;
    if (((uintptr_t)(d_bases + abase) & 15) != 0) return set_err(DBG_ERR_INVALID, "device base buffer must be 16-byte aligned");
    uint64_t n_chunks = (first_base + total_bases - abase + CB - 1) / CB;
    if (n_chunks > 0x7fffffffull || total_bases >= (1ull << 32)) return set_err(DBG_ERR_INVALID, "block too large for the exchange: split it (< 2^32 bases)");
    int rc = ensure_chunks(c, n_chunks);
    if (rc) return rc;
    const uint32_t cap0 = stage_cap(c, c->wide);
    if (cap0 == 0) return set_err(DBG_ERR_STATE, "staged scatter disabled (DBG_B200_STAGE_CAP=0)");
    // the scatter pass counts reads / occurrences / k-mer-0 lanes as it goes: keep what dbg_exchange_scatter_undo restores
    if (ensure_opt_buffers(c) != DBG_OK) return DBG_ERR_CUDA;
    if ((rc = table_ready(c, s))) return rc;        // (the extraction kernel must not run beside the table clear: see build_device)
    CU_TRY(cudaMemcpyAsync(c->d_snap, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->d_snap + CNT_N, c->d_polyA, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    c->undo_reads = n_reads;
    EvPair ev;
    rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    unsigned gb = (unsigned)((n_reads + 1 + 255) / 256);
    k_chunk_first<<<gb, 256, 0, s>>>((const u64 *)d_offs, n_reads, abase, n_chunks, c->d_chunk_first);
    CU_TRY(cudaGetLastError());
    c->launches++;
    BuildArgs a;
    a.bases = d_bases; a.offs = (const u64 *)d_offs; a.n_reads = n_reads; a.abase = abase; a.end_base = first_base + total_bases;
    a.chunk_first = c->d_chunk_first; a.read_index0 = first_read_index; a.K = c->prm.K; a.R = c->prm.max_read_len;
    a.stage_words = stage_words_for(c->prm.max_read_len); a.count_stats = 1; a.seed = c->prm.payload_mode == 1;
    const uint64_t div = (c->P + n_parts - 1) / n_parts;
    if (c->wide) {
        StagedScatterSink<true, true> st; st.t = view_of(c); st.shift = 0; st.n_buckets = (uint32_t)n_parts; st.cap = cap0; st.matrix = nullptr; st.tuples = nullptr;
        st.fill = d_fill; st.capb = cap_pair; st.flag = d_fill + n_parts; st.filled = 0;
        st.div = div; st.div_M = (uint64_t)((((unsigned __int128)1) << 64) / div); st.dst_ptrs = (u64 *const *)d_dst_ptrs; st.region_off = region_off;
        rc = launch_build<true>(c, a, st, n_chunks, s, 0, StageBuf<true>::bytes(cap0, (uint32_t)n_parts));
    } else {
        StagedScatterSink<false, true> st; st.t = view_of(c); st.shift = 0; st.n_buckets = (uint32_t)n_parts; st.cap = cap0; st.matrix = nullptr; st.tuples = nullptr;
        st.fill = d_fill; st.capb = cap_pair; st.flag = d_fill + n_parts; st.filled = 0;
        st.div = div; st.div_M = (uint64_t)((((unsigned __int128)1) << 64) / div); st.dst_ptrs = (u64 *const *)d_dst_ptrs; st.region_off = region_off;
        rc = launch_build<false>(c, a, st, n_chunks, s, 0, StageBuf<false>::bytes(cap0, (uint32_t)n_parts));
    }
    if (rc) return rc;
    c->reads_total += n_reads;
    ev.slot = 7;                               // ms[7]: the fused scatter-into-peers kernel alone (NVLink figure of bench.py)
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

// PULL exchange, source side: ONE extraction pass partitions this rank's occurrences by (owner, table slice of the owner) into
// its own send buffer d_send (bucket = owner * n_slices + slice, a fixed region of `capb` tuples per bucket, space reserved
// with atomics on d_fill); nothing crosses NVLink here.  d_fill[n_parts * n_slices] != 0: a region would have overflowed.
extern "C" int dbg_exchange_scatter_pull_device(dbg_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads, uint64_t first_base,
                                                uint64_t total_bases, uint64_t first_read_index, int32_t n_parts, void *d_send, uint32_t capb,
                                                uint32_t *d_fill, void *stream)
{
    if (!c || !d_send || !d_fill) return set_err(DBG_ERR_INVALID, "dbg_exchange_scatter_pull_device: NULL argument");
    const uint64_t nbt64 = (uint64_t)(n_parts > 0 ? n_parts : 0) * c->n_buckets;
    if (n_parts < 1 || nbt64 > 4096) return set_err(DBG_ERR_INVALID, "n_parts x table slices = %llu buckets (1..4096)", (unsigned long long)nbt64);
    if (capb < INS_TILE || capb % INS_TILE) return set_err(DBG_ERR_INVALID, "capb must be a multiple of %d", INS_TILE);
    const uint32_t nbt = (uint32_t)nbt64;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CU_TRY(cudaMemsetAsync(d_fill, 0, ((size_t)nbt + 1) * sizeof(u32), s));
    if (n_reads == 0 || total_bases == 0) return DBG_OK;
    uint64_t abase = first_base & ~15ull;
    if (((uintptr_t)(d_bases + abase) & 15) != 0) return set_err(DBG_ERR_INVALID, "device base buffer must be 16-byte aligned");
    uint64_t n_chunks = (first_base + total_bases - abase + CB - 1) / CB;
    if (n_chunks > 0x7fffffffull || total_bases >= (1ull << 32)) return set_err(DBG_ERR_INVALID, "block too large for the exchange: split it (< 2^32 bases)");
    int rc = ensure_chunks(c, n_chunks);
    if (rc) return rc;
    // the batch must fit next to nbt bucket counters in shared memory with three CTAs per SM resident
    uint32_t cap0 = stage_cap(c, c->wide);
    const size_t tb = c->wide ? 32 : 16;
    while (cap0 > (uint32_t)BLOCK * G && cap0 * tb + cap0 * 6 + (size_t)nbt * 12 > 68 * 1024) cap0 /= 2;
    if (cap0 == 0 || cap0 * tb + cap0 * 6 + (size_t)nbt * 12 > 150 * 1024) return set_err(DBG_ERR_STATE, "staged scatter not possible with %u buckets", nbt);
    if (ensure_opt_buffers(c) != DBG_OK) return DBG_ERR_CUDA;
    if ((rc = table_ready(c, s))) return rc;        // (the partition kernel must not run beside the table clear: see build_device)
    CU_TRY(cudaMemcpyAsync(c->d_snap, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->d_snap + CNT_N, c->d_polyA, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    c->undo_reads = n_reads;
    EvPair ev;
    rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    k_chunk_first<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, s>>>((const u64 *)d_offs, n_reads, abase, n_chunks, c->d_chunk_first);
    CU_TRY(cudaGetLastError());
    c->launches++;
    BuildArgs a;
    a.bases = d_bases; a.offs = (const u64 *)d_offs; a.n_reads = n_reads; a.abase = abase; a.end_base = first_base + total_bases;
    a.chunk_first = c->d_chunk_first; a.read_index0 = first_read_index; a.K = c->prm.K; a.R = c->prm.max_read_len;
    a.stage_words = stage_words_for(c->prm.max_read_len); a.count_stats = 1; a.seed = c->prm.payload_mode == 1;
    const uint64_t div = (c->P + n_parts - 1) / n_parts;
    auto fill_sink = [&](auto &st) {
        st.t = view_of(c); st.shift = c->part_shift; st.n_buckets = nbt; st.cap = cap0; st.matrix = nullptr; st.tuples = (u64 *)d_send;
        st.fill = d_fill; st.capb = capb; st.flag = d_fill + nbt; st.filled = 0;
        st.div = div; st.div_M = (uint64_t)((((unsigned __int128)1) << 64) / div); st.dst_ptrs = nullptr; st.region_off = 0; st.nb_local = c->n_buckets;
    };
    if (c->wide) { StagedScatterSink<true, true> st; fill_sink(st); rc = launch_build<true>(c, a, st, n_chunks, s, 0, StageBuf<true>::bytes(cap0, nbt)); }
    else { StagedScatterSink<false, true> st; fill_sink(st); rc = launch_build<false>(c, a, st, n_chunks, s, 0, StageBuf<false>::bytes(cap0, nbt)); }
    if (rc) return rc;
    c->reads_total += n_reads;
    ev.slot = 7;
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

// PULL exchange, owner side: insert this shard's buckets slice by slice, reading the n_src sources' regions over NVLink
// (d_src_ptrs[q] = base of source q's send buffer: a peer mapping, or the local buffer for q == this rank; d_fills =
// the all-gathered fill counters, source q's at d_fills + q * fill_stride).
extern "C" int dbg_insert_pull_device(dbg_ctx *c, void *const *d_src_ptrs, int32_t n_src, uint32_t capb, const uint32_t *d_fills,
                                      uint32_t fill_stride, uint64_t n_tuples_upper, void *stream)
{
    if (!c || !d_src_ptrs || !d_fills) return set_err(DBG_ERR_INVALID, "dbg_insert_pull_device: NULL argument");
    if (c->finalized) return set_err(DBG_ERR_STATE, "insert after finalize");
    if (n_src < 1 || n_src > 64 || capb < INS_TILE || capb % INS_TILE) return set_err(DBG_ERR_INVALID, "dbg_insert_pull_device: bad geometry");
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    { int trc = table_ready(c, s); if (trc) return trc; }
    c->guard_total += n_tuples_upper;
    PullSrc ps;
    ps.ptrs = (const u64 *const *)d_src_ptrs; ps.fills = d_fills; ps.n_src = (u32)n_src; ps.fill_stride = fill_stride;
    ps.region0 = (u32)c->prm.shard_rank * c->n_buckets; ps.capb = capb;
    if (c->n_shards <= 1) ps.region0 = 0;
    const uint64_t tiles = (uint64_t)(capb / INS_TILE) * n_src * c->n_buckets;
    const uint64_t persistent = (uint64_t)c->n_sms * INS_CTAS;
    const unsigned grid = (unsigned)(tiles < persistent ? tiles : persistent);
    EvPair ev, ei;
    int rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    rc = ev_begin(c, s, &ei);
    if (rc) return rc;
    ei.slot = 6;
    CU_TRY(cudaMemsetAsync(c->d_counters + 7, 0, sizeof(u64), s));
    if (c->wide) {
        if (c->track) k_insert_tuples_pull<true, true><<<grid, INS_BLOCK, 0, s>>>(ps, view_of(c), c->n_buckets, c->part_shift, c->d_counters + 7);
        else k_insert_tuples_pull<true, false><<<grid, INS_BLOCK, 0, s>>>(ps, view_of(c), c->n_buckets, c->part_shift, c->d_counters + 7);
    } else {
        if (c->track) k_insert_tuples_pull<false, true><<<grid, INS_BLOCK, 0, s>>>(ps, view_of(c), c->n_buckets, c->part_shift, c->d_counters + 7);
        else k_insert_tuples_pull<false, false><<<grid, INS_BLOCK, 0, s>>>(ps, view_of(c), c->n_buckets, c->part_shift, c->d_counters + 7);
    }
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->part_blocks++;
    CU_TRY(cudaEventRecord(ei.b, s));
    c->build_ev.push_back(ei);
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

// an optimistic scatter overflowed: take back what it added to this context's side counters (reads, occurrences, k-mer-0
// lanes), so that the exact exchange can redo the block
extern "C" int dbg_exchange_scatter_undo(dbg_ctx *c, void *stream)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    if (!c->d_snap) return set_err(DBG_ERR_STATE, "no optimistic scatter to undo");
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CU_TRY(cudaMemcpyAsync(c->d_counters, c->d_snap, CNT_N * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaMemcpyAsync(c->d_polyA, c->d_snap + CNT_N, 8 * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    c->reads_total -= c->undo_reads;
    c->undo_reads = 0;
    return DBG_OK;
}

// Owner side of the optimistic exchange: the receive buffer holds one region per source rank (region r = counts[r]
// tuples at d_base + r * stride_tuples); all of them go through ONE optimistic slice partition and one bucketed insert.
extern "C" int dbg_insert_tuple_regions_device(dbg_ctx *c, const void *d_base, uint32_t n_regions, uint64_t stride_tuples,
                                               const uint64_t *counts, void *stream)
{
    if (!c || !d_base || !counts || n_regions == 0) return set_err(DBG_ERR_INVALID, "dbg_insert_tuple_regions_device: bad argument");
    if (c->finalized) return set_err(DBG_ERR_STATE, "insert after finalize");
    uint64_t n = 0;
    for (uint32_t r = 0; r < n_regions; r++) { if (counts[r] > stride_tuples) return set_err(DBG_ERR_INVALID, "region %u holds more than its stride", r); n += counts[r]; }
    if (n == 0) return DBG_OK;
    c->guard_total += n;
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    { int trc = table_ready(c, s); if (trc) return trc; }
    bool part = want_partition(c, n) && c->optimistic;
    const uint64_t want_cap = n + n / 8 + (uint64_t)c->n_buckets * INS_TILE;
    if (part && ensure_tuples(c, want_cap) != DBG_OK) part = false;
    EvPair ev;
    int rc = ev_begin(c, s, &ev);
    if (rc) return rc;
    int opt = 0;
    uint32_t capb = 0;
    if (part) {
        opt = c->wide ? partition_tuples_optimistic<true>(c, d_base, n_regions, stride_tuples, counts, s, &capb)
                      : partition_tuples_optimistic<false>(c, d_base, n_regions, stride_tuples, counts, s, &capb);
        if (opt < 0) return opt;
    }
    if (opt) {
        EvPair ei;
        rc = ev_begin(c, s, &ei);
        if (rc) return rc;
        ei.slot = 6;
        rc = insert_any(c, c->d_tuples, (uint64_t)capb * c->n_buckets, nullptr, s, true, c->d_fill);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(ei.b, s));
        c->build_ev.push_back(ei);
        c->part_blocks++;
    } else {
        // few tuples per table slice, or a slice region overflowed (skew): insert the regions as they are
        const size_t tb = c->wide ? 32 : 16;
        for (uint32_t r = 0; r < n_regions; r++) {
            if (counts[r] == 0) continue;
            rc = insert_any(c, static_cast<const char *>(d_base) + (size_t)r * stride_tuples * tb, counts[r], nullptr, s);
            if (rc) return rc;
        }
        c->path_counts[0]++;
    }
    CU_TRY(cudaEventRecord(ev.b, s));
    c->build_ev.push_back(ev);
    return DBG_OK;
}

extern "C" int dbg_partition_info(const dbg_ctx *c, uint32_t *n_slices, int32_t *slice_shift)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    if (n_slices) *n_slices = c->n_buckets;
    if (slice_shift) *slice_shift = c->part_shift;
    return DBG_OK;
}

extern "C" int dbg_get_polyA_counts(dbg_ctx *c, uint64_t counts[8])
{
    if (!c || !counts) return set_err(DBG_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    int frc = flush_batch(c);
    if (frc) return frc;
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(counts, c->d_polyA, 8 * sizeof(u64), cudaMemcpyDeviceToHost));
    return DBG_OK;
}

extern "C" int dbg_set_polyA_counts(dbg_ctx *c, const uint64_t counts[8])
{
    if (!c || !counts) return set_err(DBG_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(c->d_polyA, counts, 8 * sizeof(u64), cudaMemcpyHostToDevice));
    return DBG_OK;
}

// ---------------------------------------------------------------------------------------------------
// finalize: reference slot layout + k-mer-0 node
// ---------------------------------------------------------------------------------------------------
static uint64_t nul_words(uint64_t P) { return (P / 8 + 1 + 3) / 4 + 1; }

static int fill_stats(dbg_ctx *c, const u64 *cnt)
{
    dbg_stats &s = c->st;
    s.array_size = c->P; s.max_cutoff = c->max_cutoff; s.load_factor = c->lf; s.wide = c->wide;
    s.count = cnt[CNT_NEW] + (c->finalized && (c->n_shards <= 1) && c->prm.payload_mode != 1 ? 1 : 0);
    s.conflict = cnt[CNT_CONFLICT]; s.reads = c->reads_total; s.kmers_logged = cnt[CNT_LOGGED]; s.occurrences = cnt[CNT_OCC];
    s.polyA_l = (uint32_t)c->polyA_links; s.polyA_r = (uint32_t)(c->polyA_links >> 32);
    s.shard_lo = c->shard_lo; s.shard_hi = c->shard_hi;
    return DBG_OK;
}

static int read_counters(dbg_ctx *c, u64 *cnt)
{
    int frc = flush_batch(c);
    if (frc) return frc;
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(cnt, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToHost));
    if (cnt[CNT_ERROR] == 1) return set_err(DBG_ERR_TABLE_FULL, "probe ran off the shard (%llu slots + margin): table too full", (unsigned long long)(c->shard_hi - c->shard_lo));
    return DBG_OK;
}

static const uint64_t SCRATCH_CAP = 16ull << 20;    // entries of the region scratch (long clusters + wrap region)

// global method (fallback): atomicMin priority probing over the whole table, needs owner[P]
template <bool WIDE, bool TRACK>
static int run_layout_global(dbg_ctx *c)
{
    if (c->owner_cap < c->P) {
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(c->d_owner); c->d_owner = nullptr; c->owner_cap = 0;
        CU_TRY(cudaMalloc(&c->d_owner, c->P * sizeof(u64)));
        c->owner_cap = c->P;
    }
    size_t nb = (size_t)node_bytes(c);
    CU_TRY(cudaMemsetAsync(c->d_owner, 0xFF, c->P * sizeof(u64), c->stream));
    CU_TRY(cudaMemsetAsync(c->d_out, 0, c->P * nb, c->stream));
    CU_TRY(cudaMemsetAsync(c->d_nul32, 0, nul_words(c->P) * sizeof(u32), c->stream));
    unsigned grid = c->n_sms * 16;
    k_layout_insert<WIDE, TRACK><<<grid, 256, 0, c->stream>>>((const NodeT<WIDE> *)c->d_nodes, c->n_local, c->shard_lo, c->d_owner, c->P, c->M);
    CU_TRY(cudaGetLastError());
    k_layout_place<WIDE, TRACK><<<grid, 256, 0, c->stream>>>((const NodeT<WIDE> *)c->d_nodes, c->n_local, c->shard_lo, c->d_owner, c->P, c->M, c->d_out, c->d_nul32);
    CU_TRY(cudaGetLastError());
    c->launches += 2;
    return DBG_OK;
}

// cluster-local method: one streaming pass, regions (long clusters, wrap-around) on a small scratch
static LayoutGeom layout_geom(const dbg_ctx *c)
{
    LayoutGeom g;
    g.P = c->P; g.M = c->M;
    if (c->n_shards <= 1) { g.gbase = 0; g.v_begin = 0; g.v_end = c->P; }
    else {
        // window = [imported tail cluster (end of the porch) | own home range minus the own tail cluster]
        g.gbase = (c->shard_lo + c->P - c->porch % c->P) % c->P;
        g.v_begin = c->porch - c->tail_a_in;
        g.v_end = c->porch + ((c->shard_hi - c->shard_lo) - c->tail_a_own);
    }
    g.v_first = g.v_begin; g.v_halo = g.v_end; g.v_final = g.v_end;
    return g;
}

// the cluster-local layout pass over one window: version 2 (warp-owned word pairs, default) or 1 (block-wide renumbering;
// env DBG_B200_LAYOUT_V=1, kept for A/B measurements)
template <bool WIDE, bool TRACK>
static void launch_layout_clusters(dbg_ctx *c, const NodeT<WIDE> *nodes, const LayoutGeom &geo, u32 *nul32, LayoutInfo *info, LayoutRegion *regions, cudaStream_t s)
{
    if (c->layout_v == 1)
        k_layout_clusters<WIDE, TRACK><<<c->n_sms * (2048 / LT), LT, 0, s>>>(nodes, geo, c->d_out, nul32, info, regions, SCRATCH_CAP);
    else
        k_layout_clusters2<WIDE, TRACK><<<c->n_sms * (2048 / L2N), L2N, 0, s>>>(nodes, geo, c->d_out, nul32, info, regions, SCRATCH_CAP);
}

static int ensure_layout_scratch(dbg_ctx *c)
{
    if (c->owner_cap < SCRATCH_CAP) {
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(c->d_owner); c->d_owner = nullptr; c->owner_cap = 0;
        CU_TRY(cudaMalloc(&c->d_owner, SCRATCH_CAP * sizeof(u64)));
        c->owner_cap = SCRATCH_CAP;
    }
    if (!c->d_layout_info) {
        CU_TRY(cudaMalloc(&c->d_layout_info, sizeof(LayoutInfo)));
        CU_TRY(cudaMalloc(&c->d_regions, (size_t)MAX_REGIONS * sizeof(LayoutRegion)));
    }
    return DBG_OK;
}

template <bool WIDE, bool TRACK>
static int run_layout(dbg_ctx *c)
{
    const NodeT<WIDE> *nodes = (const NodeT<WIDE> *)c->d_nodes;
    bool use_global = c->layout_mode == 1;
    if (!use_global) {
        int rc = ensure_layout_scratch(c);
        if (rc) return rc;
        const LayoutGeom geo = layout_geom(c);
        CU_TRY(cudaMemsetAsync(c->d_nul32, 0, nul_words(c->P) * sizeof(u32), c->stream));
        k_layout_wrapscan<WIDE><<<1, 32, 0, c->stream>>>(nodes, c->n_local, c->P, c->d_layout_info, c->d_regions, SCRATCH_CAP);
        CU_TRY(cudaGetLastError());
        launch_layout_clusters<WIDE, TRACK>(c, nodes, geo, c->d_nul32, c->d_layout_info, c->d_regions, c->stream);
        CU_TRY(cudaGetLastError());
        k_layout_regions<WIDE, TRACK><<<c->n_sms * 4, 256, 0, c->stream>>>(nodes, geo, c->d_out, c->d_nul32, c->d_layout_info,
                                                                            c->d_regions, c->d_owner);
        CU_TRY(cudaGetLastError());
        c->launches += 3;
        LayoutInfo li;
        CU_TRY(cudaMemcpyAsync(&li, c->d_layout_info, sizeof(li), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        c->layout_regions = li.n_regions;
        if (li.overflow) use_global = true;      // very dense table: too many / too long clusters for the scratch
    }
    if (use_global) {
        int rc = run_layout_global<WIDE, TRACK>(c);
        if (rc) return rc;
        c->layout_regions = UINT32_MAX;
    }
    // (seed index, payload_mode 1: k-mer 0 is an ordinary key there -- csrc/seedidx.cu places it by its first occurrence)
    if (c->prm.payload_mode != 1) {
        k_polyA_insert<WIDE><<<1, 32, 0, c->stream>>>(c->P, c->M, c->d_polyA, c->d_out, c->d_nul32, c->d_counters + 6);
        CU_TRY(cudaGetLastError());
        c->launches++;
    }
    return DBG_OK;
}

// sharded context: lay out this rank's window [imported tail | own range minus own tail] (dbg_shard_tail_export /
// _import came first) into d_out (virtual slot order) and derive the occupancy words of the slice in global positions.
// The k-mer-0 node is NOT placed here: it goes in last, on the merged table (dbg_host_polyA_insert).
template <bool WIDE, bool TRACK>
static int run_layout_sharded(dbg_ctx *c)
{
    const NodeT<WIDE> *base = (const NodeT<WIDE> *)c->d_nodes_alloc;       // virtual slot 0
    int rc = ensure_layout_scratch(c);
    if (rc) return rc;
    const LayoutGeom geo = layout_geom(c);
    LayoutInfo li0;
    memset(&li0, 0, sizeof(li0));
    li0.e = ~0ULL;                                                           // no wrap region inside a shard window
    CU_TRY(cudaMemcpyAsync(c->d_layout_info, &li0, sizeof(li0), cudaMemcpyHostToDevice, c->stream));
    if (geo.v_end > geo.v_begin) {
        launch_layout_clusters<WIDE, TRACK>(c, base, geo, nullptr, c->d_layout_info, c->d_regions, c->stream);
        CU_TRY(cudaGetLastError());
        k_layout_regions<WIDE, TRACK><<<c->n_sms * 4, 256, 0, c->stream>>>(base, geo, c->d_out, nullptr, c->d_layout_info, c->d_regions, c->d_owner);
        CU_TRY(cudaGetLastError());
        c->launches += 2;
    }
    LayoutInfo li;
    CU_TRY(cudaMemcpyAsync(&li, c->d_layout_info, sizeof(li), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->layout_regions = li.n_regions;
    if (li.overflow) return set_err(DBG_ERR_STATE, "shard too dense for the windowed layout (cluster scratch exhausted): merge the shard dumps instead");
    // occupancy words of the slice, in global word positions (a slice that wraps past slot P-1 produces two runs)
    const uint64_t n = geo.v_end - geo.v_begin;
    const uint64_t g_first = (geo.gbase + geo.v_begin) % c->P;
    const uint64_t words = n ? (n + 31) / 32 + 2 : 0;
    if (words > c->nul_slice_words) {
        cudaFree(c->d_nul_slice); c->d_nul_slice = nullptr; c->nul_slice_words = 0;
        CU_TRY(cudaMalloc(&c->d_nul_slice, 2 * (words + 2) * sizeof(u32)));
        c->nul_slice_words = words;
    }
    if (n) {
        // run 1: words from g_first/32 up to the end of the slice or of the table; run 2 (wrap): words from 0
        const uint64_t first_len = g_first + n <= c->P ? n : c->P - g_first;
        const uint64_t w1 = g_first / 32, nw1 = (g_first + first_len + 31) / 32 - w1;
        k_nul_from_image<WIDE><<<(unsigned)((nw1 * 32 + 255) / 256), 256, 0, c->stream>>>(c->d_out, c->P, geo.gbase, g_first, n, w1, nw1, c->d_nul_slice);
        CU_TRY(cudaGetLastError());
        c->launches++;
        if (first_len < n) {
            const uint64_t nw2 = (n - first_len + 31) / 32;
            k_nul_from_image<WIDE><<<(unsigned)((nw2 * 32 + 255) / 256), 256, 0, c->stream>>>(c->d_out, c->P, geo.gbase, g_first, n, 0, nw2,
                                                                                          c->d_nul_slice + c->nul_slice_words + 2);
            CU_TRY(cudaGetLastError());
            c->launches++;
        }
    }
    return DBG_OK;
}

// ---------------------------------------------------------------------------------------------------
// dbg_finish_export: the last block's insert, the reference layout and the copy of the image, slice group by slice group.
// The tuples of the block are bucket-ordered (bucket = table slice), so after the insert of the buckets [0, b) every slot
// below b << part_shift is FINAL (later keys have later homes and linear probing only moves forward).  Group g therefore
// runs  insert(buckets of g) -> k_layout_clusters(window of g) -> D2H(window of g)  with the copy on its own stream: the
// image is on the PCIe link -- the bound of the whole hand-over -- while later groups are still being inserted.
// Windows end GUARD slots before the group's last slot so that tiles read their halo, and long clusters are measured,
// inside final territory.  The wrap-around region, clusters longer than the shared-memory window and the k-mer-0 node are
// settled after the last group and their (few, small) slot ranges are copied again.
// Returns 0 = image delivered and context finalized; 1 = not applicable, nothing touched; 2 = table built, but the layout
// must be redone in one pass (dense table: too many long clusters) -- the caller goes through dbg_finalize / dbg_export.
// ---------------------------------------------------------------------------------------------------
template <bool WIDE, bool TRACK>
static int grouped_finish(dbg_ctx *c, uint32_t capb, FinishExport *fx)
{
    const uint32_t nb = c->n_buckets;
    const uint64_t bucket_slots = 1ull << c->part_shift;
    const uint64_t GUARD = MARGIN_SLOTS;
    const uint64_t P = c->P;
    const size_t nbytes = (size_t)node_bytes(c);
    if (c->n_shards > 1 || c->layout_mode != 0 || c->prm.payload_mode != 0) return 1;
    int G = getenv("DBG_B200_FINISH_GROUPS") ? atoi(getenv("DBG_B200_FINISH_GROUPS")) : 8;
    if (G < 2) return 1;
    uint32_t per = (nb + G - 1) / G;
    if (per == 0 || (uint64_t)per * bucket_slots < 4 * GUARD || P < 8 * GUARD) return 1;
    cudaStream_t s = c->stream, cs = c->copy_stream;
    int rc = ensure_layout_scratch(c);
    if (rc) return rc < 0 ? rc : -1;
    if (!c->d_out) {
        CU_TRY(cudaMalloc(&c->d_out, P * nbytes));
        CU_TRY(cudaMalloc(&c->d_nul32, nul_words(P) * sizeof(u32)));
    }
    if (!c->d_layout_info2) {
        CU_TRY(cudaMalloc(&c->d_layout_info2, sizeof(LayoutInfo)));
        CU_TRY(cudaMalloc(&c->d_regions2, 4 * sizeof(LayoutRegion)));
        CU_TRY(cudaMallocHost(&c->h_regions, 4096 * sizeof(LayoutRegion)));
    }
    const NodeT<WIDE> *nodes = (const NodeT<WIDE> *)c->d_nodes;
    char *array = static_cast<char *>(fx->array);
    uint8_t *nul_flag = fx->nul_flag;
    uint64_t link_bytes = 0, n_windows = 0;
    auto copy_range = [&](uint64_t s0, uint64_t s1, bool to_end) -> cudaError_t {
        if (s1 <= s0) return cudaSuccess;
        cudaError_t e = cudaMemcpyAsync(array + s0 * nbytes, static_cast<const char *>(c->d_out) + s0 * nbytes, (s1 - s0) * nbytes, cudaMemcpyDeviceToHost, cs);
        if (e != cudaSuccess) return e;
        const uint64_t by0 = s0 / 8, by1 = to_end ? P / 8 + 1 : (s1 + 7) / 8;
        link_bytes += (s1 - s0) * nbytes + (by1 - by0);
        return cudaMemcpyAsync(nul_flag + by0, reinterpret_cast<const uint8_t *>(c->d_nul32) + by0, by1 - by0, cudaMemcpyDeviceToHost, cs);
    };
    CU_TRY(cudaMemsetAsync(c->d_nul32, 0, nul_words(P) * sizeof(u32), s));
    LayoutInfo li0;
    memset(&li0, 0, sizeof(li0));
    li0.e = ~0ULL;                       // no wrap region while the groups run: it is settled at the end
    CU_TRY(cudaMemcpyAsync(c->d_layout_info, &li0, sizeof(li0), cudaMemcpyHostToDevice, s));
    EvPair ev_d2h;
    bool d2h_started = false;
    std::vector<cudaEvent_t> gev;
    uint64_t win_begin = 0, lastw = 0;       // lastw: first slot of the last window
    for (uint32_t b0 = 0; b0 < nb;) {
        uint32_t b1 = b0 + per < nb ? b0 + per : nb;
        uint64_t end_g = (uint64_t)b1 * bucket_slots;
        if (end_g >= P + GUARD || end_g - GUARD >= P || end_g - GUARD <= win_begin + GUARD) { b1 = nb; end_g = c->n_local; }   // the rest is the last group
        const bool last = b1 == nb;
        EvPair evi;
        rc = ev_begin(c, s, &evi);
        if (rc) return rc;
        evi.slot = 6;
        rc = insert_any(c, c->d_tuples, (uint64_t)capb * b1, nullptr, s, true, c->d_fill, (uint64_t)capb * b0, b0);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(evi.b, s));
        c->build_ev.push_back(evi);
        LayoutGeom geo;
        geo.P = P; geo.M = c->M; geo.gbase = 0; geo.v_first = 0; geo.v_halo = P;
        geo.v_begin = win_begin; geo.v_end = last ? P : end_g - GUARD; geo.v_final = last ? P : end_g;
        EvPair evl;
        rc = ev_begin(c, s, &evl);
        if (rc) return rc;
        evl.slot = 2;
        launch_layout_clusters<WIDE, TRACK>(c, nodes, geo, c->d_nul32, c->d_layout_info, c->d_regions, s);
        CU_TRY(cudaGetLastError());
        c->launches++;
        CU_TRY(cudaEventRecord(evl.b, s));
        c->build_ev.push_back(evl);
        if (last) lastw = geo.v_begin;
        else {
            cudaEvent_t e;
            rc = ev_get(c, &e);
            if (rc) return rc;
            gev.push_back(e);
            CU_TRY(cudaEventRecord(e, s));
            CU_TRY(cudaStreamWaitEvent(cs, e, 0));
            if (!d2h_started) { rc = ev_begin(c, cs, &ev_d2h); if (rc) return rc; d2h_started = true; }
            CU_TRY(copy_range(geo.v_begin, geo.v_end, false));
            n_windows++;
        }
        win_begin = geo.v_end;
        b0 = b1;
    }
    // ---- after the last insert: wrap-around region, long clusters, k-mer-0 node ----
    EvPair evt;
    rc = ev_begin(c, s, &evt);
    if (rc) return rc;
    evt.slot = 2;
    k_layout_wrapscan<WIDE><<<1, 32, 0, s>>>(nodes, c->n_local, P, c->d_layout_info2, c->d_regions2, SCRATCH_CAP);
    CU_TRY(cudaGetLastError());
    LayoutGeom whole;
    whole.P = P; whole.M = c->M; whole.gbase = 0; whole.v_begin = 0; whole.v_end = P; whole.v_first = 0; whole.v_halo = P; whole.v_final = P;
    k_layout_regions<WIDE, TRACK><<<c->n_sms * 4, 256, 0, s>>>(nodes, whole, c->d_out, c->d_nul32, c->d_layout_info, c->d_regions, c->d_owner);
    CU_TRY(cudaGetLastError());
    k_layout_regions<WIDE, TRACK><<<1, 256, 0, s>>>(nodes, whole, c->d_out, c->d_nul32, c->d_layout_info2, c->d_regions2, c->d_owner);
    CU_TRY(cudaGetLastError());
    k_polyA_insert<WIDE><<<1, 32, 0, s>>>(P, c->M, c->d_polyA, c->d_out, c->d_nul32, c->d_counters + 6);
    CU_TRY(cudaGetLastError());
    c->launches += 4;
    CU_TRY(cudaEventRecord(evt.b, s));
    c->build_ev.push_back(evt);
    LayoutInfo li1, li2;
    u64 cnt[CNT_N];
    CU_TRY(cudaMemcpyAsync(&li1, c->d_layout_info, sizeof(li1), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&li2, c->d_layout_info2, sizeof(li2), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(cnt, c->d_counters, CNT_N * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    for (cudaEvent_t e : gev) c->ev_free_list.push_back(e);
    auto give_up = [&]() -> int {
        // the table is complete but this layout is not usable: let the copies in flight drain, then the one-pass path
        cudaStreamSynchronize(cs);
        if (d2h_started) ev_put(c, ev_d2h);
        return 2;
    };
    if (cnt[CNT_ERROR] == 1 || cnt[CNT_NEW] + 1 > P) return give_up();          // (dbg_finalize reports DBG_ERR_TABLE_FULL)
    const uint32_t MAXR = 4096;
    if (li1.overflow || li2.overflow || li1.n_regions > MAXR) return give_up();
    if (li1.n_regions) CU_TRY(cudaMemcpyAsync(c->h_regions, c->d_regions, li1.n_regions * sizeof(LayoutRegion), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // the last window, then the ranges that changed after their window had left
    if (!d2h_started) { rc = ev_begin(c, cs, &ev_d2h); if (rc) return rc; d2h_started = true; }
    CU_TRY(copy_range(lastw, P, true));
    n_windows++;
    for (uint32_t r = 0; r < li1.n_regions; r++) {
        const LayoutRegion &rg = c->h_regions[r];
        if (rg.a < lastw) CU_TRY(copy_range(rg.a, rg.a + rg.n < lastw ? rg.a + rg.n : lastw, false));
    }
    if (li2.n_regions && li2.g > 0) CU_TRY(copy_range(0, li2.g < lastw ? li2.g : lastw, false));      // ([e, P) lies in the last window or is patched here:)
    if (li2.n_regions && li2.e < lastw) CU_TRY(copy_range(li2.e, lastw, false));
    CU_TRY(cudaEventRecord(ev_d2h.b, cs));
    CU_TRY(cudaEventSynchronize(ev_d2h.b));
    CU_TRY(cudaEventElapsedTime(&c->ms[5], ev_d2h.a, ev_d2h.b));
    ev_put(c, ev_d2h);
    CU_TRY(cudaMemcpy(&c->polyA_links, c->d_counters + 6, sizeof(u64), cudaMemcpyDeviceToHost));
    // the k-mer-0 node went into the device image after its window had left: the same insertion on the host copy
    // (the windows carried the image WITHOUT it, except the last one -- dbg_host_polyA_insert is idempotent only on a table
    // that lacks the node, so take it out of the last window's bits first if it landed there)
    {
        const uint64_t home0 = (WIDE ? hash_code_wide(0, 0) : hash_code(0)) % P;
        // find where the device put it: first slot from home0 whose image k-mer is 0 while its bit is set
        uint64_t sl = home0;
        for (;;) {
            const bool bit = (nul_flag[sl >> 3] >> (7 - (sl & 7))) & 1;
            if (!bit) break;                                              // free slot: the node is not in the host copy yet
            const uint64_t *nd = reinterpret_cast<const uint64_t *>(array + sl * nbytes);
            if (nd[0] == 0 && (!WIDE || nd[1] == 0)) { sl = ~0ULL; break; }  // already there (it travelled with its window)
            sl = sl + 1 == P ? 0 : sl + 1;
        }
        if (sl != ~0ULL) {
            uint64_t slot_out = 0;
            int hrc = dbg_host_polyA_insert(array, nul_flag, P, WIDE ? 1 : 0, (uint32_t)(c->polyA_links & 0xffffffffu), (uint32_t)(c->polyA_links >> 32), &slot_out);
            if (hrc) return hrc;
        }
    }
    for (auto &e : c->build_ev) { float t = 0; CU_TRY(cudaEventElapsedTime(&t, e.a, e.b)); c->ms[e.slot] += t; ev_put(c, e); }
    c->build_ev.clear();
    c->layout_regions = li1.n_regions + li2.n_regions;
    c->finalized = true;
    c->export_info[0] = 0; c->export_info[1] = n_windows; c->export_info[2] = link_bytes; c->export_info[3] = cnt[CNT_NEW] + 1;
    fx->done = true;
    return 0;
}

extern "C" int dbg_finalize(dbg_ctx *c, dbg_stats *stats)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    CU_TRY(cudaSetDevice(c->device));
    u64 cnt[CNT_N];
    int rc = read_counters(c, cnt);
    if (rc) return rc;
    if ((rc = table_ready(c, c->stream))) return rc;
    collect_clear_time(c);
    if (!c->finalized) {
        // build time = sum over blocks
        for (auto &e : c->build_ev) { float t = 0; CU_TRY(cudaEventElapsedTime(&t, e.a, e.b)); c->ms[e.slot] += t; ev_put(c, e); }
        c->build_ev.clear();
        const bool sharded_layout = c->n_shards > 1 && c->tail_exported && c->tail_imported;
        if (c->n_shards <= 1 || sharded_layout) {
            if (c->n_shards <= 1 && cnt[CNT_NEW] + 1 > c->P) return set_err(DBG_ERR_TABLE_FULL, "%llu nodes do not fit %llu slots", (unsigned long long)cnt[CNT_NEW] + 1, (unsigned long long)c->P);
            size_t nb = (size_t)node_bytes(c);
            if (!c->d_out) {
                // unsharded: the whole image (global slot order); sharded: the window [porch | own range] (virtual order)
                const uint64_t slots = c->n_shards <= 1 ? c->P : c->porch + (c->shard_hi - c->shard_lo) + 1;
                CU_TRY(cudaMalloc(&c->d_out, slots * nb));
                if (c->n_shards <= 1) CU_TRY(cudaMalloc(&c->d_nul32, nul_words(c->P) * sizeof(u32)));
            }
            EvPair e;
            rc = ev_begin(c, c->stream, &e);
            if (rc) return rc;
            if (sharded_layout) {
                if (c->wide) rc = c->track ? run_layout_sharded<true, true>(c) : run_layout_sharded<true, false>(c);
                else rc = c->track ? run_layout_sharded<false, true>(c) : run_layout_sharded<false, false>(c);
            } else {
                if (c->wide) rc = c->track ? run_layout<true, true>(c) : run_layout<true, false>(c);
                else rc = c->track ? run_layout<false, true>(c) : run_layout<false, false>(c);
            }
            if (rc) return rc;
            CU_TRY(cudaEventRecord(e.b, c->stream));
            CU_TRY(cudaEventSynchronize(e.b));
            CU_TRY(cudaEventElapsedTime(&c->ms[2], e.a, e.b));
            ev_put(c, e);
            if (!sharded_layout) CU_TRY(cudaMemcpy(&c->polyA_links, c->d_counters + 6, sizeof(u64), cudaMemcpyDeviceToHost));
            else {
                // the k-mer-0 node's link words from the (all-reduced) side counters; it is placed on the merged table
                u64 pa[8];
                CU_TRY(cudaMemcpy(pa, c->d_polyA, sizeof(pa), cudaMemcpyDeviceToHost));
                u64 l = 0, r = 0;
                for (int b = 0; b < 4; b++) { l |= (pa[b] > 255 ? 255 : pa[b]) << (24 - 8 * b); r |= (pa[4 + b] > 255 ? 255 : pa[4 + b]) << (24 - 8 * b); }
                c->polyA_links = l | (r << 32);
            }
        }
        c->finalized = true;
    }
    fill_stats(c, cnt);
    if (stats) *stats = c->st;
    return DBG_OK;
}

// ---------------------------------------------------------------------------------------------------
// cross-shard reference layout: tail hand-off, slice export, host-side merge helpers
// ---------------------------------------------------------------------------------------------------
extern "C" int dbg_shard_tail_export(dbg_ctx *c, void *blob, uint64_t cap_bytes, uint64_t *n_bytes)
{
    if (!c || !n_bytes) return set_err(DBG_ERR_INVALID, "NULL argument");
    { cudaSetDevice(c->device); int trc = table_ready(c, c->stream); if (trc) return trc; }
    if (c->n_shards <= 1) return set_err(DBG_ERR_STATE, "dbg_shard_tail_export: unsharded context");
    if (c->finalized) return set_err(DBG_ERR_STATE, "dbg_shard_tail_export after finalize");
    CU_TRY(cudaSetDevice(c->device));
    u64 cnt[CNT_N];
    int rc = read_counters(c, cnt);        // waits for the inserts; reports a shard that overflowed
    if (rc) return rc;
    const size_t nbb = build_node_bytes(c);
    const uint64_t own = c->shard_hi - c->shard_lo;
    if (!c->tail_exported) {
        if (!c->d_tail) CU_TRY(cudaMalloc(&c->d_tail, sizeof(TailInfo)));
        CU_TRY(cudaMemsetAsync(c->d_tail, 0, sizeof(TailInfo), c->stream));
        // the receiver keeps the in-range part in its porch: a run longer than the porch cannot be handed over
        const uint64_t max_a = c->porch;
        if (c->wide) k_shard_tail_scan<true><<<1, 32, 0, c->stream>>>((const NodeT<true> *)c->d_nodes, own, c->n_local, max_a, c->d_tail);
        else k_shard_tail_scan<false><<<1, 32, 0, c->stream>>>((const NodeT<false> *)c->d_nodes, own, c->n_local, max_a, c->d_tail);
        CU_TRY(cudaGetLastError());
        c->launches++;
        TailInfo ti;
        CU_TRY(cudaMemcpyAsync(&ti, c->d_tail, sizeof(ti), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (ti.a > max_a || (own && ti.a >= own)) return set_err(DBG_ERR_STATE, "boundary cluster of %llu slots does not fit the hand-off porch (%llu): merge the shard dumps instead",
                                                         (unsigned long long)ti.a, (unsigned long long)max_a);
        c->tail_a_own = ti.a; c->tail_mt_own = ti.mt;
        c->tail_exported = true;
    }
    const uint64_t n_nodes = c->tail_a_own + c->tail_mt_own;
    const uint64_t need = 4 * sizeof(uint64_t) + n_nodes * nbb;
    *n_bytes = need;
    if (!blob) return DBG_OK;                                    // size query
    if (cap_bytes < need) return set_err(DBG_ERR_BUFFER, "tail blob needs %llu bytes", (unsigned long long)need);
    uint64_t hdr[4] = {c->tail_a_own, c->tail_mt_own, (uint64_t)nbb, (uint64_t)c->prm.shard_rank};
    memcpy(blob, hdr, sizeof(hdr));
    if (n_nodes) {
        const char *src = static_cast<const char *>(c->d_nodes) + (own - c->tail_a_own) * nbb;      // [tail run | margin run] is contiguous
        CU_TRY(cudaMemcpyAsync(static_cast<char *>(blob) + sizeof(hdr), src, n_nodes * nbb, cudaMemcpyDeviceToHost, c->stream));
        // the margin nodes now belong to the right neighbour (their slots are in its home range)
        if (c->tail_mt_own && !c->tail_moved) {
            CU_TRY(cudaMemsetAsync(static_cast<char *>(c->d_nodes) + own * nbb, 0, c->tail_mt_own * nbb, c->stream));
            c->nodes_delta -= (int64_t)c->tail_mt_own;
            c->tail_moved = true;
        }
        CU_TRY(cudaStreamSynchronize(c->stream));
    }
    return DBG_OK;
}

extern "C" int dbg_shard_tail_import(dbg_ctx *c, const void *blob, uint64_t n_bytes)
{
    if (!c || !blob) return set_err(DBG_ERR_INVALID, "NULL argument");
    { cudaSetDevice(c->device); int trc = table_ready(c, c->stream); if (trc) return trc; }
    if (c->n_shards <= 1) return set_err(DBG_ERR_STATE, "dbg_shard_tail_import: unsharded context");
    if (!c->tail_exported) return set_err(DBG_ERR_STATE, "dbg_shard_tail_import before dbg_shard_tail_export (the own tail is fixed first)");
    if (c->tail_imported) return set_err(DBG_ERR_STATE, "tail already imported");
    CU_TRY(cudaSetDevice(c->device));
    const size_t nbb = build_node_bytes(c);
    uint64_t hdr[4];
    if (n_bytes < sizeof(hdr)) return set_err(DBG_ERR_INVALID, "short tail blob");
    memcpy(hdr, blob, sizeof(hdr));
    const uint64_t a = hdr[0], mt = hdr[1];
    if (hdr[2] != nbb || n_bytes < sizeof(hdr) + (a + mt) * nbb) return set_err(DBG_ERR_INVALID, "tail blob does not match this context");
    if (a > c->porch) return set_err(DBG_ERR_STATE, "incoming boundary cluster exceeds the porch");
    const char *nodes_h = static_cast<const char *>(blob) + sizeof(hdr);
    const uint64_t own = c->shard_hi - c->shard_lo;
    // in-range part: verbatim into the end of the porch (same relative slots: they end at the shard boundary)
    if (a) CU_TRY(cudaMemcpyAsync(static_cast<char *>(c->d_nodes) - a * nbb, nodes_h, a * nbb, cudaMemcpyHostToDevice, c->stream));
    if (mt) {
        if (mt > c->tail_nodes_cap) {
            cudaFree(c->d_tail_nodes); c->d_tail_nodes = nullptr; c->tail_nodes_cap = 0;
            CU_TRY(cudaMalloc(&c->d_tail_nodes, mt * nbb));
            c->tail_nodes_cap = mt;
        }
        CU_TRY(cudaMemcpyAsync(c->d_tail_nodes, nodes_h + a * nbb, mt * nbb, cudaMemcpyHostToDevice, c->stream));
        if (c->wide) k_shard_adopt<true><<<1, 32, 0, c->stream>>>((NodeT<true> *)c->d_nodes, c->n_local, (const NodeT<true> *)c->d_tail_nodes, mt, c->d_tail);
        else k_shard_adopt<false><<<1, 32, 0, c->stream>>>((NodeT<false> *)c->d_nodes, c->n_local, (const NodeT<false> *)c->d_tail_nodes, mt, c->d_tail);
        CU_TRY(cudaGetLastError());
        c->launches++;
        TailInfo ti;
        CU_TRY(cudaMemcpyAsync(&ti, c->d_tail, sizeof(ti), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        c->nodes_delta += (int64_t)ti.adopted;
        // the adopted chain must end before this rank's own tail cluster starts (otherwise the hand-off cascades
        // through the whole shard: tiny tables only) -- the caller then merges the shard dumps instead
        if (ti.landing_max == ~0ULL || ti.landing_max + 1 >= own - c->tail_a_own)
            return set_err(DBG_ERR_STATE, "boundary hand-off cascades through the whole shard: merge the shard dumps instead");
    }
    CU_TRY(cudaStreamSynchronize(c->stream));      // the blob may go away
    c->tail_a_in = a;
    c->tail_imported = true;
    return DBG_OK;
}

extern "C" int dbg_shard_slice_info(dbg_ctx *c, uint64_t *g_first, uint64_t *n_slots, void **d_slice)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    if (c->n_shards <= 1 || !c->finalized || !c->d_out || !c->tail_imported) return set_err(DBG_ERR_STATE, "no laid-out shard slice (tail export/import + finalize first)");
    const LayoutGeom g = layout_geom(c);
    if (g_first) *g_first = (g.gbase + g.v_begin) % c->P;
    if (n_slots) *n_slots = g.v_end - g.v_begin;
    if (d_slice) *d_slice = static_cast<char *>(c->d_out) + g.v_begin * (size_t)node_bytes(c);
    return DBG_OK;
}

// copy this rank's slice into the FULL host table: array[P] (reference node size) and the nul_flag bytes this slice
// covers completely; the (at most two) bytes it shares with its neighbours are returned in edge_slots[4] (UINT64_MAX =
// unused) for dbg_host_fix_nul_bytes once every rank has exported
extern "C" int dbg_export_shard_slice(dbg_ctx *c, void *array, uint8_t *nul_flag, uint64_t edge_slots[4])
{
    if (!c || !array) return set_err(DBG_ERR_INVALID, "NULL argument");
    uint64_t g_first = 0, n = 0;
    void *d_slice = nullptr;
    int rc = dbg_shard_slice_info(c, &g_first, &n, &d_slice);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(c->device));
    if (edge_slots) for (int i = 0; i < 4; i++) edge_slots[i] = UINT64_MAX;
    if (n == 0) return DBG_OK;
    EvPair e;
    rc = ev_begin(c, c->stream, &e);
    if (rc) return rc;
    const size_t nb = (size_t)node_bytes(c);
    const uint64_t first_len = g_first + n <= c->P ? n : c->P - g_first;
    CU_TRY(cudaMemcpyAsync(static_cast<char *>(array) + g_first * nb, d_slice, first_len * nb, cudaMemcpyDeviceToHost, c->stream));
    if (first_len < n) CU_TRY(cudaMemcpyAsync(array, static_cast<char *>(d_slice) + first_len * nb, (n - first_len) * nb, cudaMemcpyDeviceToHost, c->stream));
    if (nul_flag) {
        // run 1 covers global slots [g_first, g_first + first_len), its words start at word g_first / 32
        int k = 0;
        auto copy_run = [&](uint64_t s0, uint64_t len, const u32 *d_words, uint64_t w0) -> int {
            if (len == 0) return DBG_OK;
            const uint64_t s1 = s0 + len;                               // exclusive
            uint64_t b0 = (s0 + 7) / 8, b1 = s1 / 8;                     // bytes covered completely
            if (s1 == c->P) b1 = c->P / 8 + 1;                           // the table's last byte belongs to the slice that ends it
            if (s0 % 8 && edge_slots && k < 4) edge_slots[k++] = s0;
            if (s1 % 8 && s1 != c->P && edge_slots && k < 4) edge_slots[k++] = s1 - 1;
            if (b1 > b0) CU_TRY(cudaMemcpyAsync(nul_flag + b0, reinterpret_cast<const uint8_t *>(d_words) + (b0 - w0 * 4), b1 - b0, cudaMemcpyDeviceToHost, c->stream));
            return DBG_OK;
        };
        rc = copy_run(g_first, first_len, c->d_nul_slice, g_first / 32);
        if (rc) return rc;
        if (first_len < n) { rc = copy_run(0, n - first_len, c->d_nul_slice + c->nul_slice_words + 2, 0); if (rc) return rc; }
    }
    CU_TRY(cudaEventRecord(e.b, c->stream));
    CU_TRY(cudaEventSynchronize(e.b));
    CU_TRY(cudaEventElapsedTime(&c->ms[5], e.a, e.b));
    ev_put(c, e);
    return DBG_OK;
}

// host helpers for whoever merges the slices (no device work): recompute the nul_flag bytes that contain the given
// slots from the table image, and append the k-mer-0 node last (add_node_to_kmerset, kmerSet.cpp:253-273,
// DBGgraph.cpp:418).  Fix the edge bytes BEFORE the k-mer-0 node goes in (its key is 0, like an empty slot).
extern "C" int dbg_host_fix_nul_bytes(const void *array, uint8_t *nul_flag, uint64_t P, int32_t wide, const uint64_t *slots, uint64_t n)
{
    if (!array || !nul_flag || (!slots && n)) return DBG_ERR_INVALID;
    for (uint64_t i = 0; i < n; i++) {
        if (slots[i] == UINT64_MAX || slots[i] >= P) continue;
        const uint64_t byte = slots[i] / 8;
        uint8_t v = 0;
        for (uint64_t s = byte * 8; s < byte * 8 + 8 && s < P; s++) {
            bool occ;
            if (wide) { const dbg_node32 *nd = static_cast<const dbg_node32 *>(array) + s; occ = (nd->kmer_lo | nd->kmer_hi) != 0; }
            else occ = static_cast<const dbg_node16 *>(array)[s].kmer != 0;
            if (occ) v |= (uint8_t)(0x80u >> (s & 7));
        }
        nul_flag[byte] = v;
    }
    return DBG_OK;
}

extern "C" int dbg_host_polyA_insert(void *array, uint8_t *nul_flag, uint64_t P, int32_t wide, uint32_t l_link, uint32_t r_link, uint64_t *slot_out)
{
    if (!array || !nul_flag || P == 0) return DBG_ERR_INVALID;
    uint64_t s = (wide ? hash_code_wide(0, 0) : hash_code(0)) % P;
    uint64_t steps = 0;
    while (nul_flag[s >> 3] & (0x80u >> (s & 7))) { s = (s + 1 == P) ? 0 : s + 1; if (++steps > P) return DBG_ERR_TABLE_FULL; }
    if (wide) { dbg_node32 *nd = static_cast<dbg_node32 *>(array) + s; nd->kmer_lo = 0; nd->kmer_hi = 0; nd->l_link = l_link; nd->r_link = r_link; nd->pad = 0; }
    else { dbg_node16 *nd = static_cast<dbg_node16 *>(array) + s; nd->kmer = 0; nd->l_link = l_link; nd->r_link = r_link; }
    nul_flag[s >> 3] |= (uint8_t)(0x80u >> (s & 7));
    if (slot_out) *slot_out = s;
    return DBG_OK;
}

extern "C" int dbg_get_stats(dbg_ctx *c, dbg_stats *stats)
{
    if (!c || !stats) return set_err(DBG_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    u64 cnt[CNT_N];
    int rc = read_counters(c, cnt);
    if (rc) return rc;
    fill_stats(c, cnt);
    *stats = c->st;
    return DBG_OK;
}

extern "C" int dbg_device_image(dbg_ctx *c, void **d_array, void **d_nul_flag)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    if (!c->finalized || !c->d_out || c->n_shards > 1) return set_err(DBG_ERR_STATE, "no finalized image (unsharded contexts only; sharded: dbg_shard_slice_info)");
    if (d_array) *d_array = c->d_out;
    if (d_nul_flag) *d_nul_flag = c->d_nul32;
    return DBG_OK;
}

extern "C" int dbg_device_build_table(dbg_ctx *c, void **d_nodes, uint64_t *n_local, uint64_t **d_polyA)
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    { cudaSetDevice(c->device); int trc = table_ready(c, c->stream); if (trc) return trc; }
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (d_nodes) *d_nodes = c->d_nodes;
    if (n_local) *n_local = c->n_local;
    if (d_polyA) *d_polyA = (uint64_t *)c->d_polyA;
    return DBG_OK;
}

extern "C" int dbg_export_kmerset(dbg_ctx *c, void *array, uint8_t *nul_flag)
{
    if (!c || !array || !nul_flag) return set_err(DBG_ERR_INVALID, "NULL argument");
    if (!c->finalized || !c->d_out || c->n_shards > 1) return set_err(DBG_ERR_STATE, "dbg_export_kmerset needs dbg_finalize on an unsharded context (sharded: dbg_export_shard_slice)");
    CU_TRY(cudaSetDevice(c->device));
    const uint64_t image_bytes = c->P * (uint64_t)node_bytes(c), nul_bytes = c->P / 8 + 1;
    // DBG_B200_EXPORT=pipe: pipelined hand-over (export_pipe.cu) -- occupied nodes only over the link, host threads expand
    // them into `array`; it steps aside (returns 1) for small tables, without host threads, or when its buffers cannot be
    // allocated.  Opt-in: it halves the bytes on the link but triples the host-memory traffic, and on the measured hosts
    // (16-vCPU VMs, ~80 GB/s of host memory bandwidth) the plain DMA copy is faster (profiles/README.md); same bytes either way.
    const char *mode = getenv("DBG_B200_EXPORT");
    if (mode && strcmp(mode, "pipe") == 0) {
        char msg[256] = "";
        float ms = 0;
        int prc = export_pipe_run(&c->pipe, c->device, c->stream, c->d_out, c->d_nul32, nul_words(c->P), c->P, node_bytes(c), array, nul_flag,
                                  &ms, c->export_info, msg, (int)sizeof(msg));
        if (prc < 0) return set_err(DBG_ERR_CUDA, "%s", msg);
        if (prc == 0) { c->ms[5] = ms; c->launches += 2 + c->export_info[0] + c->export_info[1]; return DBG_OK; }
    }
    EvPair e;
    int rc = ev_begin(c, c->stream, &e);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(array, c->d_out, image_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(nul_flag, c->d_nul32, nul_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaEventRecord(e.b, c->stream));
    CU_TRY(cudaEventSynchronize(e.b));
    CU_TRY(cudaEventElapsedTime(&c->ms[5], e.a, e.b));
    ev_put(c, e);
    c->export_info[0] = 0; c->export_info[1] = 1; c->export_info[2] = image_bytes + nul_bytes; c->export_info[3] = c->st.count;
    return DBG_OK;
}

extern "C" int dbg_finish_export(dbg_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads, dbg_stats *stats, void *array, uint8_t *nul_flag)
{
    if (!c || !array || !nul_flag) return set_err(DBG_ERR_INVALID, "NULL argument");
    if (c->n_shards > 1) return set_err(DBG_ERR_STATE, "dbg_finish_export: unsharded contexts only");
    FinishExport fx{array, nul_flag, false};
    if (n_reads) {
        int rc = submit_impl(c, bases, offs, n_reads, &fx);
        if (rc) return rc;
    }
    int rc = dbg_finalize(c, stats);
    if (rc) return rc;
    return fx.done ? DBG_OK : dbg_export_kmerset(c, array, nul_flag);
}

extern "C" int dbg_export_info(const dbg_ctx *c, uint64_t info[4])
{
    if (!c || !info) return set_err(DBG_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 4; i++) info[i] = c->export_info[i];
    return DBG_OK;
}

extern "C" uint64_t dbg_host_expand_nodes(const uint8_t *bits, uint64_t n_slots, const void *nodes, void *array, int32_t wide)
{
    if (!bits || !nodes || !array) return 0;
    return expand_nodes(bits, n_slots, nodes, array, wide ? 32 : 16);
}

// ---------------------------------------------------------------------------------------------------
// calculate_kmer_links + compaction
// ---------------------------------------------------------------------------------------------------
static int run_links(dbg_ctx *c, int cutoff)
{
    if (!c->finalized || !c->d_out || c->n_shards > 1) return set_err(DBG_ERR_STATE, "needs dbg_finalize on an unsharded context");
    if (c->links_cutoff == cutoff) return DBG_OK;
    uint64_t n_tiles = (c->P + TILE - 1) / TILE;
    if (!c->d_klink) {
        CU_TRY(cudaMalloc(&c->d_klink, c->P * sizeof(unsigned short)));
        CU_TRY(cudaMalloc(&c->d_del32, nul_words(c->P) * sizeof(u32)));
        CU_TRY(cudaMalloc(&c->d_tile_counts, n_tiles * 4 * sizeof(u32)));
        CU_TRY(cudaMalloc(&c->d_tile_offs, n_tiles * 4 * sizeof(u64)));
        CU_TRY(cudaMalloc(&c->d_small, (256 + 3 + 4 + 1) * sizeof(u64)));
    }
    EvPair e;
    int rc = ev_begin(c, c->stream, &e);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(c->d_small, 0, (256 + 3 + 4 + 1) * sizeof(u64), c->stream));
    CU_TRY(cudaMemsetAsync(c->d_del32, 0, nul_words(c->P) * sizeof(u32), c->stream));
    unsigned grid = (unsigned)(n_tiles < 148ull * 8 ? n_tiles : 148ull * 8);
    k_links_classify<<<grid, 256, 0, c->stream>>>(c->d_out, c->d_nul32, c->P, c->wide ? 1 : 0, cutoff, c->d_klink, c->d_del32,
                                                  c->d_small, c->d_small + 256, c->d_tile_counts);
    CU_TRY(cudaGetLastError());
    k_scan_tiles<<<1, 1024, 0, c->stream>>>(c->d_tile_counts, n_tiles, c->d_tile_offs, c->d_small + 259);
    CU_TRY(cudaGetLastError());
    c->launches += 2;
    CU_TRY(cudaEventRecord(e.b, c->stream));
    CU_TRY(cudaEventSynchronize(e.b));
    CU_TRY(cudaEventElapsedTime(&c->ms[3], e.a, e.b));
    ev_put(c, e);
    c->links_cutoff = cutoff;
    return DBG_OK;
}

extern "C" int dbg_export_links(dbg_ctx *c, int32_t freq_cutoff, uint8_t *klink, uint8_t *del_flag, int64_t depth_hist[256],
                                uint64_t *tips, uint64_t *n_tips, uint64_t *branches, uint64_t *n_branches, int64_t stats3[3])
{
    if (!c) return set_err(DBG_ERR_INVALID, "NULL ctx");
    CU_TRY(cudaSetDevice(c->device));
    int rc = run_links(c, freq_cutoff);
    if (rc) return rc;
    u64 small[264];
    CU_TRY(cudaMemcpy(small, c->d_small, sizeof(small), cudaMemcpyDeviceToHost));
    if (depth_hist) for (int i = 0; i < 256; i++) depth_hist[i] = (int64_t)small[i];
    if (stats3) for (int i = 0; i < 3; i++) stats3[i] = (int64_t)small[256 + i];
    uint64_t nt = small[259], nb = small[260];
    if (klink) CU_TRY(cudaMemcpy(klink, c->d_klink, c->P * 2, cudaMemcpyDeviceToHost));
    if (del_flag) CU_TRY(cudaMemcpy(del_flag, c->d_del32, c->P / 8 + 1, cudaMemcpyDeviceToHost));
    if ((tips && n_tips) || (branches && n_branches)) {
        uint64_t cap_t = (tips && n_tips) ? *n_tips : 0, cap_b = (branches && n_branches) ? *n_branches : 0;
        if ((tips && cap_t < nt) || (branches && cap_b < nb)) {
            if (n_tips) *n_tips = nt;
            if (n_branches) *n_branches = nb;
            return set_err(DBG_ERR_BUFFER, "tip/branch list capacity too small (%llu/%llu needed)", (unsigned long long)nt, (unsigned long long)nb);
        }
        u64 *d_t = nullptr, *d_b = nullptr;
        CU_TRY(cudaMalloc(&d_t, (nt + 1) * sizeof(u64)));
        CU_TRY(cudaMalloc(&d_b, (nb + 1) * sizeof(u64)));
        uint64_t n_tiles = (c->P + TILE - 1) / TILE;
        unsigned grid = (unsigned)(n_tiles < 148ull * 8 ? n_tiles : 148ull * 8);
        k_links_lists<<<grid, 256, 0, c->stream>>>(c->d_klink, c->d_nul32, c->P, c->d_tile_offs, d_t, nt, d_b, nb);
        c->launches++;
        cudaError_t e1 = cudaGetLastError();
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->stream);
        if (e1 == cudaSuccess && tips && nt) e1 = cudaMemcpy(tips, d_t, nt * sizeof(u64), cudaMemcpyDeviceToHost);
        if (e1 == cudaSuccess && branches && nb) e1 = cudaMemcpy(branches, d_b, nb * sizeof(u64), cudaMemcpyDeviceToHost);
        cudaFree(d_t); cudaFree(d_b);
        CU_TRY(e1);
    }
    if (n_tips) *n_tips = nt;
    if (n_branches) *n_branches = nb;
    return DBG_OK;
}

extern "C" int dbg_dump_compact(dbg_ctx *c, int32_t freq_cutoff, uint64_t *slots, uint64_t *kmers_lo, uint64_t *kmers_hi,
                                uint32_t *l_link, uint32_t *r_link, uint64_t *n)
{
    if (!c || !n) return set_err(DBG_ERR_INVALID, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    int which = freq_cutoff < 0 ? 3 : 2;
    int rc = run_links(c, freq_cutoff < 0 ? (c->links_cutoff != INT32_MIN ? c->links_cutoff : 0) : freq_cutoff);
    if (rc) return rc;
    u64 totals[4];
    CU_TRY(cudaMemcpy(totals, c->d_small + 259, sizeof(totals), cudaMemcpyDeviceToHost));
    uint64_t m = totals[which], cap = *n;
    *n = m;
    if (!slots && !kmers_lo && !kmers_hi && !l_link && !r_link) return DBG_OK;   // size query
    if (cap < m) return set_err(DBG_ERR_BUFFER, "dump capacity %llu < %llu", (unsigned long long)cap, (unsigned long long)m);
    u64 *d_s = nullptr, *d_lo = nullptr, *d_hi = nullptr; u32 *d_l = nullptr, *d_r = nullptr;
    cudaError_t e1 = cudaSuccess;
    if (slots) e1 = cudaMalloc(&d_s, (m + 1) * 8);
    if (e1 == cudaSuccess && kmers_lo) e1 = cudaMalloc(&d_lo, (m + 1) * 8);
    if (e1 == cudaSuccess && kmers_hi) e1 = cudaMalloc(&d_hi, (m + 1) * 8);
    if (e1 == cudaSuccess && l_link) e1 = cudaMalloc(&d_l, (m + 1) * 4);
    if (e1 == cudaSuccess && r_link) e1 = cudaMalloc(&d_r, (m + 1) * 4);
    if (e1 == cudaSuccess) {
        uint64_t n_tiles = (c->P + TILE - 1) / TILE;
        unsigned grid = (unsigned)(n_tiles < 148ull * 8 ? n_tiles : 148ull * 8);
        k_compact_nodes<<<grid, 256, 0, c->stream>>>(c->d_out, c->d_nul32, c->d_del32, c->P, c->wide ? 1 : 0, which, c->d_tile_offs, m,
                                                     d_s, d_lo, d_hi, d_l, d_r);
        c->launches++;
        e1 = cudaGetLastError();
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->stream);
    }
    if (e1 == cudaSuccess && slots && m) e1 = cudaMemcpy(slots, d_s, m * 8, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && kmers_lo && m) e1 = cudaMemcpy(kmers_lo, d_lo, m * 8, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && kmers_hi && m) e1 = cudaMemcpy(kmers_hi, d_hi, m * 8, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && l_link && m) e1 = cudaMemcpy(l_link, d_l, m * 4, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && r_link && m) e1 = cudaMemcpy(r_link, d_r, m * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_s); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_l); cudaFree(d_r);
    CU_TRY(e1);
    return DBG_OK;
}

extern "C" int dbg_dump_shard(dbg_ctx *c, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *l_link, uint32_t *r_link,
                              uint64_t *first_ordinal, uint64_t *n)
{
    if (!c || !n) return set_err(DBG_ERR_INVALID, "NULL argument");
    { cudaSetDevice(c->device); int trc = table_ready(c, c->stream); if (trc) return trc; }
    CU_TRY(cudaSetDevice(c->device));
    u64 cnt[CNT_N];
    int rc = read_counters(c, cnt);
    if (rc) return rc;
    uint64_t m = (uint64_t)((int64_t)cnt[CNT_NEW] + c->nodes_delta), cap = *n;      // what is physically in this shard's table
    *n = m;
    if (!kmers_lo && !kmers_hi && !l_link && !r_link && !first_ordinal) return DBG_OK;   // size query
    if (cap < m) return set_err(DBG_ERR_BUFFER, "dump capacity %llu < %llu", (unsigned long long)cap, (unsigned long long)m);
    u64 *d_lo = nullptr, *d_hi = nullptr, *d_ord = nullptr, *d_cur = nullptr; u32 *d_l = nullptr, *d_r = nullptr;
    cudaError_t e1 = cudaMalloc(&d_cur, 8);
    if (e1 == cudaSuccess) e1 = cudaMemset(d_cur, 0, 8);
    if (e1 == cudaSuccess && kmers_lo) e1 = cudaMalloc(&d_lo, (m + 1) * 8);
    if (e1 == cudaSuccess && kmers_hi) e1 = cudaMalloc(&d_hi, (m + 1) * 8);
    if (e1 == cudaSuccess && first_ordinal) e1 = cudaMalloc(&d_ord, (m + 1) * 8);
    if (e1 == cudaSuccess && l_link) e1 = cudaMalloc(&d_l, (m + 1) * 4);
    if (e1 == cudaSuccess && r_link) e1 = cudaMalloc(&d_r, (m + 1) * 4);
    if (e1 == cudaSuccess) {
        if (c->wide) k_dump_shard<true><<<148 * 8, 256, 0, c->stream>>>((const NodeT<true> *)c->d_nodes, c->n_local, m, d_cur, d_lo, d_hi, d_l, d_r, d_ord);
        else k_dump_shard<false><<<148 * 8, 256, 0, c->stream>>>((const NodeT<false> *)c->d_nodes, c->n_local, m, d_cur, d_lo, d_hi, d_l, d_r, d_ord);
        c->launches++;
        e1 = cudaGetLastError();
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(c->stream);
    }
    if (e1 == cudaSuccess && kmers_lo && m) e1 = cudaMemcpy(kmers_lo, d_lo, m * 8, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && kmers_hi && m) e1 = cudaMemcpy(kmers_hi, d_hi, m * 8, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && first_ordinal && m) e1 = cudaMemcpy(first_ordinal, d_ord, m * 8, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && l_link && m) e1 = cudaMemcpy(l_link, d_l, m * 4, cudaMemcpyDeviceToHost);
    if (e1 == cudaSuccess && r_link && m) e1 = cudaMemcpy(r_link, d_r, m * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_cur); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_ord); cudaFree(d_l); cudaFree(d_r);
    CU_TRY(e1);
    return DBG_OK;
}

extern "C" int dbg_get_timings(dbg_ctx *c, float ms[8])
{
    if (!c || !ms) return set_err(DBG_ERR_INVALID, "NULL argument");
    collect_clear_time(c);
    for (int i = 0; i < 8; i++) ms[i] = c->ms[i];
    return DBG_OK;
}

extern "C" uint64_t dbg_launch_count(const dbg_ctx *c) { return c ? c->launches : 0; }

extern "C" int dbg_path_counts(const dbg_ctx *c, uint64_t counts[4])
{
    if (!c || !counts) return set_err(DBG_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 4; i++) counts[i] = c->path_counts[i];
    return DBG_OK;
}

// ---------------------------------------------------------------------------------------------------
// roofline denominator
// ---------------------------------------------------------------------------------------------------
extern "C" int dbg_measure_random_rmw(int32_t device, uint64_t bytes, uint64_t n_ops, int32_t mode, float *ms)
{
    if (!ms || bytes < sizeof(Rec32) || mode < 0 || mode > 5) return set_err(DBG_ERR_INVALID, "bad argument");
    if (dbg_device_count() == 0) return set_err(DBG_ERR_CUDA, "no CUDA device visible");
    CU_TRY(cudaSetDevice(device));
    Rec32 *tab = nullptr;
    CU_TRY(cudaMalloc(&tab, bytes + 8));
    cudaError_t e1 = cudaMemset(tab, 0, bytes + 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    uint64_t n_nodes = bytes / sizeof(Rec32);
    u64 *sink = reinterpret_cast<u64 *>(reinterpret_cast<char *>(tab) + (bytes & ~7ull));
    unsigned grid = 148 * 32;
    float best = 1e30f;
    for (int it = 0; it < 4 && e1 == cudaSuccess; it++) {   // first iteration is warm-up
        cudaEventRecord(a);
        u64 seed = 0x1234567ull * (it + 1);
        switch (mode) {
        case 0: k_random_rmw<0><<<grid, 256>>>(tab, n_nodes, n_ops, seed, sink); break;
        case 1: k_random_rmw<1><<<grid, 256>>>(tab, n_nodes, n_ops, seed, sink); break;
        case 2: k_random_rmw<2><<<grid, 256>>>(tab, n_nodes, n_ops, seed, sink); break;
        case 3: k_random_rmw<3><<<grid, 256>>>(tab, n_nodes, n_ops, seed, sink); break;
        case 4: k_random_rmw<4><<<grid, 256>>>(tab, n_nodes, n_ops, seed, sink); break;
        default: k_random_rmw<5><<<grid, 256>>>(tab, n_nodes, n_ops, seed, sink); break;
        }
        cudaEventRecord(b);
        e1 = cudaEventSynchronize(b);
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        if (it > 0 && t < best) best = t;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(tab);
    CU_TRY(e1);
    *ms = best;
    return DBG_OK;
}
