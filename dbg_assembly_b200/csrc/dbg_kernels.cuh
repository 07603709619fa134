// dbg_kernels.cuh -- the sm_100a kernels of the De Bruijn graph build.
//
//   k_chunk_first       : reads -> fixed-size base chunks (one CTA per chunk later)
//   k_build<..>         : FUSED  stage ASCII -> 2-bit pack in shared memory -> canonical k-mer + neighbour
//                         bases per occurrence -> sink.  Sinks: InsertSink (direct hash insert),
//                         StagedScatterSink<OPT> (radix partition by table slice through shared-memory batches
//                         copied out in bucket order; OPT = single pass into fixed bucket regions),
//                         PartitionSink<0/1> (exact two-pass partition: counts, then offsets from the scan),
//                         PeerStagedSink (owner-sorted runs into the peers' buffers over NVLink),
//                         FreqSink (kfreq.cu: direct-index k-mer counts).
//   k_part_scan1/2/3    : column scan of the per-chunk bucket counts -> exact write offsets
//   k_opt_finish        : bucket offsets + tail zeroing after an optimistic partition
//   k_tuple_partition, k_tuple_scatter_staged : the same partition over received tuples (owner side of the exchange)
//   k_insert_tuples     : bucket-ordered hash insert through L2-resident table slices
//   k_layout_*          : rebuild the reference's slot layout (first-occurrence priority linear probing)
//   k_links_* / k_compact_* / k_dump_shard : calculate_kmer_links + ordered stream compaction, shard dump
//
// Reference semantics: DBGgraph.cpp:38-213 (parse + update), kmerSet.cpp:253-273 (poly-A node),
// contig.cpp:107-205 (link pass).  Nothing here is a translation of the pthread code: the reference
// materialises (kmer,left,right) per block and lets T threads rescan it; here one CTA owns a chunk of
// bases, keeps it packed in shared memory and pushes occurrences straight into the HBM table or into
// slice-ordered tuples.
#pragma once
#include "dbg_core.cuh"

namespace dbg {

constexpr int CB = 16384;      // bases per chunk (one CTA)
constexpr int BLOCK = 256;     // threads per CTA
constexpr int MAXR = 1024;     // reads per shared-memory table pass
constexpr int G = 4;           // consecutive occurrences per thread run (rolling k-mer + 4 probes in flight)
#ifndef DBG_MIN_CTAS
#define DBG_MIN_CTAS 2
#endif
constexpr int MIN_CTAS = DBG_MIN_CTAS;    // register budget of the insert kernels: CTAs (x8 warps) per SM
constexpr u64 EMPTY_PRI = ~0ULL;
constexpr u64 POLYA_PRI = ~0ULL - 1;

struct Occ {
    u64 klo, khi;
    u32 lb, rb;     // 0..3 or 4 (= no neighbour: read end)
    u64 ord;        // read_index << 16 | j
};

struct BuildArgs {
    const char *bases;        // device pointer; global offsets index into it
    const u64 *offs;          // n_reads + 1
    u64 n_reads;
    u64 abase;                // 16-B aligned byte where chunk 0 starts
    u64 end_base;             // one past the last valid base byte
    const u64 *chunk_first;   // n_chunks + 1: first read starting in each chunk
    u64 read_index0;          // global index of read 0 (ordinals)
    int K, R;
    u32 stage_words;          // packed words staged per chunk (incl. slack)
    int count_stats;          // add this launch's reads/occurrences to the context counters
    int seed;                 // 1: contig seed index (link_scaffold/map_func.cpp:119-172): occurrences carry no neighbour bases but
                              // the strand of the k-mer (ordinal << 1 | direct) and count into lane 0; tie -> direct 0
};

// seed-index occurrences (BuildArgs::seed): lb = 0 (the occurrence count lives in l lane A), rb = RB_SEED: "no right
// neighbour" for every insert path (rb >= 4), and the mark by which the k-mer-0 side node also keeps its first ordinal
constexpr u32 RB_SEED = 5;

// the k-mer-0 side node (DBGgraph.cpp:153-164): counts above 255 change nothing after clamping, so stop adding once a
// lane is saturated (bounds contention).  Seed index: polyA[0] counts the occurrences, polyA[7] keeps ~(first ordinal).
__device__ __forceinline__ void polyA_bump(u64 *polyA, u32 lb, u32 rb, u64 ord)
{
    if (lb < 4 && __ldcg(polyA + lb) < 255) atomicAdd(polyA + lb, 1ULL);
    if (rb < 4 && __ldcg(polyA + 4 + rb) < 255) atomicAdd(polyA + 4 + rb, 1ULL);
    if (rb == RB_SEED && ~ord > __ldcg(polyA + 7)) atomicMax(polyA + 7, ~ord);
}

// ---------------------------------------------------------------------------------------------------
// chunk index: chunk_first[c] = first read whose start offset lies in chunk >= c
// ---------------------------------------------------------------------------------------------------
static __global__ void k_chunk_first(const u64 *__restrict__ offs, u64 n_reads, u64 abase, u64 n_chunks,
                              u64 *__restrict__ chunk_first)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    u64 c = n_chunks;
    if (r < n_reads) { c = (offs[r] - abase) / CB; if (c > n_chunks) c = n_chunks; }
    u64 cprev_plus1 = 0;
    if (r > 0) { u64 cp = (offs[r - 1] - abase) / CB; if (cp > n_chunks) cp = n_chunks; cprev_plus1 = cp + 1; }
    for (u64 cc = cprev_plus1; cc <= c; cc++) chunk_first[cc] = r;
}

// ---------------------------------------------------------------------------------------------------
// sinks
// ---------------------------------------------------------------------------------------------------
// thread_updatekmers for ONE occurrence (DBGgraph.cpp:167-205), lock-free: load the whole node with one
// 256-bit request, claim an empty slot with a CAS (the only atomic with a return value, and only for NEW
// keys), count with one fire-and-forget vector RED, keep the earliest ordinal with a RED.max.
template <bool WIDE, bool TRACK>
__device__ __forceinline__ void insert_probe(const TableView &t, NodeT<WIDE> *p, NodeRegs n, u64 klo, u64 khi, u32 lb, u32 rb, u64 ord,
                                             u32 &n_new, u32 &n_conf);

template <bool WIDE, bool TRACK>
__device__ __forceinline__ void insert_one(const TableView &t, u64 klo, u64 khi, u32 lb, u32 rb, u64 ord, u32 &n_new, u32 &n_conf)
{
    typedef NodeT<WIDE> Nd;
    u64 h = WIDE ? hash_code_wide(klo, khi) : hash_code(klo);
    Nd *p = static_cast<Nd *>(t.nodes) + (mod_P(h, t.P, t.M) - t.lo);
    NodeRegs n;
    load_node(p, n);
    insert_probe<WIDE, TRACK>(t, p, n, klo, khi, lb, rb, ord, n_new, n_conf);
}

// A (nearly) full table must be REPORTED, not scanned slot by slot to its end (the front end then rebuilds with a larger
// device table).  DBG_GUARD picks how (measured on C2, profiles/README.md): 1-3 = a budget of long probe steps checked inside
// the probe loop -- correct, but the extra code spills at the insert kernel's 32-register budget and costs 9 % of it;
// 4 (default) = the probe length is capped through the loop's own end pointer, which costs nothing.
#ifndef DBG_GUARD
#define DBG_GUARD 4
#endif
constexpr long long PROBE_CAP = 1ll << 22;      // slots a probe may walk from its home slot (longest legitimate clusters: ~1e5)
#if DBG_GUARD == 2
static __device__ __noinline__ bool probe_guard(u64 *counters, u64 budget)
{
    if (atomicAdd(counters + CNT_GUARD, 1ULL) >= budget || __ldcg(counters + CNT_ERROR)) {
        atomicExch(counters + CNT_ERROR, 1ULL); atomicExch(counters + 7, 1ULL << 40);
        return true;
    }
    return false;
}
#endif

// the probe loop, entered with the home slot's node already loaded (callers may have several loads in flight)
template <bool WIDE, bool TRACK>
__device__ __forceinline__ void insert_probe(const TableView &t, NodeT<WIDE> *p, NodeRegs n, u64 klo, u64 khi, u32 lb, u32 rb, u64 ord,
                                             u32 &n_new, u32 &n_conf)
{
    typedef NodeT<WIDE> Nd;
#if DBG_GUARD == 4
    // a probe ends at the end of the shard (+ margin) or PROBE_CAP slots from home, whichever comes first: the check is
    // the one the loop has anyway, so a (nearly) full table is REPORTED -- the first probe that runs out raises the error
    // and poisons the tile counter of k_insert_tuples (counters[7]), the probes in flight end within PROBE_CAP slots --
    // instead of being scanned slot by slot to its end
    Nd *p_end = static_cast<Nd *>(t.nodes) + t.n_local;
    if (p_end - p > (long long)PROBE_CAP) p_end = p + PROBE_CAP;
#else
    Nd *const p_end = static_cast<Nd *>(t.nodes) + t.n_local;
#endif
    for (;;) {
        if ((n.klo | n.khi) == 0) {
            if (WIDE) {
                u64 olo, ohi;
                if (cas128(p, klo, khi, olo, ohi)) { n_new++; n.klo = klo; n.khi = khi; }
                else { n.klo = olo; n.khi = ohi; }
            } else {
                u64 old = atomicCAS(&p->klo, 0ULL, klo);
                if (old == 0) { n_new++; n.klo = klo; }
                else n.klo = old;                       // somebody else got the slot
            }
        }
        if (n.klo == klo && (!WIDE || n.khi == khi)) {
            // a stale (too small) loaded count only costs a redundant add; export clamps to 255
            u32 l = lb < 4 ? ((u32)(n.c0 >> (16 * lb)) & 0xFFFFu) : 0xFFFFu;
            u32 r = rb < 4 ? ((u32)(n.c1 >> (16 * rb)) & 0xFFFFu) : 0xFFFFu;
            bump_counts<WIDE>(p, l | (r << 16), lb, rb);
            if (TRACK) {
                u64 mn = ~ord;
                if (mn > n.nord) atomicMax(&p->nord, mn);
            }
            return;
        }
        n_conf++;                                       // occupied by another key: next slot (DBGgraph.cpp:201-204)
        p++;
        if (p >= p_end) { atomicExch(t.counters + CNT_ERROR, 1ULL); atomicExch(t.counters + 7, 1ULL << 40); return; }
#if DBG_GUARD == 3
        // every 16384 probe steps of this thread (its running conflict counter is live anyway: no extra register): long
        // probes draw on a budget of the whole build (see probe_guard)
        if ((n_conf & 0x3FFFu) == 0) {
            if (atomicAdd(t.counters + CNT_GUARD, 1ULL) >= t.guard_budget || __ldcg(t.counters + CNT_ERROR)) {
                atomicExch(t.counters + CNT_ERROR, 1ULL); atomicExch(t.counters + 7, 1ULL << 40); return;
            }
        }
#elif DBG_GUARD == 2
        if ((reinterpret_cast<unsigned long long>(p) & (1024 * sizeof(Nd) - 1)) == 0 && probe_guard(t.counters, t.guard_budget)) return;
#elif DBG_GUARD == 1
        if ((reinterpret_cast<unsigned long long>(p) & (1024 * sizeof(Nd) - 1)) == 0) {
            if (atomicAdd(t.counters + CNT_GUARD, 1ULL) >= t.guard_budget || __ldcg(t.counters + CNT_ERROR)) {
                atomicExch(t.counters + CNT_ERROR, 1ULL); atomicExch(t.counters + 7, 1ULL << 40); return;
            }
        }
#endif
        load_node(p, n);
    }
}

template <bool WIDE, bool TRACK>
struct InsertSink {
    static constexpr int RUN = G;
    static constexpr int MIN_BLOCKS = MIN_CTAS;
    TableView t;
    u32 n_new, n_conf;     // per-thread, reduced at kernel end
    bool dead;             // the table was already found full: skip the inserts (the host reports DBG_ERR_TABLE_FULL)

    __device__ __forceinline__ void init(u32 *) { n_new = 0; n_conf = 0; dead = __ldcg(t.counters + CNT_ERROR) != 0; }

    __device__ __forceinline__ void polyA(const Occ &o)
    {
        // k-mer 0 never enters the table during the build (DBGgraph.cpp:153-164)
        polyA_bump(t.polyA, o.lb, o.rb, o.ord);
    }

    __device__ __forceinline__ void consume(const Occ (&o)[G], int nv)
    {
        if (dead) return;
        for (int g = 0; g < nv; g++) {
            if ((o[g].klo | o[g].khi) == 0) polyA(o[g]);
            else insert_one<WIDE, TRACK>(t, o[g].klo, o[g].khi, o[g].lb, o[g].rb, o[g].ord, n_new, n_conf);
        }
    }

    __device__ __forceinline__ void finish()
    {
        u32 a = n_new, b = n_conf;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, s); b += __shfl_xor_sync(0xffffffffu, b, s); }
        if ((threadIdx.x & 31) == 0) {
            if (a) atomicAdd(t.counters + CNT_NEW, (u64)a);
            if (b) atomicAdd(t.counters + CNT_CONFLICT, (u64)b);
        }
    }
};

// tuples (partitioned build and multi-GPU exchange): 16 B (narrow) {kmer, ord<<8 | rb<<4 | lb}, 32 B (wide) {lo, hi, meta, 0}
// Radix partition by home-slot range (bucket = home >> shift), exact and atomic-free in global memory:
//   pass 1 (MODE 0)  every chunk CTA histograms its occurrences per bucket in shared memory and stores the row
//                    M[chunk][bucket];
//   k_part_scan*     column-wise exclusive scan turns M into the exact write offset of every (chunk, bucket);
//   pass 2 (MODE 1)  the chunk CTA regenerates its occurrences and writes each tuple to
//                    M[chunk][bucket] + (shared-memory rank): no global atomics, no barriers per round.
// The insert pass then walks the tuples bucket by bucket, so every bucket's slice of the table (16 MB)
// is pulled into the 126 MB L2 once, updated there and written back once, instead of one random DRAM
// read-modify-write per occurrence.
template <bool WIDE, int MODE>
struct PartitionSink {
    static constexpr int RUN = G;
    static constexpr int MIN_BLOCKS = MIN_CTAS;
    TableView t;
    int shift;         // bucket = (home - lo) >> shift          (table slices), or, when div != 0,
    u64 div, div_M;    // bucket = home / div                    (owner ranks of the multi-GPU exchange), or, when
    u32 nb_local;      // nb_local != 0: owner * nb_local + ((home - owner*div) >> shift)   (owner AND its table slice)
    u32 n_buckets;
    u32 *matrix;       // [n_chunks][n_buckets]: counts after pass 1, write offsets after the scan
    u64 *tuples;
    // fused exchange (multi-GPU): bucket q's tuples are stored straight into rank q's receive buffer over
    // NVLink peer mappings, at dst_base[q] + (this rank's running offset for q); NULL = local packed output
    u64 *const *dst_ptrs;
    const u64 *dst_base;   // first tuple index reserved for this rank in each peer's buffer
    const u64 *roffs;      // this rank's packed start of each bucket (subtracted from the matrix offsets)
    u32 *hist;         // shared: n_buckets running counts of this chunk
    u32 *base;         // shared: n_buckets write offsets of this chunk (MODE 1)

    __device__ __forceinline__ void init(u32 *extra)
    {
        hist = extra;
        base = extra + n_buckets;
        const u32 *row = matrix + (size_t)blockIdx.x * n_buckets;
        for (u32 b = threadIdx.x; b < n_buckets; b += BLOCK) {
            hist[b] = 0;
            if (MODE == 1) base[b] = dst_ptrs ? (u32)((u64)row[b] - roffs[b] + dst_base[b]) : row[b];
        }
        __syncthreads();
    }

    __device__ __forceinline__ void consume(const Occ (&o)[RUN], int nv)
    {
#pragma unroll
        for (int g = 0; g < RUN; g++) {
            if (g < nv) {
                if ((o[g].klo | o[g].khi) == 0) {
                    if (MODE == 1) {   // the k-mer-0 side node is accumulated once, in the scatter pass
                        polyA_bump(t.polyA, o[g].lb, o[g].rb, o[g].ord);
                    }
                } else {
                    u64 hh = WIDE ? hash_code_wide(o[g].klo, o[g].khi) : hash_code(o[g].klo);
                    u64 home = mod_P(hh, t.P, t.M);
                    u32 bkt;
                    if (div) {
                        bkt = (u32)__umul64hi(home, div_M); if ((u64)(bkt + 1) * div <= home) bkt++;
                        if (nb_local) bkt = bkt * nb_local + (u32)((home - (u64)bkt * div) >> shift);
                    } else bkt = (u32)((home - t.lo) >> shift);
                    u32 rank = atomicAdd(&hist[bkt], 1u);
                    if (MODE == 1) {
                        u64 pos = (u64)base[bkt] + rank;
                        u64 meta = (o[g].ord << 8) | (o[g].rb << 4) | o[g].lb;
                        u64 *tuples = dst_ptrs ? dst_ptrs[bkt] : this->tuples;      // peer (or own) receive buffer
                        if (WIDE) {
                            ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(tuples) + 2 * pos;
                            __stcg(dst, make_ulonglong2(o[g].klo, o[g].khi));
                            __stcg(dst + 1, make_ulonglong2(meta, 0ULL));
                        } else {
                            __stcg(reinterpret_cast<ulonglong2 *>(tuples) + pos, make_ulonglong2(o[g].klo, meta));
                        }
                    }
                }
            }
        }
    }

    __device__ __forceinline__ void finish()
    {
        if (MODE == 0) {
            __syncthreads();
            u32 *row = matrix + (size_t)blockIdx.x * n_buckets;
            for (u32 b = threadIdx.x; b < n_buckets; b += BLOCK) row[b] = hist[b];
        }
    }
};

// Shared-memory staging for the scatter passes: tuples are appended to a batch in shared memory (with their bucket
// and their rank inside the batch's bucket), and when the batch is full it is copied out IN BUCKET ORDER, adjacent
// lanes writing adjacent tuples.  A bucket then receives runs of several tuples (whole 32-B sectors) per batch
// instead of isolated 16-B half-sector stores -- the scattered-store transaction rate, not bandwidth, bounds the
// scatter (profiles/README.md).
template <bool WIDE>
struct StageBuf {
    static constexpr int TW = WIDE ? 4 : 2;   // u64 words per tuple
    u64 *tup;            // [cap * TW]
    unsigned short *bk;  // [cap] bucket of entry e
    unsigned short *rk;  // [cap] rank of entry e inside its bucket (this batch)
    unsigned short *idx; // [cap] entry at sorted position i
    u32 *bh;             // [nb] batch counts
    u32 *boff;           // [nb] exclusive scan of bh
    u32 *base;           // [nb] running global write offset of this CTA per bucket
    u32 *misc;           // [0] cursor, [1..8] warp totals
    u32 cap, nb;

    static __host__ __device__ size_t bytes(u32 cap, u32 nb)
    {
        return (size_t)cap * TW * 8 + (size_t)cap * 3 * 2 + (size_t)nb * 3 * 4 + 16 * 4;
    }
    __device__ __forceinline__ void carve(void *mem, u32 cap_, u32 nb_)
    {
        cap = cap_; nb = nb_;
        tup = reinterpret_cast<u64 *>(mem);
        bh = reinterpret_cast<u32 *>(tup + (size_t)cap * TW);
        boff = bh + nb; base = boff + nb; misc = base + nb;
        bk = reinterpret_cast<unsigned short *>(misc + 16);
        rk = bk + cap; idx = rk + cap;
    }
    // all threads of the CTA; a __syncthreads() must separate the last append from this call; ends with one.
    // OPT (optimistic single-pass partition): no precomputed offsets; every bucket has a fixed region of `capb`
    // tuples and the batch reserves its run with one atomicAdd per touched bucket on fill[]; a bucket that would
    // overflow raises *flag and its tuples are dropped (the host then redoes the block with the exact two-pass
    // partition).
    // dst_ptrs != nullptr (optimistic peer exchange): bucket b lives in its own buffer dst_ptrs[b] (the owner's receive
    // buffer, over NVLink), and this source's region starts at tuple `region_off` of every one of them
    template <bool OPT = false>
    __device__ __forceinline__ void flush(u64 *dst, u32 *fill = nullptr, u32 capb = 0, u32 *flag = nullptr, u64 *const *dst_ptrs = nullptr,
                                          u64 region_off = 0)
    {
        const u32 t = threadIdx.x, nthr = blockDim.x;
        const u32 n = misc[0];
        // OPT: reserve this batch's run in every touched bucket's region.  The atomics are issued first and their
        // results are picked up after the scan below, so the round trips hide behind it (nb <= 8 * nthr: up to 2048
        // buckets, the (owner, slice) buckets of the pull exchange on 8 GPUs included; larger bucket counts take the
        // results at once, chunk by chunk)
        constexpr int RRN = 8;       // reservations a thread holds in registers across the scan: nb <= 8 * nthr buckets
        u32 rr[RRN];
#pragma unroll
        for (int j = 0; j < RRN; j++) rr[j] = 0;
        const bool held = nb <= RRN * nthr;
        if (OPT) {
            if (held) {
#pragma unroll
                for (int j = 0; j < RRN; j++) {
                    if (j * nthr >= nb) break;                 // block-uniform: 382 buckets (one GPU) take two rounds, not eight
                    const u32 b = j * nthr + t;
                    const u32 cnt = b < nb ? bh[b] : 0u;
                    if (cnt) rr[j] = atomicAdd(&fill[b], cnt);
                }
            } else {
                for (u32 c0 = 0; c0 < nb; c0 += 4 * nthr) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const u32 b = c0 + j * nthr + t;
                        const u32 cnt = b < nb ? bh[b] : 0u;
                        rr[j] = cnt ? atomicAdd(&fill[b], cnt) : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const u32 b = c0 + j * nthr + t;
                        const u32 cnt = b < nb ? bh[b] : 0u;
                        if (cnt) { if (rr[j] + cnt > capb) { *flag = 1u; base[b] = 0xffffffffu; } else base[b] = (dst_ptrs ? 0u : b * capb) + rr[j]; }
                    }
                }
            }
        }
        // exclusive scan of bh (nb <= 4096): a contiguous segment per thread, warp scan, warp totals
        const u32 seg = (nb + nthr - 1) / nthr;
        const u32 b0 = t * seg, b1 = (b0 + seg < nb) ? b0 + seg : nb;
        u32 sum = 0;
        for (u32 b = b0; b < b1; b++) sum += bh[b];
        u32 incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, d); if ((t & 31) >= (u32)d) incl += v; }
        if ((t & 31) == 31) misc[1 + (t >> 5)] = incl;
        __syncthreads();
        u32 run = incl - sum;
        for (u32 w = 0; w < (t >> 5); w++) run += misc[1 + w];
        for (u32 b = b0; b < b1; b++) { boff[b] = run; run += bh[b]; }
        if (OPT && held) {
#pragma unroll
            for (int j = 0; j < RRN; j++) {
                if (j * nthr >= nb) break;
                const u32 b = j * nthr + t;
                const u32 cnt = b < nb ? bh[b] : 0u;
                if (cnt) { if (rr[j] + cnt > capb) { *flag = 1u; base[b] = 0xffffffffu; } else base[b] = (dst_ptrs ? 0u : b * capb) + rr[j]; }
            }
        }
        __syncthreads();
        for (u32 e = t; e < n; e += nthr) idx[boff[bk[e]] + rk[e]] = (unsigned short)e;
        __syncthreads();
        for (u32 i = t; i < n; i += nthr) {
            const u32 e = idx[i], b = bk[e];
            if (OPT && base[b] == 0xffffffffu) continue;
            const u64 pos = (u64)base[b] + (i - boff[b]) + region_off;
            const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(tup);
            if (OPT && dst_ptrs) dst = dst_ptrs[b];
            if (WIDE) {
                ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dst) + 2 * pos;
                __stcg(d, q[2 * e]);
                __stcg(d + 1, q[2 * e + 1]);
            } else {
                __stcg(reinterpret_cast<ulonglong2 *>(dst) + pos, q[e]);
            }
        }
        __syncthreads();
        for (u32 b = t; b < nb; b += nthr) { if (!OPT) base[b] += bh[b]; bh[b] = 0; }
        if (t == 0) misc[0] = 0;
        __syncthreads();
    }
    // append one tuple at entry e (reserved by the caller)
    __device__ __forceinline__ void put(u32 e, u32 b, u64 klo, u64 khi, u64 meta)
    {
        ulonglong2 *q = reinterpret_cast<ulonglong2 *>(tup);          // tup is 16-B aligned: one 128-bit store per half
        if (WIDE) { q[2 * e] = make_ulonglong2(klo, khi); q[2 * e + 1] = make_ulonglong2(meta, 0ULL); }
        else q[e] = make_ulonglong2(klo, meta);
        bk[e] = (unsigned short)b;
        rk[e] = (unsigned short)atomicAdd(&bh[b], 1u);
    }
};

// scatter pass of the single-GPU partitioned build through the staging above (slice buckets)
template <bool WIDE, bool OPT = false>
struct StagedScatterSink {
    static constexpr int RUN = G;
#ifndef DBG_SCATTER_MIN_BLOCKS
#define DBG_SCATTER_MIN_BLOCKS 3
#endif
    static constexpr int MIN_BLOCKS = DBG_SCATTER_MIN_BLOCKS;   // 3: <= 85 registers, three CTAs per SM next to a 2048-tuple batch
    TableView t;
    int shift;
    u32 n_buckets, cap;
    u32 *matrix;       // exact: [n_chunks][n_buckets] write offsets
    u64 *tuples;
    u32 *fill;         // OPT: tuples stored so far per bucket (global)
    u32 capb;          // OPT: region size per bucket
    u32 *flag;         // OPT: overflow
    // optimistic PEER exchange (multi-GPU, OPT only): buckets = owner ranks (home / div), bucket b is stored into the
    // owner's receive buffer dst_ptrs[b] over NVLink, inside the region [region_off, region_off + capb) that owner keeps
    // for THIS source -- no counting pass, no offsets exchanged beforehand
    u64 div = 0, div_M = 0;
    u64 *const *dst_ptrs = nullptr;
    u64 region_off = 0;
    // PULL exchange (OPT only, dst_ptrs == nullptr): buckets = (owner, table slice of the owner) = owner * nb_local + slice, all
    // stored into this source's LOCAL send buffer `tuples`; the owners read their regions over NVLink (k_insert_tuples_pull)
    u32 nb_local = 0;
    StageBuf<WIDE> sb;
    u32 filled;

    __device__ __forceinline__ void init(u32 *extra)
    {
        filled = 0;
        sb.carve(extra, cap, n_buckets);
        const u32 *row = OPT ? nullptr : matrix + (size_t)blockIdx.x * n_buckets;
        for (u32 b = threadIdx.x; b < n_buckets; b += BLOCK) { sb.bh[b] = 0; sb.base[b] = OPT ? 0u : row[b]; }
        if (threadIdx.x == 0) sb.misc[0] = 0;
        __syncthreads();
    }

    __device__ __forceinline__ void consume(const Occ (&o)[RUN], int nv)
    {
        // `filled` is a block-uniform upper bound of the batch cursor kept in registers (never read the shared
        // cursor to decide: a fast warp may already have bumped it for this round)
        if (filled + BLOCK * RUN > sb.cap) { __syncthreads(); sb.template flush<OPT>(tuples, fill, capb, flag, dst_ptrs, region_off); filled = 0; }
        filled += BLOCK * RUN;
        u32 bkt[RUN];
        u32 mine = 0;
#pragma unroll
        for (int g = 0; g < RUN; g++) {
            bkt[g] = 0xffffffffu;
            if (g < nv) {
                if ((o[g].klo | o[g].khi) == 0) {
                    polyA_bump(t.polyA, o[g].lb, o[g].rb, o[g].ord);
                } else {
                    u64 hh = WIDE ? hash_code_wide(o[g].klo, o[g].khi) : hash_code(o[g].klo);
                    const u64 home = mod_P(hh, t.P, t.M);
                    if (OPT && div) {
                        u32 b = (u32)__umul64hi(home, div_M); if ((u64)(b + 1) * div <= home) b++;
                        if (nb_local) b = b * nb_local + (u32)((home - (u64)b * div) >> shift);
                        bkt[g] = b;
                    } else bkt[g] = (u32)((home - t.lo) >> shift);
                    mine++;
                }
            }
        }
        // reserve entries: warp-aggregated bump of the batch cursor
        u32 incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= (u32)d) incl += v; }
        u32 wbase = 0;
        if ((threadIdx.x & 31) == 31 && incl) wbase = atomicAdd(&sb.misc[0], incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        u32 e = wbase + incl - mine;
#pragma unroll
        for (int g = 0; g < RUN; g++)
            if (bkt[g] != 0xffffffffu) sb.put(e++, bkt[g], o[g].klo, o[g].khi, (o[g].ord << 8) | (o[g].rb << 4) | o[g].lb);
    }

    __device__ __forceinline__ void finish()
    {
        __syncthreads();
        sb.template flush<OPT>(tuples, fill, capb, flag, dst_ptrs, region_off);
    }
};

// Scatter pass of the FUSED multi-GPU exchange, owner buckets only (n_buckets = ranks <= 32).  Every round of
// BLOCK*G occurrences is first sorted by owner in shared memory, then copied out with adjacent lanes writing
// adjacent 16-B tuples: each owner receives contiguous runs (hundreds of bytes) instead of isolated 16-B stores,
// which is what NVLink wants (a store per tuple straight from the generator reached only a fraction of the link
// rate: every 16-B store became its own packet).
template <bool WIDE>
struct PeerStagedSink {
    static constexpr int RUN = G;
    static constexpr int MIN_BLOCKS = MIN_CTAS;
    static constexpr int TW = WIDE ? 4 : 2;            // u64 words per tuple
    TableView t;
    u64 div, div_M;
    u32 n_buckets;
    u32 *matrix;            // [n_chunks][n_buckets] write offsets (after the scan)
    u64 *const *dst_ptrs;   // owner receive buffers (peer mappings)
    const u64 *dst_base;    // first tuple index reserved for this rank in each owner's buffer
    const u64 *roffs;       // this rank's packed start of each bucket (subtracted from the matrix offsets)
    u32 *cnt;               // shared [2][32] round counts (ping-pong)
    u32 *base;              // shared [32] running write offset per owner
    u64 *stage;             // shared [BLOCK*G] tuples
    u32 parity;

    static __host__ __device__ size_t smem_bytes() { return (size_t)(3 * 32) * sizeof(u32) + (size_t)BLOCK * G * TW * sizeof(u64); }

    __device__ __forceinline__ void init(u32 *extra)
    {
        cnt = extra; base = extra + 64;
        stage = reinterpret_cast<u64 *>(extra + 96);
        const u32 *row = matrix + (size_t)blockIdx.x * n_buckets;
        if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
        if (threadIdx.x < n_buckets) base[threadIdx.x] = (u32)((u64)row[threadIdx.x] - roffs[threadIdx.x] + dst_base[threadIdx.x]);
        parity = 0;
        __syncthreads();
    }

    __device__ __forceinline__ void consume(const Occ (&o)[RUN], int nv)
    {
        u32 *c = cnt + 32 * parity;
        u32 bkt[RUN], rank[RUN];
#pragma unroll
        for (int g = 0; g < RUN; g++) {
            bkt[g] = 0xffffffffu; rank[g] = 0;
            if (g < nv) {
                if ((o[g].klo | o[g].khi) == 0) {     // the k-mer-0 side node is accumulated here, once
                    polyA_bump(t.polyA, o[g].lb, o[g].rb, o[g].ord);
                } else {
                    u64 hh = WIDE ? hash_code_wide(o[g].klo, o[g].khi) : hash_code(o[g].klo);
                    u64 home = mod_P(hh, t.P, t.M);
                    u32 b = (u32)__umul64hi(home, div_M); if ((u64)(b + 1) * div <= home) b++;
                    bkt[g] = b;
                    rank[g] = atomicAdd(&c[b], 1u);
                }
            }
        }
        __syncthreads();                                   // round counted
        u32 eoff[RUN] = {};                                // exclusive prefix of the round's counts at my owners
        u32 total = 0;
        for (u32 b = 0; b < n_buckets; b++) {
#pragma unroll
            for (int g = 0; g < RUN; g++) if (bkt[g] == b) eoff[g] = total;
            total += c[b];
        }
#pragma unroll
        for (int g = 0; g < RUN; g++) {
            if (bkt[g] != 0xffffffffu) {
                const u32 e = eoff[g] + rank[g];
                u64 meta = (o[g].ord << 8) | (o[g].rb << 4) | o[g].lb;
                if (WIDE) { stage[4 * e] = o[g].klo; stage[4 * e + 1] = o[g].khi; stage[4 * e + 2] = meta; stage[4 * e + 3] = 0; }
                else { stage[2 * e] = o[g].klo; stage[2 * e + 1] = meta; }
            }
        }
        __syncthreads();                                   // round sorted by owner in shared memory
        for (u32 e = threadIdx.x; e < total; e += BLOCK) {
            u32 b = 0, lo = 0;
            for (; b + 1 < n_buckets; b++) { const u32 cq = c[b]; if (e < lo + cq) break; lo += cq; }
            const u64 pos = (u64)base[b] + (e - lo);
            if (WIDE) {
                ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(dst_ptrs[b]) + 2 * pos;
                __stcg(dst, make_ulonglong2(stage[4 * e], stage[4 * e + 1]));
                __stcg(dst + 1, make_ulonglong2(stage[4 * e + 2], 0ULL));
            } else {
                __stcg(reinterpret_cast<ulonglong2 *>(dst_ptrs[b]) + pos, make_ulonglong2(stage[2 * e], stage[2 * e + 1]));
            }
        }
        __syncthreads();                                   // copied out: stage, base and this round's counts are free
        if (threadIdx.x < n_buckets) { base[threadIdx.x] += c[threadIdx.x]; c[threadIdx.x] = 0; }
        parity ^= 1;                                       // next round counts in the other slot while these reset
    }

    __device__ __forceinline__ void finish() {}
};

// column-wise exclusive scan of M[n_chunks][n_buckets] (+ per-bucket bases), three small kernels:
//   1: tile sums  ts[tile][b] = sum of M[c][b] over the tile's chunks            grid (tiles, ceil(nb/256))
//   2: per bucket: exclusive scan of ts over tiles, bucket totals -> boffs (exclusive over buckets), one CTA
//   3: M[c][b] <- boffs[b] + ts_excl[tile][b] + running sum inside the tile       grid like 1
constexpr int PT_CHUNKS = 128;     // chunks per scan tile

static __global__ void k_offsets_to_counts(const u64 *__restrict__ offs, u32 n, u64 *counts)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = offs[i + 1] - offs[i];
}

static __global__ void __launch_bounds__(256) k_part_scan1(const u32 *__restrict__ matrix, u64 n_chunks, u32 nb, u64 *__restrict__ ts)
{
    const u32 b = blockIdx.y * 256 + threadIdx.x;
    if (b >= nb) return;
    const u64 c0 = (u64)blockIdx.x * PT_CHUNKS, c1 = c0 + PT_CHUNKS < n_chunks ? c0 + PT_CHUNKS : n_chunks;
    u64 sum = 0;
    for (u64 c = c0; c < c1; c++) sum += matrix[c * nb + b];
    ts[(u64)blockIdx.x * nb + b] = sum;
}

static __global__ void __launch_bounds__(1024) k_part_scan2(u64 *ts, u64 n_tiles, u32 nb, u64 *boffs)
{
    __shared__ u64 wsum[32];
    __shared__ u64 carry;
    const int tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (u32 b0 = 0; b0 < nb; b0 += 1024) {
        const u32 b = b0 + tid;
        u64 tot = 0;
        if (b < nb)
            for (u64 tl = 0; tl < n_tiles; tl++) { u64 v = ts[tl * nb + b]; ts[tl * nb + b] = tot; tot += v; }   // exclusive over tiles
        u64 inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { u64 x = __shfl_up_sync(0xffffffffu, inc, s); if ((tid & 31) >= s) inc += x; }
        if ((tid & 31) == 31) wsum[tid >> 5] = inc;
        __syncthreads();
        u64 wb = 0;
        for (int w = 0; w < (tid >> 5); w++) wb += wsum[w];
        if (b < nb) boffs[b] = carry + wb + inc - tot;
        __syncthreads();
        if (tid == 1023) carry += wb + inc;
        __syncthreads();
    }
    if (tid == 0) boffs[nb] = carry;
}

static __global__ void __launch_bounds__(256) k_part_scan3(u32 *matrix, u64 n_chunks, u32 nb, const u64 *__restrict__ ts, const u64 *__restrict__ boffs)
{
    const u32 b = blockIdx.y * 256 + threadIdx.x;
    if (b >= nb) return;
    const u64 c0 = (u64)blockIdx.x * PT_CHUNKS, c1 = c0 + PT_CHUNKS < n_chunks ? c0 + PT_CHUNKS : n_chunks;
    u64 run = boffs[b] + ts[(u64)blockIdx.x * nb + b];
    for (u64 c = c0; c < c1; c++) { u32 v = matrix[c * nb + b]; matrix[c * nb + b] = (u32)run; run += v; }
}

// ---------------------------------------------------------------------------------------------------
// the fused build kernel: one CTA per chunk of CB bases
// ---------------------------------------------------------------------------------------------------
template <bool WIDE, class Sink>
__global__ void __launch_bounds__(BLOCK, Sink::MIN_BLOCKS) k_build(BuildArgs a, Sink sink)
{
    extern __shared__ u32 smem[];
    u32 *pk = smem;                          // a.stage_words
    u32 *rstart = pk + a.stage_words;        // MAXR
    u32 *rpre = rstart + MAXR;               // MAXR + 1
    u32 *extra = rpre + MAXR + 2;            // sink-private shared memory, 16-byte aligned
    extra += (4 - ((a.stage_words + 2 * MAXR + 2) & 3)) & 3;
    __shared__ u32 warp_tot[BLOCK / 32];

    const int tid = threadIdx.x;
    const u64 chunk = blockIdx.x;
    const u64 cbase = a.abase + chunk * CB;
    const int K = a.K;
    constexpr int RUN = Sink::RUN;

    sink.init(extra);

    // ---- (1) coalesced 16-B loads of ASCII bases, 2-bit pack, stage in shared memory -------------------
    {
        u64 avail = a.end_base > cbase ? a.end_base - cbase : 0;
        u64 want = (u64)(a.stage_words - 4) * 16;
        u32 nvec = (u32)(((avail < want ? avail : want) + 15) / 16);
        const uint4 *src = reinterpret_cast<const uint4 *>(a.bases + cbase);
        for (u32 v = tid; v < a.stage_words; v += BLOCK)
            pk[v] = (v < nvec) ? pack16(__ldg(src + v)) : 0u;
    }

    const u64 r_lo = a.chunk_first[chunk], r_hi = a.chunk_first[chunk + 1];
    u64 logged = 0, n_occ = 0;

    for (u64 rb = r_lo; rb < r_hi; rb += MAXR) {
        const u32 nr = (u32)((r_hi - rb) < (u64)MAXR ? (r_hi - rb) : (u64)MAXR);
        __syncthreads();   // staging done / previous pass finished with the tables

        // ---- (2) per-read k-mer counts + exclusive scan (4 reads per thread) ---------------------------
        u32 c[4]; u32 tsum = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            u32 i = tid * 4 + e;
            c[e] = 0;
            if (i < nr) {
                u64 s0 = a.offs[rb + i], s1 = a.offs[rb + i + 1];
                u64 len = s1 - s0;
                u64 klen = len < (u64)a.R ? len : (u64)a.R;                 // trim to -r (DBGgraph.cpp:63)
                if (klen >= (u64)K) c[e] = (u32)(klen - K + 1);             // skip reads shorter than K (:51-53)
                if (len >= (u64)K) logged += len - K + 1;                   // Kmer_total_num quirk (:101)
                rstart[i] = (u32)(s0 - cbase);
            }
            tsum += c[e];
        }
        u32 incl = tsum;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, s); if ((tid & 31) >= s) incl += v; }
        if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        u32 wbase = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; w++) if (w < (tid >> 5)) wbase += warp_tot[w];
        u32 ex = wbase + incl - tsum;
#pragma unroll
        for (int e = 0; e < 4; e++) { rpre[tid * 4 + e] = ex; ex += c[e]; }
        if (tid == BLOCK - 1) rpre[MAXR] = ex;
        __syncthreads();
        const u32 S = rpre[MAXR];
        if (tid == 0) n_occ += S;

        // ---- (3) occurrences: each thread takes runs of RUN consecutive ones ----------------------------
        // (block-uniform trip count: every thread calls sink.consume, so sinks may use full-warp collectives)
        for (u32 ob = 0; ob < S; ob += BLOCK * RUN) {
            const u32 o0 = ob + tid * RUN;
            Occ occ[RUN];
            int nv = 0;
            if (o0 < S) {
            u32 lo = 0, hi = nr;
            while (lo < hi) { u32 mid = (lo + hi) >> 1; if (rpre[mid] <= o0) lo = mid + 1; else hi = mid; }
            u32 i = lo - 1;
            u32 j = o0 - rpre[i];
            u32 ci = rpre[i + 1] - rpre[i];
            bool fresh = true;
            u64 flo = 0, fhi = 0, rlo = 0, rhi = 0;   // forward / reverse-complement words
#pragma unroll
            for (int g = 0; g < RUN; g++) {
                if (o0 + g < S) {
                    if (j >= ci) {
                        do { i++; ci = rpre[i + 1] - rpre[i]; } while (ci == 0);
                        j = 0; fresh = true;
                    }
                    const u32 p = rstart[i] + j;
                    if (fresh) {
                        // first k-mer of a run: read the window straight out of the packed stream
                        if (WIDE) { U128 f = window128(pk, p, K); U128 r = revcomp128(f, K); flo = f.lo; fhi = f.hi; rlo = r.lo; rhi = r.hi; }
                        else { flo = window64(pk, p, K); rlo = revcomp64(flo, K); }
                        fresh = false;
                    } else {
                        // rolling update (DBGgraph.cpp:71-73)
                        u32 b = code_at(pk, p + K - 1);
                        if (WIDE) {
                            fhi = (fhi << 2) | (flo >> 62); flo = (flo << 2) | b;
                            int top = 2 * K - 64;   // bits of the k-mer living in the high word (K > 32) or <= 0
                            if (top > 0) fhi &= (top >= 64 ? ~0ULL : ((1ULL << top) - 1));
                            else { fhi = 0; if (2 * K < 64) flo &= (1ULL << (2 * K)) - 1; }
                            rlo = (rlo >> 2) | (rhi << 62); rhi >>= 2;
                            int sh = 2 * (K - 1);
                            if (sh >= 64) rhi |= (u64)(3 - b) << (sh - 64); else rlo |= (u64)(3 - b) << sh;
                        } else {
                            flo = ((flo << 2) | b) & ((1ULL << (2 * K)) - 1);
                            rlo = (rlo >> 2) | ((u64)(3 - b) << (2 * (K - 1)));
                        }
                    }
                    u32 left = (j > 0) ? code_at(pk, p - 1) : 4u;
                    u32 right = (j + 1 < ci) ? code_at(pk, p + K) : 4u;
                    bool fwd = WIDE ? (fhi < rhi || (fhi == rhi && flo <= rlo)) : (flo <= rlo);   // tie -> forward (:80)
                    Occ &q = occ[g];
                    if (fwd) { q.klo = flo; q.khi = fhi; q.lb = left; q.rb = right; }
                    else {
                        q.klo = rlo; q.khi = rhi;
                        q.rb = (left < 4) ? 3 - left : 4u;       // DBGgraph.cpp:85-89
                        q.lb = (right < 4) ? 3 - right : 4u;
                    }
                    q.ord = ((a.read_index0 + rb + i) << 16) | j;
                    if (a.seed) {
                        // chop_contig_to_kmerset, map_func.cpp:154-163: direct = 1 only for kbit < rc_kbit (a tie is direct 0)
                        const bool direct = WIDE ? (fhi < rhi || (fhi == rhi && flo < rlo)) : (flo < rlo);
                        q.lb = 0; q.rb = RB_SEED; q.ord = (q.ord << 1) | (direct ? 1u : 0u);
                    }
                    j++; nv = g + 1;
                }
            }
            }
            sink.consume(occ, nv);
        }
    }

    sink.finish();
    // Kmer_total_num / occurrence counters (skipped by the counting pass of the partitioned build)
    if (a.count_stats) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) logged += __shfl_xor_sync(0xffffffffu, logged, s);
        if ((tid & 31) == 0 && logged) atomicAdd(sink.t.counters + CNT_LOGGED, logged);
        if (tid == 0 && n_occ) atomicAdd(sink.t.counters + CNT_OCC, n_occ);
    }
}

// ---------------------------------------------------------------------------------------------------
// owner side of the multi-GPU exchange: received tuples arrive in source order; the same exact, atomic-free
// radix partition (count rows -> column scan -> scatter) puts them in table-slice order for the bucketed insert
// ---------------------------------------------------------------------------------------------------
#ifndef DBG_TP_TILE
#define DBG_TP_TILE 2048
#endif
constexpr int TP_TILE = DBG_TP_TILE;     // tuples per CTA (2048: 50 KB of staging, four CTAs per SM; 4096 left only two resident)
constexpr int TP_CTAS = TP_TILE <= 2048 ? 4 : 2;

// after an optimistic scatter that did not overflow: bucket b occupies [b*capb, b*capb + fill[b]); publish the
// strided bucket offsets the insert kernel walks and zero the keys behind the last tuple of every region (the insert
// kernel skips tuples whose key is 0: tuples never carry the k-mer-0 key)
template <bool WIDE>
__global__ void __launch_bounds__(256) k_opt_finish(u64 *tuples, const u32 *__restrict__ fill, u32 capb, u32 nb, u64 *boffs, u32 tile)
{
    const u32 b = blockIdx.x;
    if (threadIdx.x == 0) { boffs[b] = (u64)b * capb; if (b == nb - 1) boffs[nb] = (u64)nb * capb; }
    const u32 f = fill[b];
    ulonglong2 *t = reinterpret_cast<ulonglong2 *>(tuples);
    // capb is a multiple of the insert tile, so regions start on tile boundaries; the insert kernel skips the tiles
    // that lie entirely in the unused tail, only the partly used tile needs its tail zeroed
    const u32 z_end = min(capb, (f + tile - 1) / tile * tile);
    for (u32 i = f + threadIdx.x; i < z_end; i += 256) {
        const u64 pos = (u64)b * capb + i;
        if (WIDE) t[2 * pos] = make_ulonglong2(0ULL, 0ULL);
        else t[pos] = make_ulonglong2(0ULL, 0ULL);
    }
}

// scatter pass of the tuple partition through the shared-memory staging (one batch = the CTA's tile of TP_TILE tuples)
template <bool WIDE, bool OPT>
__global__ void __launch_bounds__(256, TP_CTAS) k_tuple_scatter_staged(const u64 *__restrict__ src, u64 n, TableView t, int shift, u32 nb,
                                                              const u32 *__restrict__ matrix, u64 *dst, u32 *fill, u32 capb, u32 *flag)
{
    extern __shared__ u32 tp_smem[];
    StageBuf<WIDE> sb;
    sb.carve(tp_smem, TP_TILE, nb);
    const u32 *row = OPT ? nullptr : matrix + (size_t)blockIdx.x * nb;
    for (u32 b = threadIdx.x; b < nb; b += 256) { sb.bh[b] = 0; sb.base[b] = OPT ? 0u : row[b]; }
    const u64 i0 = (u64)blockIdx.x * TP_TILE;
    const u32 cnt = (u32)((n - i0) < (u64)TP_TILE ? (n - i0) : (u64)TP_TILE);
    if (threadIdx.x == 0) sb.misc[0] = cnt;
    __syncthreads();
    for (u32 e = threadIdx.x; e < cnt; e += 256) {
        const u64 i = i0 + e;
        u64 klo, khi = 0, meta, z = 0;
        if (WIDE) ld256_cs(reinterpret_cast<const ulonglong2 *>(src) + 2 * i, klo, khi, meta, z);
        else { ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2 *>(src) + i); klo = x.x; meta = x.y; }
        u64 h = WIDE ? hash_code_wide(klo, khi) : hash_code(klo);
        sb.put(e, (u32)((mod_P(h, t.P, t.M) - t.lo) >> shift), klo, khi, meta);
    }
    __syncthreads();
    sb.template flush<OPT>(dst, fill, capb, flag);
}

template <bool WIDE, int MODE>
__global__ void __launch_bounds__(256) k_tuple_partition(const u64 *__restrict__ src, u64 n, TableView t, int shift, u32 nb,
                                                         u32 *matrix, u64 *dst)
{
    extern __shared__ u32 tp_smem[];
    u32 *hist = tp_smem, *base = tp_smem + nb;
    u32 *row = matrix + (size_t)blockIdx.x * nb;
    for (u32 b = threadIdx.x; b < nb; b += 256) { hist[b] = 0; if (MODE == 1) base[b] = row[b]; }
    __syncthreads();
    const u64 i0 = (u64)blockIdx.x * TP_TILE;
    for (int r = 0; r < TP_TILE / 256; r++) {
        const u64 i = i0 + (u64)r * 256 + threadIdx.x;
        if (i >= n) break;
        u64 klo, khi = 0, meta, z = 0;
        if (WIDE) ld256_cs(reinterpret_cast<const ulonglong2 *>(src) + 2 * i, klo, khi, meta, z);
        else { ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2 *>(src) + i); klo = x.x; meta = x.y; }
        u64 h = WIDE ? hash_code_wide(klo, khi) : hash_code(klo);
        u32 bkt = (u32)((mod_P(h, t.P, t.M) - t.lo) >> shift);
        u32 rank = atomicAdd(&hist[bkt], 1u);
        if (MODE == 1) {
            u64 pos = (u64)base[bkt] + rank;
            if (WIDE) {
                ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dst) + 2 * pos;
                d[0] = make_ulonglong2(klo, khi); d[1] = make_ulonglong2(meta, 0ULL);
            } else {
                reinterpret_cast<ulonglong2 *>(dst)[pos] = make_ulonglong2(klo, meta);
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (u32 b = threadIdx.x; b < nb; b += 256) row[b] = hist[b];
    }
}

// ---------------------------------------------------------------------------------------------------
// owner side: insert tuples (partitioned build and multi-GPU exchange)
// ---------------------------------------------------------------------------------------------------
// Persistent CTAs; CTA c takes tiles c, c+grid, ... of 256 tuples, ONE tuple per thread per tile: the kernel is
// a memory-latency machine (a dependent tuple -> node -> [claim] chain per occurrence), so it keeps the code
// and the register footprint small and the number of independent chains per SM at the hardware maximum.
// Partitioned build: tuples are in bucket order, so all CTAs work inside a window of grid*256 tuples, i.e. one
// or two 16-MB table slices that stay in the 126 MB L2.  Random first touches of a slice would still reach
// DRAM one sector at a time; instead every tile also streams its proportional share of the NEXT slice into
// L2 with coalesced prefetches.
constexpr int INS_BLOCK = 256;
#ifndef DBG_INS_ROUNDS
#define DBG_INS_ROUNDS 2
#endif
constexpr int INS_ROUNDS = DBG_INS_ROUNDS;            // tuples per thread per tile
constexpr int INS_TILE = INS_BLOCK * INS_ROUNDS;
#ifndef DBG_INS_CTAS
#define DBG_INS_CTAS 8
#endif
constexpr int INS_CTAS = DBG_INS_CTAS;      // per SM (x8 warps)

template <bool WIDE, bool TRACK>
__global__ void __launch_bounds__(INS_BLOCK, INS_CTAS) k_insert_tuples(const u64 *__restrict__ tuples, u64 n, const u64 *__restrict__ n_ptr,
                                                                       TableView t, const u64 *__restrict__ boffs, u32 n_buckets, int shift,
                                                                       u64 *tile_counter, const u32 *__restrict__ fill,
                                                                       u64 t_begin = 0, u32 b_begin = 0)
{
    // (t_begin, b_begin: a launch over the bucket range [b_begin, ...) of the stream, tuples [t_begin, n) -- the grouped
    // insert of dbg_finish_export; boffs / fill / n stay in whole-stream coordinates)
    if (n_ptr) n = *n_ptr;      // exact count produced on the device (partitioned build): no host round trip
    u32 n_new = 0, n_conf = 0;
    u32 b = b_begin;            // current bucket of this CTA's tile (monotone)
    u64 b0 = 0, b1 = 0;         // the bucket's region in the tuple stream
    u64 used = 0;               // tuples actually stored in it (== b1 - b0 unless the regions are fixed-size: `fill`)
    float ratio = 0.f;
    // lines of the NEXT table slice per tuple of the current bucket: every tile prefetches its share
    auto next_slice_ratio = [&]() -> float {
        if (!(b + 1 < n_buckets && used > 0 && ((u64)(b + 1) << shift) < t.n_local)) return 0.f;
        const u64 slice_lo = (u64)(b + 1) << shift;
        u64 slice_n = (u64)1 << shift;
        if (slice_lo + slice_n > t.n_local) slice_n = t.n_local - slice_lo;
        return (float)(slice_n * sizeof(NodeT<WIDE>) / 128) / (float)used;      // 128-B L2 lines per tuple
    };
    if (boffs) {
        b0 = __ldg(boffs + b); b1 = __ldg(boffs + b + 1);
        used = fill ? (u64)__ldg(fill + b) : b1 - b0;
        ratio = next_slice_ratio();
    }
    __shared__ u64 s_tile[2];
    // tiles are handed out by a global counter, so at any moment the resident CTAs hold the NEXT gridDim tiles of
    // the bucket-ordered stream: the window of table slices they touch cannot drift apart (static round-robin
    // let fast CTAs run buckets ahead and the slices fell out of L2).  The counter is read one tile ahead (its
    // round trip hides behind the current tile) and published through a ping-pong slot: one barrier per tile.
    // (a table already found full: nothing to do; a probe that finds it full poisons the tile counter, see insert_probe)
    u64 next_tile = 0;
    if (threadIdx.x == 0) next_tile = __ldcg(t.counters + CNT_ERROR) ? (1ULL << 40) : atomicAdd(tile_counter, 1ULL);
    for (u32 par = 0;; par ^= 1) {
        if (threadIdx.x == 0) s_tile[par] = next_tile;
        __syncthreads();
        const u64 tile = t_begin + s_tile[par] * INS_TILE;
        if (tile >= n) break;
        if (threadIdx.x == 0) next_tile = atomicAdd(tile_counter, 1ULL);
        if (boffs && tile >= b1 && b + 1 < n_buckets) {
            // entered a new bucket (rare path): its region, its tuple count, the prefetch ratio of the next slice
            while (b + 1 < n_buckets && tile >= b1) { b++; b0 = b1; b1 = __ldg(boffs + b + 1); }
            used = fill ? (u64)__ldg(fill + b) : b1 - b0;
            ratio = next_slice_ratio();
        }
        if (fill && tile - b0 >= used) continue;        // fixed-size regions: this tile lies in the unused tail (block-uniform)
        u64 klo[INS_ROUNDS], khi[INS_ROUNDS], meta[INS_ROUNDS];
#pragma unroll
        for (int r = 0; r < INS_ROUNDS; r++) {
            const u64 i = tile + (u64)r * INS_BLOCK + threadIdx.x;
            klo[r] = 0; khi[r] = 0; meta[r] = 0;
            if (i < n) {
                if (WIDE) { u64 z; ld256_cs(reinterpret_cast<const ulonglong2 *>(tuples) + 2 * i, klo[r], khi[r], meta[r], z); }   // read once: evict first
                else { ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2 *>(tuples) + i); klo[r] = x.x; meta[r] = x.y; }
            }
        }
        if (ratio > 0.f) {
            // this tile's proportional share of the next slice (approximate shares are fine: overlaps and small
            // gaps only cost a few redundant or late lines)
            const float t0 = (float)(tile - b0);
            const u64 l0 = (u64)(t0 * ratio), l1 = (u64)((t0 + (float)INS_TILE) * ratio) + 1;
            const char *basep = reinterpret_cast<const char *>(static_cast<const NodeT<WIDE> *>(t.nodes) + ((u64)(b + 1) << shift));
            const u64 lmax = (((u64)1 << shift) * sizeof(NodeT<WIDE>)) / 128;
            for (u64 l = l0 + threadIdx.x; l < l1 && l < lmax; l += INS_BLOCK)
                if (((u64)(b + 1) << shift) + l * (128 / sizeof(NodeT<WIDE>)) < t.n_local)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(basep + l * 128));
        }
#pragma unroll
        for (int r = 0; r < INS_ROUNDS; r++)
            if ((klo[r] | khi[r]) != 0)      // tuples never carry the k-mer-0 key, so 0 = past the end
                insert_one<WIDE, TRACK>(t, klo[r], khi[r], (u32)(meta[r] & 15), (u32)((meta[r] >> 4) & 15), meta[r] >> 8, n_new, n_conf);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { n_new += __shfl_xor_sync(0xffffffffu, n_new, s); n_conf += __shfl_xor_sync(0xffffffffu, n_conf, s); }
    if ((threadIdx.x & 31) == 0) {
        if (n_new) atomicAdd(t.counters + CNT_NEW, (u64)n_new);
        if (n_conf) atomicAdd(t.counters + CNT_CONFLICT, (u64)n_conf);
    }
}

// Owner side of the PULL exchange: every source rank q partitioned its occurrences by (owner, table slice) into its own send
// buffer (region of `capb` tuples per bucket, fills[q * fill_stride + bucket] of them used); the owner walks ITS buckets slice
// by slice and reads the n_src regions of a slice straight from the sources' memory over NVLink peer mappings -- long
// contiguous reads, no receive buffer, no owner-side partition pass.  Same tile scheduling, L2 prefetch of the next slice
// and insert as k_insert_tuples.
struct PullSrc {
    const u64 *const *ptrs;   // [n_src] base of every source's send buffer (peer mappings; the own one is local)
    const u32 *fills;         // all-gathered fill counters
    u32 n_src, fill_stride;
    u32 region0;              // first bucket of this owner in a source's buffer: rank * n_buckets
    u32 capb;                 // tuples per region (multiple of INS_TILE)
};

template <bool WIDE, bool TRACK>
__global__ void __launch_bounds__(INS_BLOCK, INS_CTAS) k_insert_tuples_pull(PullSrc ps, TableView t, u32 n_buckets, int shift, u64 *tile_counter)
{
    constexpr int TW = WIDE ? 4 : 2;
    u32 n_new = 0, n_conf = 0;
    const u32 tpr = ps.capb / INS_TILE;                       // tiles per region
    const u64 tiles_per_bucket = (u64)tpr * ps.n_src;
    const u64 n_tiles = tiles_per_bucket * n_buckets;
    __shared__ u64 s_tile[2];
    u64 next_tile = 0;
    if (threadIdx.x == 0) next_tile = __ldcg(t.counters + CNT_ERROR) ? (1ULL << 40) : atomicAdd(tile_counter, 1ULL);
    for (u32 par = 0;; par ^= 1) {
        if (threadIdx.x == 0) s_tile[par] = next_tile;
        __syncthreads();
        const u64 ti = s_tile[par];
        if (ti >= n_tiles) break;
        if (threadIdx.x == 0) next_tile = atomicAdd(tile_counter, 1ULL);
        const u32 b = (u32)(ti / tiles_per_bucket);
        const u32 within = (u32)(ti - (u64)b * tiles_per_bucket);
        const u32 q = within / tpr;
        const u32 off = (within - q * tpr) * INS_TILE;
        // this tile's share of the NEXT slice goes to L2 (by position inside the bucket's span, used or not)
        if (b + 1 < n_buckets && ((u64)(b + 1) << shift) < t.n_local) {
            const u64 slice_lo = (u64)(b + 1) << shift;
            u64 slice_n = (u64)1 << shift;
            if (slice_lo + slice_n > t.n_local) slice_n = t.n_local - slice_lo;
            const u64 lines = slice_n * sizeof(NodeT<WIDE>) / 128;
            const u64 l0 = lines * within / tiles_per_bucket, l1 = lines * (within + 1) / tiles_per_bucket;
            const char *basep = reinterpret_cast<const char *>(static_cast<const NodeT<WIDE> *>(t.nodes) + slice_lo);
            for (u64 l = l0 + threadIdx.x; l < l1; l += INS_BLOCK) asm volatile("prefetch.global.L2 [%0];" ::"l"(basep + l * 128));
        }
        const u32 used = __ldg(ps.fills + (size_t)q * ps.fill_stride + ps.region0 + b);
        if (off >= used) continue;                             // unused tail of the region (block-uniform)
        const u64 *src = ps.ptrs[q] + ((u64)(ps.region0 + b) * ps.capb + off) * TW;
        u64 klo[INS_ROUNDS], khi[INS_ROUNDS], meta[INS_ROUNDS];
#pragma unroll
        for (int r = 0; r < INS_ROUNDS; r++) {
            const u32 i = (u32)r * INS_BLOCK + threadIdx.x;
            klo[r] = 0; khi[r] = 0; meta[r] = 0;
            if (off + i < used) {
                if (WIDE) { u64 z; ld256_cs(reinterpret_cast<const ulonglong2 *>(src) + 2 * i, klo[r], khi[r], meta[r], z); }
                else { ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2 *>(src) + i); klo[r] = x.x; meta[r] = x.y; }
            }
        }
#pragma unroll
        for (int r = 0; r < INS_ROUNDS; r++)
            if ((klo[r] | khi[r]) != 0)
                insert_one<WIDE, TRACK>(t, klo[r], khi[r], (u32)(meta[r] & 15), (u32)((meta[r] >> 4) & 15), meta[r] >> 8, n_new, n_conf);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { n_new += __shfl_xor_sync(0xffffffffu, n_new, s); n_conf += __shfl_xor_sync(0xffffffffu, n_conf, s); }
    if ((threadIdx.x & 31) == 0) {
        if (n_new) atomicAdd(t.counters + CNT_NEW, (u64)n_new);
        if (n_conf) atomicAdd(t.counters + CNT_CONFLICT, (u64)n_conf);
    }
}

// ---------------------------------------------------------------------------------------------------
// reference slot layout: the reference places a key in the first free
// slot from hash%P at the moment of its FIRST occurrence.  That layout is the unique one in which every
// key k at slot s has only earlier-first-seen keys in [home(k), s); priority linear probing with
// atomicMin on the ordinal builds it in any execution order.  owner[] holds the ordinal per slot.
// ---------------------------------------------------------------------------------------------------
template <bool WIDE, bool TRACK>
__global__ void k_layout_insert(const NodeT<WIDE> *__restrict__ nodes, u64 n_local, u64 lo_slot, u64 *owner, u64 P, u64 M)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += stride) {
        NodeRegs n;
        if (WIDE) { u64 pad; ld256_cg(nodes + i, n.klo, n.khi, n.nord, pad); }
        else { ulonglong2 kk = __ldcg(reinterpret_cast<const ulonglong2 *>(nodes + i)); n.klo = kk.x; n.nord = kk.y; n.khi = 0; }
        u64 klo = n.klo, khi = n.khi;
        if ((klo | khi) == 0) continue;
        u64 cur = TRACK ? ~n.nord : (lo_slot + i);
        u64 h = WIDE ? hash_code_wide(klo, khi) : hash_code(klo);
        u64 s = mod_P(h, P, M);
        for (;;) {
            u64 old = atomicMin(owner + s, cur);
            if (old == EMPTY_PRI) break;
            if (old > cur) cur = old;      // we took the slot; carry the displaced key onwards
            s = (s + 1 == P) ? 0 : s + 1;
        }
    }
}

template <bool WIDE, bool TRACK>
__global__ void k_layout_place(const NodeT<WIDE> *__restrict__ nodes, u64 n_local, u64 lo_slot, const u64 *__restrict__ owner,
                               u64 P, u64 M, void *out, u32 *nul32)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += stride) {
        NodeRegs n;
        load_node(nodes + i, n);
        u64 klo = n.klo, khi = n.khi;
        if ((klo | khi) == 0) continue;
        u64 links = (u64)pack_link(n.c0) | ((u64)pack_link(n.c1) << 32);
        u64 pri = TRACK ? ~n.nord : (lo_slot + i);
        u64 h = WIDE ? hash_code_wide(klo, khi) : hash_code(klo);
        u64 s = mod_P(h, P, M);
        while (__ldg(owner + s) != pri) s = (s + 1 == P) ? 0 : s + 1;
        if (WIDE) {
            ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out) + 2 * s;
            dst[0] = make_ulonglong2(klo, khi);
            dst[1] = make_ulonglong2(links, 0ULL);
        } else {
            reinterpret_cast<ulonglong2 *>(out)[s] = make_ulonglong2(klo, links);
        }
        atomicOr(nul32 + (s >> 5), flag_mask(s));
    }
}

// ---- cluster-local layout (the fast path) ------------------------------------------------------------
// Linear probing never moves a key across an empty slot, and the set of occupied slots does not depend on
// insertion order.  So the reference layout can be rebuilt one CLUSTER (maximal run of occupied slots) at a
// time: replay the cluster's keys in first-occurrence order, each into the first free slot at or after its
// home.  One streaming pass over the table, no atomics, no scratch: the thread that sees a cluster start
// walks it (clusters average 2-3 slots at load 0.5), simulates the replay on a 64-bit occupancy mask and
// writes the 16-B image nodes; clusters longer than 64 slots and the region where the table wraps around
// go to k_layout_regions, which runs the atomicMin priority probing on a per-region scratch.
// Window of the table a layout pass works on, in VIRTUAL slot numbers v (index into the physical node array):
//   unsharded : v = global slot, window [0, P), gbase = 0
//   sharded   : the physical array of rank r is [porch | own home range | margin]; v = 0 is global slot gbase =
//               (lo_r - porch) mod P.  The porch holds the neighbour's tail cluster (cross-shard hand-off), so a
//               cluster that crosses the shard boundary is replayed as one unit by the rank on its right.
// Every key found inside the window has its home inside [gbase, gbase + window) (mod P), so home -> virtual is
// home - gbase (+ P when that is negative).
struct LayoutGeom {
    u64 P, M;
    u64 gbase;        // global slot of v = 0
    u64 v_begin;      // first slot laid out
    u64 v_end;        // one past the last slot laid out
    // A pass over a SUB-window of a table that exists beyond it (dbg_finish_export lays the table out group by group,
    // behind the inserts): slots in [v_first, v_halo) exist -- tiles read their halo there, and a cluster that starts
    // inside the window is replayed (and written) to its end even past v_end; slots >= v_final are not final yet, a long
    // cluster that reaches them raises `overflow` (the caller redoes the layout in one pass).  Whole windows:
    // v_first = v_begin, v_halo = v_final = v_end.
    u64 v_first, v_halo, v_final;
};

__device__ __forceinline__ u64 home_virtual(const LayoutGeom &g, u64 klo, u64 khi, bool wide)
{
    const u64 h = wide ? hash_code_wide(klo, khi) : hash_code(klo);
    const u64 home = mod_P(h, g.P, g.M);
    return home >= g.gbase ? home - g.gbase : home + (g.P - g.gbase);
}

struct LayoutInfo {
    u64 e;              // first slot of the cluster that runs over the end of the table (== P: none)
    u64 g;              // slots [0, g) take part in the wrap-around region
    u64 mt;             // occupied slots in the overflow margin (>= P)
    u64 scratch_used;   // entries of the region scratch handed out
    u32 n_regions;
    u32 overflow;       // region list / scratch exhausted: host falls back to the global method
};
struct LayoutRegion { u64 a, n, off; u64 wrap; };
constexpr u32 MAX_REGIONS = 1u << 20;

template <bool WIDE>
__device__ __forceinline__ bool slot_occupied(const NodeT<WIDE> *nodes, u64 i)
{
    if (WIDE) { ulonglong2 k = __ldcg(reinterpret_cast<const ulonglong2 *>(nodes + i)); return (k.x | k.y) != 0; }
    return __ldcg(&nodes[i].klo) != 0;
}

template <bool WIDE>
__device__ __forceinline__ void write_image(void *out, u64 slot, u64 klo, u64 khi, u64 links)
{
    if (WIDE) {
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out) + 2 * slot;
        dst[0] = make_ulonglong2(klo, khi);
        dst[1] = make_ulonglong2(links, 0ULL);
    } else {
        reinterpret_cast<ulonglong2 *>(out)[slot] = make_ulonglong2(klo, links);
    }
}

// one thread: where does the table wrap?  (tiny sequential scans around slot P-1 and slot 0)
template <bool WIDE>
__global__ void k_layout_wrapscan(const NodeT<WIDE> *__restrict__ nodes, u64 n_local, u64 P, LayoutInfo *info, LayoutRegion *regions,
                                  u64 scratch_cap)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    LayoutInfo li; li.e = P; li.g = 0; li.mt = 0; li.scratch_used = 0; li.n_regions = 0; li.overflow = 0;
    u64 mt = 0;
    while (P + mt < n_local && slot_occupied<WIDE>(nodes, P + mt)) mt++;
    if (mt > 0) {
        u64 e = P - 1;                       // slot P-1 is occupied (a probe chain ran through it)
        while (e > 0 && slot_occupied<WIDE>(nodes, e - 1)) e--;
        u64 need = mt, g = 0;
        while (need > 0 && g < e) { if (!slot_occupied<WIDE>(nodes, g)) need--; g++; }
        li.e = e; li.g = g; li.mt = mt;
        u64 R = (P - e) + g;
        if (need > 0 || R > scratch_cap) li.overflow = 1;
        else { regions[0].a = e; regions[0].n = R; regions[0].off = 0; regions[0].wrap = 1; li.n_regions = 1; li.scratch_used = R; }
    }
    *info = li;
}

#ifndef DBG_LT
#define DBG_LT 256
#endif
constexpr int LT = DBG_LT;  // slots per layout tile (one CTA round) = threads per CTA
constexpr int LH = 64;      // halo: the longest cluster handled in shared memory
constexpr int LW = (LT + LH) / 32;

// cluster [start, end) around occupied position k of the tile's occupancy words (positions 0 .. LT+LH-1);
// start = -1: the cluster began before the tile; end = LT+LH+1: it runs past the window
__device__ __forceinline__ void cluster_bounds(const u32 *W, int k, bool prev_occ, int &start, int &end)
{
    int w = k >> 5;
    u32 m = ~W[w] & ((1u << (k & 31)) - 1u);
    start = -2;
    for (;;) {
        if (m) { start = (w << 5) + 32 - __clz(m); break; }
        if (--w < 0) break;
        m = ~W[w];
    }
    if (start == -2) start = prev_occ ? -1 : 0;
    w = k >> 5;
    m = ~W[w] & ~((2u << (k & 31)) - 1u);            // zeros above k in its word
    end = LT + LH + 1;
    for (;;) {
        if (m) { end = (w << 5) + __ffs(m) - 1; break; }
        if (++w >= LW) break;
        m = ~W[w];
    }
}

// index (0 .. LT+LH-1) of the i-th occupied slot of the window: chunk from the per-32 counts, then the n-th set bit
__device__ __forceinline__ int nth_occupied(const u32 *W, const u32 *wcnt, u32 i)
{
    int c = 0;
    u32 before = 0;
    bool found = false;
#pragma unroll
    for (int q = 0; q < LW; q++) {
        const u32 n = wcnt[q];
        if (!found && before + n <= i) { before += n; c = q + 1; } else found = true;
    }
    if (c >= LW) c = LW - 1;                                  // (not reached for i < n_occ)
    // position of the r-th (0-based) set bit of W[c]: binary search on population counts
    u32 w = W[c], r = i - before, pos = 0, t;
    t = __popc(w & 0xFFFFu); if (r >= t) { pos += 16; r -= t; w >>= 16; }
    t = __popc(w & 0xFFu);   if (r >= t) { pos += 8;  r -= t; w >>= 8; }
    t = __popc(w & 0xFu);    if (r >= t) { pos += 4;  r -= t; w >>= 4; }
    t = __popc(w & 0x3u);    if (r >= t) { pos += 2;  r -= t; w >>= 2; }
    t = w & 1u;              if (r >= t) { pos += 1; }
    return (c << 5) + (int)(pos & 31u);
}

template <bool WIDE, bool TRACK>
__global__ void __launch_bounds__(LT, 2048 / LT) k_layout_clusters(const NodeT<WIDE> *__restrict__ nodes, LayoutGeom geo, void *out, u32 *nul32,
                                                        LayoutInfo *info, LayoutRegion *regions, u64 scratch_cap)
{
    // One tile = LT slots (+ LH halo).  (1) every thread loads one node (coalesced 256-bit loads), the raw nodes go to
    // shared memory, empty slots are written out at once; (2) the OCCUPIED slots are renumbered densely, so the
    // expensive part -- hash -> home slot, cluster bounds, the replay with priority probing (atomicMin on the ordinal,
    // the same algorithm as the global method, but on chip) -- runs on full warps instead of on the half-empty ones a
    // table at load 0.5 gives; (3) each key finds its slot and the 16-B image node is written.
    // nul32 == nullptr (sharded windows): the occupancy bitmap is produced afterwards from the image (k_nul_from_image),
    // because a window's 32-slot groups are not aligned with the global bitmap words.
    constexpr int NQ = WIDE ? 3 : 2;     // 16-byte quarters kept per node: {klo,nord|khi} [{nord,-}] {c0,c1}
    __shared__ ulonglong2 s_raw[NQ * (LT + LH)];
    __shared__ u64 s_ord[LT + LH], s_owner[LT + LH];
    __shared__ u32 s_home[LT + LH];     // home - tile start (>= 0 for every cluster that starts in the tile); ~0: not replayed here
    __shared__ u32 s_W[LW], s_wcnt[LW]; // occupancy bitmap of the window, occupied slots per 32
    __shared__ int s_prev;
    const u64 e_skip = info->e, g_skip = info->g;
    const u64 v_end = geo.v_end;
    const int t = threadIdx.x;
    const u32 lane = t & 31;
    for (u64 i0 = geo.v_begin + (u64)blockIdx.x * LT; i0 < v_end; i0 += (u64)gridDim.x * LT) {
        for (int k = t; k < LT + LH; k += LT) {          // k = t, and t + LT for the first LH threads (warp uniform)
            const u64 s = i0 + k;
            NodeRegs nd; nd.klo = 0; nd.khi = 0; nd.nord = 0; nd.c0 = 0; nd.c1 = 0;
            if (s < geo.v_halo) load_node(nodes + s, nd);
            const bool o = (nd.klo | nd.khi) != 0;
            if (WIDE) { s_raw[3 * k] = make_ulonglong2(nd.klo, nd.khi); s_raw[3 * k + 1] = make_ulonglong2(nd.nord, 0ULL); s_raw[3 * k + 2] = make_ulonglong2(nd.c0, nd.c1); }
            else { s_raw[2 * k] = make_ulonglong2(nd.klo, nd.nord); s_raw[2 * k + 1] = make_ulonglong2(nd.c0, nd.c1); }
            s_owner[k] = EMPTY_PRI;
            const u32 bal = __ballot_sync(0xffffffffu, o);
            if (lane == 0) {
                s_W[k >> 5] = bal; s_wcnt[k >> 5] = __popc(bal);
                if (nul32 && k < LT && s < v_end) nul32[s >> 5] = __byte_perm(__brev(bal), 0, 0x0123);   // MSB-first bitmap word of 32 slots
            }
            if (k < LT && s < v_end && !o) write_image<WIDE>(out, s, 0, 0, 0);
        }
        if (t == 0) s_prev = (i0 > geo.v_first) && slot_occupied<WIDE>(nodes, i0 - 1);
        __syncthreads();
        u32 n_occ = 0;
#pragma unroll
        for (int q = 0; q < LW; q++) n_occ += s_wcnt[q];
        for (u32 i = t; i < n_occ; i += LT) {
            const int k = nth_occupied(s_W, s_wcnt, i);
            u64 klo, khi = 0, nord;
            if (WIDE) { const ulonglong2 a = s_raw[3 * k]; klo = a.x; khi = a.y; nord = s_raw[3 * k + 1].x; }
            else { const ulonglong2 a = s_raw[2 * k]; klo = a.x; nord = a.y; }
            u64 cur = TRACK ? ~nord : i0 + k;
            u32 pos = (u32)(home_virtual(geo, klo, khi, WIDE) - i0);
            s_ord[k] = cur;
            s_home[k] = 0xffffffffu;
            int start, end;
            cluster_bounds(s_W, k, s_prev != 0, start, end);
            if (start < 0 || start >= LT) continue;                   // belongs to the previous / next tile
            const u64 cs = i0 + start;
            if (cs < g_skip || cs == e_skip) continue;                // wrap-around region: k_layout_regions
            if (end > LT + LH || end - start > LH) {
                // longer than the shared-memory window: measure it in global memory, hand it to k_layout_regions
                if (k == start) {
                    u64 len = (u64)(end > LT + LH ? LT + LH - start : end - start);
                    while (cs + len < geo.v_halo && slot_occupied<WIDE>(nodes, cs + len)) len++;
                    if (cs + len >= geo.v_final && geo.v_final < geo.v_halo) info->overflow = 1;      // ran into slots still being built
                    u32 rr = atomicAdd(&info->n_regions, 1u);
                    u64 off = atomicAdd(&info->scratch_used, len);
                    if (rr >= MAX_REGIONS || off + len > scratch_cap) info->overflow = 1;
                    else { regions[rr].a = cs; regions[rr].n = len; regions[rr].off = off; regions[rr].wrap = 0; }
                }
                continue;
            }
            if (end - start == 1) {
                // a key alone between two empty slots sits at its home slot (probing never crosses an empty slot): nothing
                // to replay -- about a third of the keys of a table at load 0.5 -- write its image node at once
                const ulonglong2 cc = s_raw[NQ * k + NQ - 1];
                write_image<WIDE>(out, i0 + k, klo, khi, (u64)pack_link(cc.x) | ((u64)pack_link(cc.y) << 32));
                continue;
            }
            s_home[k] = pos;
            for (;;) {
                u64 old = atomicMin(&s_owner[pos], cur);
                if (old == EMPTY_PRI) break;
                if (old > cur) cur = old;      // we took the slot; carry the displaced key onwards
                pos++;
            }
        }
        __syncthreads();
        for (u32 i = t; i < n_occ; i += LT) {
            const int k = nth_occupied(s_W, s_wcnt, i);
            u32 pos = s_home[k];
            if (pos == 0xffffffffu) continue;
            const u64 pri = s_ord[k];
            while (s_owner[pos] != pri) pos++;
            const ulonglong2 a = s_raw[NQ * k], cc = s_raw[NQ * k + NQ - 1];
            write_image<WIDE>(out, i0 + pos, a.x, WIDE ? a.y : 0ULL, (u64)pack_link(cc.x) | ((u64)pack_link(cc.y) << 32));
        }
        __syncthreads();
    }
}

// ---- version 2 of the cluster-local layout pass ------------------------------------------------------------------
// Same algorithm, different work assignment.  A tile is L2T slots + LH halo = L2S slots = L2W occupancy words, worked on by
// L2S/2 threads: every thread LOADS two slots, and in the replay / write-out phases every WARP owns two occupancy words
// (64 slots, about 32 of them occupied at load 0.5), lane l taking the l-th occupied slot of the pair -- full warps without
// the block-wide renumbering (which cost a 10-word prefix walk per key, twice), and the hash is computed only for keys that
// are replayed here (not for singleton clusters, not for halo keys that belong to the next tile).
#ifndef DBG_L2T
#define DBG_L2T 448
#endif
constexpr int L2T = DBG_L2T;            // slots per tile (multiple of 64)
constexpr int L2S = L2T + LH;           // slots staged
constexpr int L2W = L2S / 32;           // occupancy words
constexpr int L2N = L2S / 2;            // threads per CTA
static_assert(L2T % 64 == 0 && L2N % 32 == 0, "tile geometry");

__device__ __forceinline__ void cluster_bounds2(const u32 *W, int k, bool prev_occ, int &start, int &end)
{
    int w = k >> 5;
    u32 m = ~W[w] & ((1u << (k & 31)) - 1u);
    start = -2;
    for (;;) {
        if (m) { start = (w << 5) + 32 - __clz(m); break; }
        if (--w < 0) break;
        m = ~W[w];
    }
    if (start == -2) start = prev_occ ? -1 : 0;
    w = k >> 5;
    m = ~W[w] & ~((2u << (k & 31)) - 1u);
    end = L2S + 1;
    for (;;) {
        if (m) { end = (w << 5) + __ffs(m) - 1; break; }
        if (++w >= L2W) break;
        m = ~W[w];
    }
}

// position of the r-th (0-based) set bit of w
__device__ __forceinline__ u32 nth_set_bit(u32 w, u32 r)
{
    u32 pos = 0, t;
    t = __popc(w & 0xFFFFu); if (r >= t) { pos += 16; r -= t; w >>= 16; }
    t = __popc(w & 0xFFu);   if (r >= t) { pos += 8;  r -= t; w >>= 8; }
    t = __popc(w & 0xFu);    if (r >= t) { pos += 4;  r -= t; w >>= 4; }
    t = __popc(w & 0x3u);    if (r >= t) { pos += 2;  r -= t; w >>= 2; }
    t = w & 1u;              if (r >= t) { pos += 1; }
    return pos & 31u;
}

template <bool WIDE, bool TRACK>
__global__ void __launch_bounds__(L2N, 2048 / L2N) k_layout_clusters2(const NodeT<WIDE> *__restrict__ nodes, LayoutGeom geo, void *out, u32 *nul32,
                                                                      LayoutInfo *info, LayoutRegion *regions, u64 scratch_cap)
{
    constexpr int NQ = WIDE ? 3 : 2;
    __shared__ ulonglong2 s_raw[NQ * L2S];
    __shared__ u64 s_owner[L2S];
    __shared__ u32 s_W[L2W];
    __shared__ int s_prev;
    const u64 e_skip = info->e, g_skip = info->g;
    const u64 v_end = geo.v_end;
    const int t = threadIdx.x;
    const u32 lane = t & 31, warp = t >> 5;
    for (u64 i0 = geo.v_begin + (u64)blockIdx.x * L2T; i0 < v_end; i0 += (u64)gridDim.x * L2T) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int k = t + j * L2N;
            const u64 s = i0 + k;
            NodeRegs nd; nd.klo = 0; nd.khi = 0; nd.nord = 0; nd.c0 = 0; nd.c1 = 0;
            if (s < geo.v_halo) load_node(nodes + s, nd);
            const bool o = (nd.klo | nd.khi) != 0;
            if (WIDE) { s_raw[3 * k] = make_ulonglong2(nd.klo, nd.khi); s_raw[3 * k + 1] = make_ulonglong2(nd.nord, 0ULL); s_raw[3 * k + 2] = make_ulonglong2(nd.c0, nd.c1); }
            else { s_raw[2 * k] = make_ulonglong2(nd.klo, nd.nord); s_raw[2 * k + 1] = make_ulonglong2(nd.c0, nd.c1); }
            s_owner[k] = EMPTY_PRI;
            const u32 bal = __ballot_sync(0xffffffffu, o);
            if (lane == 0) {
                s_W[k >> 5] = bal;
                if (nul32 && k < L2T && s < v_end) nul32[s >> 5] = __byte_perm(__brev(bal), 0, 0x0123);
            }
            if (k < L2T && s < v_end && !o) write_image<WIDE>(out, s, 0, 0, 0);
        }
        if (t == 0) s_prev = (i0 > geo.v_first) && slot_occupied<WIDE>(nodes, i0 - 1);
        __syncthreads();
        // ---- replay: this warp's two words ----
        const u32 w0 = s_W[2 * warp], w1 = s_W[2 * warp + 1];
        const u32 c0 = __popc(w0), n_mine = c0 + __popc(w1);
        u32 mine[2];                    // per pass: (home - i0) << 16 | slot index in the tile; 0xffffffff: nothing left for phase 3
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const u32 idx = lane + 32 * pass;
            mine[pass] = 0xffffffffu;
            if (idx >= n_mine) continue;
            const int k = idx < c0 ? (int)(64 * warp + nth_set_bit(w0, idx)) : (int)(64 * warp + 32 + nth_set_bit(w1, idx - c0));
            int start, end;
            cluster_bounds2(s_W, k, s_prev != 0, start, end);
            if (start < 0 || start >= L2T) continue;                  // belongs to the previous / next tile
            const u64 cs = i0 + start;
            if (cs < g_skip || cs == e_skip) continue;                // wrap-around region: k_layout_regions
            if (end > L2S || end - start > LH) {
                if (k == start) {
                    u64 len = (u64)(end > L2S ? L2S - start : end - start);
                    while (cs + len < geo.v_halo && slot_occupied<WIDE>(nodes, cs + len)) len++;
                    if (cs + len >= geo.v_final && geo.v_final < geo.v_halo) info->overflow = 1;
                    u32 rr = atomicAdd(&info->n_regions, 1u);
                    u64 off = atomicAdd(&info->scratch_used, len);
                    if (rr >= MAX_REGIONS || off + len > scratch_cap) info->overflow = 1;
                    else { regions[rr].a = cs; regions[rr].n = len; regions[rr].off = off; regions[rr].wrap = 0; }
                }
                continue;
            }
            u64 klo, khi = 0, nord;
            if (WIDE) { const ulonglong2 a = s_raw[3 * k]; klo = a.x; khi = a.y; nord = s_raw[3 * k + 1].x; }
            else { const ulonglong2 a = s_raw[2 * k]; klo = a.x; nord = a.y; }
            if (end - start == 1) {
                // alone between two empty slots: it sits at its home slot, nothing to replay (no hash needed either)
                const ulonglong2 cc = s_raw[NQ * k + NQ - 1];
                write_image<WIDE>(out, i0 + k, klo, khi, (u64)pack_link(cc.x) | ((u64)pack_link(cc.y) << 32));
                continue;
            }
            u64 cur = TRACK ? ~nord : i0 + k;
            u32 pos = (u32)(home_virtual(geo, klo, khi, WIDE) - i0);
            mine[pass] = (pos << 16) | (u32)k;
            for (;;) {
                u64 old = atomicMin(&s_owner[pos], cur);
                if (old == EMPTY_PRI) break;
                if (old > cur) cur = old;      // we took the slot; carry the displaced key onwards
                pos++;
            }
        }
        __syncthreads();
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            if (mine[pass] == 0xffffffffu) continue;
            u32 pos = mine[pass] >> 16;
            const int k = (int)(mine[pass] & 0xffffu);
            const u64 pri = TRACK ? ~(WIDE ? s_raw[3 * k + 1].x : s_raw[2 * k].y) : i0 + k;
            while (s_owner[pos] != pri) pos++;
            const ulonglong2 a = s_raw[NQ * k], cc = s_raw[NQ * k + NQ - 1];
            write_image<WIDE>(out, i0 + pos, a.x, WIDE ? a.y : 0ULL, (u64)pack_link(cc.x) | ((u64)pack_link(cc.y) << 32));
        }
        __syncthreads();
    }
}

// long clusters and the wrap-around region: priority probing (atomicMin on the ordinal) inside a private scratch
template <bool WIDE, bool TRACK>
__global__ void __launch_bounds__(256) k_layout_regions(const NodeT<WIDE> *__restrict__ nodes, LayoutGeom geo, void *out, u32 *nul32,
                                                        const LayoutInfo *__restrict__ info, const LayoutRegion *__restrict__ regions, u64 *scratch)
{
    if (info->overflow) return;
    const u64 P = geo.P;
    const u32 n_regions = info->n_regions < MAX_REGIONS ? info->n_regions : MAX_REGIONS;
    const u64 e = info->e, mt = info->mt, g = info->g;
    for (u32 r = blockIdx.x; r < n_regions; r += gridDim.x) {
        const LayoutRegion rg = regions[r];
        u64 *sc = scratch + rg.off;
        const u64 tail = rg.wrap ? (P - e) + mt : 0;          // wrap: sources are A-slots [e, P+mt) then [0, g)
        const u64 nsrc = rg.wrap ? tail + g : rg.n;
        for (u64 k = threadIdx.x; k < rg.n; k += blockDim.x) sc[k] = EMPTY_PRI;
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {
            for (u64 k = threadIdx.x; k < nsrc; k += blockDim.x) {
                const u64 s = rg.wrap ? (k < tail ? e + k : k - tail) : rg.a + k;
                NodeRegs nd;
                load_node(nodes + s, nd);
                if ((nd.klo | nd.khi) == 0) continue;
                u64 home = home_virtual(geo, nd.klo, nd.khi, WIDE);      // (wrap regions exist only in unsharded windows: virtual == global)
                u64 pos = rg.wrap ? (home >= e ? home - e : home + (P - e)) : home - rg.a;
                const u64 pri = TRACK ? ~nd.nord : s;
                if (pass == 0) {
                    u64 cur = pri;
                    for (;;) {
                        u64 old = atomicMin(sc + pos, cur);
                        if (old == EMPTY_PRI) break;
                        if (old > cur) cur = old;      // we took the slot; carry the displaced key onwards
                        pos++;
                    }
                } else {
                    while (__ldcg(sc + pos) != pri) pos++;
                    u64 slot = rg.wrap ? (e + pos >= P ? e + pos - P : e + pos) : rg.a + pos;
                    u64 links = (u64)pack_link(nd.c0) | ((u64)pack_link(nd.c1) << 32);
                    write_image<WIDE>(out, slot, nd.klo, nd.khi, links);
                    if (nul32) atomicOr(nul32 + (slot >> 5), flag_mask(slot));
                }
            }
            __syncthreads();
        }
    }
}

// ---- cross-shard hand-off (sharded contexts) ---------------------------------------------------------------
// A probe cluster that crosses the boundary between two shards must be replayed as ONE unit.  The rank on the left
// hands its TAIL UNIT to the rank on its right: the run of occupied slots that ends at its last home slot (`a` nodes,
// kept in place by the receiver, in its porch) plus whatever its inserts pushed into the overflow margin (`mt` nodes:
// in the global table they sit in the right neighbour's first free slots, so the receiver adopts them by probing from
// its slot 0).  The receiver then lays out [porch tail | own range minus its own tail unit].
struct TailInfo { u64 a, mt, landing_max, adopted; };

template <bool WIDE>
__global__ void k_shard_tail_scan(const NodeT<WIDE> *__restrict__ nodes /* local slot 0 */, u64 own_size, u64 n_local, u64 max_a, TailInfo *ti)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u64 mt = 0, a = 0;
    while (own_size + mt < n_local && slot_occupied<WIDE>(nodes, own_size + mt)) mt++;
    while (a < own_size && a <= max_a && slot_occupied<WIDE>(nodes, own_size - 1 - a)) a++;
    ti->a = a; ti->mt = mt;
}

// adopt the left neighbour's margin nodes: first free slot from local slot 0, whole node (key, ordinal, counts) copied.
// The table is quiescent (no inserts in flight), one thread, plain stores.
template <bool WIDE>
__global__ void k_shard_adopt(NodeT<WIDE> *nodes /* local slot 0 */, u64 n_local, const NodeT<WIDE> *__restrict__ src, u64 n, TailInfo *ti)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u64 pos = 0, last = 0;
    bool any = false;
    for (u64 i = 0; i < n; i++) {
        while (pos < n_local && slot_occupied<WIDE>(nodes, pos)) pos++;
        if (pos >= n_local) { ti->landing_max = ~0ULL; ti->adopted = i; return; }
        nodes[pos] = src[i];
        last = pos; any = true;
    }
    // the chain continues through the occupied run behind the last adopted node: report where it ends
    if (any) { u64 q = last; while (q + 1 < n_local && slot_occupied<WIDE>(nodes, q + 1)) q++; ti->landing_max = q; }
    else ti->landing_max = 0;
    ti->adopted = n;
}

// occupancy bitmap of a range of GLOBAL slots from the laid-out image (sharded windows): word w covers global slots
// 32 w .. 32 w + 31, MSB-first bytes like nul_flag (kmerSet.h:144-155).  img is indexed by virtual slot; a global slot
// outside [g_first, g_first + n) (mod P) contributes 0.
template <bool WIDE>
__global__ void __launch_bounds__(256) k_nul_from_image(const void *__restrict__ img, u64 P, u64 gbase, u64 g_first, u64 n, u64 w_first,
                                                        u64 n_words, u32 *nul32)
{
    const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 w = gid >> 5;
    if (w >= n_words) return;
    const u64 gslot = (w_first + w) * 32 + (threadIdx.x & 31);
    bool o = false;
    if (gslot < P) {
        const u64 rel = gslot >= g_first ? gslot - g_first : gslot + (P - g_first);     // distance from the range start (mod P)
        if (rel < n) {
            const u64 v = gslot >= gbase ? gslot - gbase : gslot + (P - gbase);
            if (WIDE) { ulonglong2 k = __ldg(reinterpret_cast<const ulonglong2 *>(img) + 2 * v); o = (k.x | k.y) != 0; }
            else o = __ldg(reinterpret_cast<const u64 *>(img) + 2 * v) != 0;
        }
    }
    const u32 bal = __ballot_sync(0xffffffffu, o);
    if ((threadIdx.x & 31) == 0) nul32[w] = __byte_perm(__brev(bal), 0, 0x0123);
}

// add_node_to_kmerset(kset, PolyA) (kmerSet.cpp:253-273, DBGgraph.cpp:418): last, always, first null slot
template <bool WIDE>
__global__ void k_polyA_insert(u64 P, u64 M, const u64 *__restrict__ polyA, void *out, u32 *nul32, u64 *links_out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 l = 0, r = 0;
    for (int b = 0; b < 4; b++) {
        u64 cl = polyA[b], cr = polyA[4 + b];
        l |= (u32)(cl > 255 ? 255 : cl) << (24 - 8 * b);
        r |= (u32)(cr > 255 ? 255 : cr) << (24 - 8 * b);
    }
    u64 links = (u64)l | ((u64)r << 32);
    *links_out = links;
    u64 s = mod_P(WIDE ? hash_code_wide(0, 0) : hash_code(0), P, M);
    while (nul32[s >> 5] & flag_mask(s)) s = (s + 1 == P) ? 0 : s + 1;
    write_image<WIDE>(out, s, 0, 0, links);
    nul32[s >> 5] |= flag_mask(s);
}

// ---------------------------------------------------------------------------------------------------
// calculate_kmer_links (contig.cpp:107-205) + ordered compaction
// ---------------------------------------------------------------------------------------------------
constexpr int TILE = 2048;   // slots per compaction tile (one CTA iteration: 256 threads x 8)

__device__ __forceinline__ u32 load_links_of_slot(const void *img, u64 i, bool wide, u64 &klo, u64 &khi, u64 &links)
{
    if (wide) {
        ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(img) + 2 * i);
        ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(img) + 2 * i + 1);
        klo = a.x; khi = a.y; links = b.x;
    } else {
        ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(img) + i);
        klo = a.x; khi = 0; links = a.y;
    }
    return 0;
}

// classify one link word: number of lanes above the cutoff (capped at 3) and the arg-max base
// (first maximum wins: strict `<`, contig.cpp:137-141)
__device__ __forceinline__ void classify(u32 link, int cutoff, u32 &num, u32 &base, u32 *hist_nonzero, u32 &zeros)
{
    num = 0; base = 0; int maxd = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int d = (link >> (24 - 8 * j)) & 0xFF;
        if (d == 0) zeros++; else atomicAdd(hist_nonzero + d, 1u);
        if (d > cutoff) {
            if (num < 3) num++;
            if (maxd < d) { maxd = d; base = j; }
        }
    }
}

// pass 1: klink, del_flag, depth histogram, per-tile counts of tips / branches / survivors / filled
static __global__ void __launch_bounds__(256) k_links_classify(const void *__restrict__ img, const u32 *__restrict__ nul32, u64 P, int wide,
                                                        int cutoff, unsigned short *klink, u32 *del32,
                                                        u64 *depth_hist, u64 *stats3, u32 *tile_counts /* 4 per tile */)
{
    __shared__ u32 hist[256];
    __shared__ u32 cnt[4];
    const int tid = threadIdx.x;
    hist[tid] = 0;
    u32 zeros = 0; u64 total = 0, deleted = 0, linear = 0;
    const u64 n_tiles = (P + TILE - 1) / TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid < 4) cnt[tid] = 0;
        __syncthreads();
        u32 c_tip = 0, c_br = 0, c_surv = 0, c_fill = 0;
#pragma unroll
        for (int e = 0; e < TILE / 256; e++) {
            u64 i = tile * TILE + (u64)e * 256 + tid;
            bool in = i < P;
            bool filled = in && (nul32[i >> 5] & flag_mask(i));
            bool del = false;
            unsigned short kl = 0;
            if (filled) {
                u64 klo, khi, links;
                load_links_of_slot(img, i, wide, klo, khi, links);
                u32 ln, lb, rn, rb;
                classify((u32)links, cutoff, ln, lb, hist, zeros);
                classify((u32)(links >> 32), cutoff, rn, rb, hist, zeros);
                u32 lin = (ln == 1 && rn == 1);
                kl = (unsigned short)(ln | (lb << 2) | (rn << 4) | (rb << 6) | (lin << 8));
                total++; linear += lin; c_fill++;
                if (ln == 0 && rn == 0) { del = true; deleted++; } else c_surv++;
                if (ln + rn == 1) c_tip++;
                if (ln > 1 || rn > 1) c_br++;
            }
            if (in) klink[i] = kl;
            // del_flag: 32 consecutive slots -> one u32 of the MSB-first byte bitmap
            u32 bal = __ballot_sync(0xffffffffu, del);
            if ((tid & 31) == 0 && in) del32[i >> 5] = __byte_perm(__brev(bal), 0, 0x0123);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            c_tip += __shfl_xor_sync(0xffffffffu, c_tip, s); c_br += __shfl_xor_sync(0xffffffffu, c_br, s);
            c_surv += __shfl_xor_sync(0xffffffffu, c_surv, s); c_fill += __shfl_xor_sync(0xffffffffu, c_fill, s);
        }
        if ((tid & 31) == 0) { atomicAdd(&cnt[0], c_tip); atomicAdd(&cnt[1], c_br); atomicAdd(&cnt[2], c_surv); atomicAdd(&cnt[3], c_fill); }
        __syncthreads();
        if (tid < 4) tile_counts[tile * 4 + tid] = cnt[tid];
        __syncthreads();
    }
    // flush
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        zeros += __shfl_xor_sync(0xffffffffu, zeros, s);
        total += __shfl_xor_sync(0xffffffffu, total, s); deleted += __shfl_xor_sync(0xffffffffu, deleted, s);
        linear += __shfl_xor_sync(0xffffffffu, linear, s);
    }
    if ((tid & 31) == 0) {
        if (zeros) atomicAdd(depth_hist, (u64)zeros);
        if (total) atomicAdd(stats3, total);
        if (deleted) atomicAdd(stats3 + 1, deleted);
        if (linear) atomicAdd(stats3 + 2, linear);
    }
    __syncthreads();
    if (tid > 0 && hist[tid]) atomicAdd(depth_hist + tid, (u64)hist[tid]);
}

// exclusive scan of tile_counts (4 interleaved streams) by one CTA; totals[4] out
static __global__ void __launch_bounds__(1024) k_scan_tiles(const u32 *__restrict__ tile_counts, u64 n_tiles, u64 *tile_offs, u64 *totals)
{
    __shared__ u64 wsum[32][4];
    __shared__ u64 carry[4];
    const int tid = threadIdx.x;
    if (tid < 4) carry[tid] = 0;
    __syncthreads();
    for (u64 base = 0; base < n_tiles; base += 1024) {
        u64 t = base + tid;
        u64 v[4], inc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (t < n_tiles) ? tile_counts[t * 4 + k] : 0; inc[k] = v[k]; }
#pragma unroll
        for (int s = 1; s < 32; s <<= 1)
#pragma unroll
            for (int k = 0; k < 4; k++) { u64 x = __shfl_up_sync(0xffffffffu, inc[k], s); if ((tid & 31) >= s) inc[k] += x; }
        if ((tid & 31) == 31)
#pragma unroll
            for (int k = 0; k < 4; k++) wsum[tid >> 5][k] = inc[k];
        __syncthreads();
        u64 wb[4] = {0, 0, 0, 0};
        for (int w = 0; w < (tid >> 5); w++)
#pragma unroll
            for (int k = 0; k < 4; k++) wb[k] += wsum[w][k];
        if (t < n_tiles)
#pragma unroll
            for (int k = 0; k < 4; k++) tile_offs[t * 4 + k] = carry[k] + wb[k] + inc[k] - v[k];
        __syncthreads();
        if (tid == 1023)
#pragma unroll
            for (int k = 0; k < 4; k++) carry[k] += wb[k] + inc[k];
        __syncthreads();
    }
    if (tid < 4) totals[tid] = carry[tid];
}

// block-wide exclusive scan of a 0/1 flag over TILE slots laid out e-major (slot = e*256 + tid):
// returns this thread's rank for element e in ranks[e].  wcnt: TILE/32 + 1 words of shared memory.
__device__ __forceinline__ void tile_rank(const bool (&flag)[TILE / 256], u32 (&ranks)[TILE / 256], u32 *wcnt)
{
    const int tid = threadIdx.x;
#pragma unroll
    for (int e = 0; e < TILE / 256; e++) {
        u32 bal = __ballot_sync(0xffffffffu, flag[e]);
        ranks[e] = __popc(bal & ((1u << (tid & 31)) - 1));
        if ((tid & 31) == 0) wcnt[e * 8 + (tid >> 5)] = __popc(bal);   // slot order: e major, warp minor
    }
    __syncthreads();
    if (tid < 64) {   // exclusive scan of the 64 warp counts by two warps
        u32 v = wcnt[tid], inc = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { u32 x = __shfl_up_sync(0xffffffffu, inc, s); if ((tid & 31) >= s) inc += x; }
        if (tid == 31) wcnt[64] = inc;
        __syncwarp();
        wcnt[tid] = inc - v;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < TILE / 256; e++) {
        int me = e * 8 + (tid >> 5);
        ranks[e] += wcnt[me] + (me >= 32 ? wcnt[64] : 0u);
    }
    __syncthreads();
}

// pass 2: index-ordered tip and branch lists (contig.cpp:175-180 push_back order = slot order)
static __global__ void __launch_bounds__(256) k_links_lists(const unsigned short *__restrict__ klink, const u32 *__restrict__ nul32, u64 P,
                                                     const u64 *__restrict__ tile_offs, u64 *tips, u64 cap_tips,
                                                     u64 *branches, u64 cap_br)
{
    __shared__ u32 wcnt[TILE / 32 + 1];
    const int tid = threadIdx.x;
    const u64 n_tiles = (P + TILE - 1) / TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        bool is_tip[TILE / 256], is_br[TILE / 256];
#pragma unroll
        for (int e = 0; e < TILE / 256; e++) {
            u64 i = tile * TILE + (u64)e * 256 + tid;
            bool filled = i < P && (nul32[i >> 5] & flag_mask(i));
            u32 kl = filled ? klink[i] : 0;
            u32 ln = kl & 3, rn = (kl >> 4) & 3;
            is_tip[e] = filled && (ln + rn == 1);
            is_br[e] = filled && (ln > 1 || rn > 1);
        }
        u32 rk[TILE / 256];
        tile_rank(is_tip, rk, wcnt);
#pragma unroll
        for (int e = 0; e < TILE / 256; e++)
            if (is_tip[e]) { u64 pos = tile_offs[tile * 4 + 0] + rk[e]; if (pos < cap_tips) tips[pos] = tile * TILE + (u64)e * 256 + tid; }
        tile_rank(is_br, rk, wcnt);
#pragma unroll
        for (int e = 0; e < TILE / 256; e++)
            if (is_br[e]) { u64 pos = tile_offs[tile * 4 + 1] + rk[e]; if (pos < cap_br) branches[pos] = tile * TILE + (u64)e * 256 + tid; }
    }
}

// pass 2': slot-ordered dump of the surviving (or all filled) nodes
static __global__ void __launch_bounds__(256) k_compact_nodes(const void *__restrict__ img, const u32 *__restrict__ nul32,
                                                       const u32 *__restrict__ del32, u64 P, int wide, int which /*2 surv, 3 filled*/,
                                                       const u64 *__restrict__ tile_offs, u64 cap, u64 *slots, u64 *klo_out,
                                                       u64 *khi_out, u32 *l_out, u32 *r_out)
{
    __shared__ u32 wcnt[TILE / 32 + 1];
    const int tid = threadIdx.x;
    const u64 n_tiles = (P + TILE - 1) / TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        bool keep[TILE / 256];
#pragma unroll
        for (int e = 0; e < TILE / 256; e++) {
            u64 i = tile * TILE + (u64)e * 256 + tid;
            bool filled = i < P && (nul32[i >> 5] & flag_mask(i));
            bool del = filled && which == 2 && (del32[i >> 5] & flag_mask(i));
            keep[e] = filled && !del;
        }
        u32 rk[TILE / 256];
        tile_rank(keep, rk, wcnt);
#pragma unroll
        for (int e = 0; e < TILE / 256; e++) {
            if (!keep[e]) continue;
            u64 i = tile * TILE + (u64)e * 256 + tid;
            u64 pos = tile_offs[tile * 4 + which] + rk[e];
            if (pos >= cap) continue;
            u64 klo, khi, links;
            load_links_of_slot(img, i, wide, klo, khi, links);
            if (slots) slots[pos] = i;
            if (klo_out) klo_out[pos] = klo;
            if (khi_out) khi_out[pos] = khi;
            if (l_out) l_out[pos] = (u32)links;
            if (r_out) r_out[pos] = (u32)(links >> 32);
        }
    }
}

// unordered dump of a shard's build table (multi-GPU: every rank hands its nodes to whoever merges them)
template <bool WIDE>
__global__ void __launch_bounds__(256) k_dump_shard(const NodeT<WIDE> *__restrict__ nodes, u64 n_local, u64 cap, u64 *cursor,
                                                    u64 *klo_out, u64 *khi_out, u32 *l_out, u32 *r_out, u64 *ord_out)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < n_local; i0 += stride) {
        u64 i = i0 + threadIdx.x;
        bool keep = false;
        NodeRegs n; n.klo = 0; n.khi = 0; n.nord = 0; n.c0 = 0; n.c1 = 0;
        if (i < n_local) { load_node(nodes + i, n); keep = (n.klo | n.khi) != 0; }
        const u64 klo = n.klo, khi = n.khi, nord = n.nord;
        u32 bal = __ballot_sync(0xffffffffu, keep);
        u64 base = 0;
        if ((threadIdx.x & 31) == 0 && bal) base = atomicAdd(cursor, (u64)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            u64 pos = base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1));
            if (pos < cap) {
                if (klo_out) klo_out[pos] = klo;
                if (khi_out) khi_out[pos] = WIDE ? khi : 0;
                if (l_out) l_out[pos] = pack_link(n.c0);
                if (r_out) r_out[pos] = pack_link(n.c1);
                if (ord_out) ord_out[pos] = ~nord;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// roofline denominator: uniformly random 32-B sector read-modify-writes
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 splitmix(u64 z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct __align__(32) Rec32 { u64 a, b, links, d; };

// MODE 0: 32-B load + store, 1: 32-B load + 64-bit CAS, 2: u32 RED, 3: f16x8 vector RED, 4: 32-B load only,
// 5: 32-B load then f16x8 RED on the same sector (the insert kernel's hit path)
template <int MODE>
__global__ void __launch_bounds__(256) k_random_rmw(Rec32 *tab, u64 n_nodes, u64 n_ops, u64 seed, u64 *sink)
{
    const u64 stride = (u64)gridDim.x * blockDim.x * G;
    u64 acc = 0;
    for (u64 base = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * G; base < n_ops; base += stride) {
        Rec32 *p[G]; u64 links[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
            u64 h = splitmix(seed + base + g);
            p[g] = tab + __umul64hi(h, n_nodes);
            links[g] = 0;
            if (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 5) {
                u64 a, b, c, d;
                ld256_cg(p[g], a, b, c, d);
                links[g] = c + (a & 1) + (b & 1) + (d & 1);
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (base + g < n_ops) {
                if (MODE == 0) p[g]->links = links[g] + 1;
                else if (MODE == 1) atomicCAS(&p[g]->links, links[g], links[g] + 1);
                else if (MODE == 2) atomicAdd(reinterpret_cast<u32 *>(&p[g]->links), 1u);
                else if (MODE == 3 || MODE == 5) {
                    u32 one = HALF_ONE + (u32)(links[g] >> 63);
                    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(&p[g]->links), "r"(one), "r"(0u), "r"(0u), "r"(0u) : "memory");
                } else acc += links[g];
            }
        }
    }
    if (MODE == 4 && acc == 0x123456789ULL) *sink = acc;
}

}  // namespace dbg
