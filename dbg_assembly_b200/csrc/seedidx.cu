// seedidx.cu -- the contig SEED INDEX of link_scaffold on the B200.
//
// What it replaces in the reference (paths relative to fanagislab/DBG_assembly, link_scaffold/):
//   init_kmerset                      kmerSet.cpp:82-107     table of find_next_prime(3 x contig length), load factor 0.5
//   chop_contig_to_kmerset            map_func.cpp:119-172   every k-mer of every N-free block of every contig, canonical,
//     + add_kmerset                   kmerSet.cpp:168-210    first occurrence keeps {id, pos, direct}; a second one clears `freq`
//   get_align_seed                    map_func.cpp:181-237   first read position whose k-mer and the k-mer SeedKmerNum further
//     + exist_kmerset                 kmerSet.cpp:216-238    are both unique in the contigs, on the same contig, SeedKmerNum apart
// Node (kmerSet.h:53-60, GCC bit-field layout): {u64 kmer; u64 id:32, pos:30, freq:1, direct:1}.
//
// It is the graph build's machinery with another payload (dbg_params::payload_mode = 1): contigs are cut at runs of 'N'
// (scaffold_to_contig, map_func.cpp:303-326) and into pieces of <= 32768 k-mers on the host, the pieces go through the same
// fused extraction + insert kernels as reads (an occurrence carries its strand instead of neighbour bases, the count lives
// in one half-float lane, the first-occurrence ordinal decides the node's {id, pos, direct}), and the same layout kernels
// rebuild the reference's SLOT ORDER (first free slot from hash % size in first-occurrence order), so the exported table is
// the reference's byte for byte.  k_seed_payload then turns (first ordinal, count) into the node's value word, and
// k_seed_kmer0 places the all-A k-mer (the graph build keeps it in a side node; here it is an ordinary key) by priority
// insertion.  Lookups: the extraction kernel with a sink that probes the finished table (k_build<SeedLookupSink>) leaves one
// hit word per read position; k_seed_pick takes, per read, the first position where get_align_seed would stop.
//
// Not reproduced: enlarge_kmerset (kmerSet.cpp:111-164).  map_pair / map_reads size the table at 3 x the contig length
// with load factor 0.5 (map_pair.cpp:122-124), so it never runs there; a table that would grow is refused
// (DBG_ERR_TABLE_FULL).  Blocks shorter than K: the reference's loop bound underflows (undefined behaviour, it crashes);
// they are skipped here.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dbg_b200.h"
#include "dbg_kernels.cuh"

using namespace dbg;

static thread_local char s_err[512] = "";
static int sset_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(s_err, sizeof(s_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" const char *seedidx_last_error(void) { return s_err; }

#define SCU(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return sset_err(e_ == cudaErrorMemoryAllocation ? DBG_ERR_NOMEM : DBG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define SDBG(call)                                                                                      \
    do {                                                                                                \
        int rc_ = (call);                                                                               \
        if (rc_ != DBG_OK) return sset_err(rc_, "%s: %s", #call, dbg_last_error());                    \
    } while (0)

static const uint32_t PIECE_KMERS = 32768;        // k-mers per piece: position inside a piece fits the ordinal's 16 bits
static const uint64_t SEED_MISS = ~0ULL;

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// value word of a node from its first-occurrence ordinal ((piece << 16 | j) << 1 | direct) and "seen once"
__device__ __forceinline__ u64 seed_value(u64 ord2, bool unique, const u32 *__restrict__ piece_id, const u32 *__restrict__ piece_start)
{
    const u64 piece = ord2 >> 17;
    const u32 j = (u32)(ord2 >> 1) & 0xFFFFu;
    const u64 id = __ldg(piece_id + piece);
    const u64 pos = (u64)((__ldg(piece_start + piece) + j) & 0x3FFFFFFFu);
    return id | (pos << 32) | ((u64)(unique ? 1 : 0) << 62) | ((ord2 & 1) << 63);
}

// every node of the BUILD table finds its slot in the laid-out image (probing from its home like a lookup would) and
// stores its value word there
static __global__ void __launch_bounds__(256) k_seed_payload(const NodeT<false> *__restrict__ nodes, u64 n_local, u64 P, u64 M,
                                                             ulonglong2 *img, const u32 *__restrict__ piece_id,
                                                             const u32 *__restrict__ piece_start, u64 *errors)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += stride) {
        NodeRegs n;
        load_node(nodes + i, n);
        if (n.klo == 0) continue;
        const u64 v = seed_value(~n.nord, (u32)(n.c0 & 0xFFFFu) == HALF_ONE, piece_id, piece_start);
        u64 s = mod_P(hash_code(n.klo), P, M), steps = 0;
        while (__ldcg(reinterpret_cast<const u64 *>(img + s)) != n.klo) {
            s = (s + 1 == P) ? 0 : s + 1;
            if (++steps > P) { atomicAdd(errors, 1ULL); break; }
        }
        if (steps <= P) img[s].y = v;
    }
}

// k-mer 0, inserted by priority: a key sits in the first slot from its home that no EARLIER-first-seen key holds, so the
// new key walks from its home, takes the slot of the first later-first-seen occupant and carries that one on, until an
// empty slot ends the chain (the same displacement rule as the layout kernels; first-occurrence order == (id, pos) order)
static __global__ void k_seed_kmer0(ulonglong2 *img, u32 *nul32, u64 P, u64 M, const u64 *__restrict__ polyA,
                                    const u32 *__restrict__ piece_id, const u32 *__restrict__ piece_start, u64 *placed)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const u64 cnt = polyA[0], nord = polyA[7];
    *placed = 0;
    if (cnt == 0 || nord == 0) return;
    ulonglong2 cur = make_ulonglong2(0ULL, seed_value(~nord, cnt == 1, piece_id, piece_start));
    auto prio = [](u64 v) { return ((v & 0xFFFFFFFFULL) << 30) | ((v >> 32) & 0x3FFFFFFFULL); };
    u64 s = mod_P(hash_code(0), P, M);
    for (u64 steps = 0; steps <= P; steps++) {
        if (!(nul32[s >> 5] & flag_mask(s))) { img[s] = cur; nul32[s >> 5] |= flag_mask(s); *placed = 1; return; }
        const ulonglong2 occ = img[s];
        if (prio(occ.y) > prio(cur.y)) { img[s] = cur; cur = occ; }
        s = (s + 1 == P) ? 0 : s + 1;
    }
}

// exist_kmerset (kmerSet.cpp:216-238) on the device image
__device__ __forceinline__ u64 seed_find(const ulonglong2 *__restrict__ img, const u32 *__restrict__ nul32, u64 P, u64 M, u64 kmer)
{
    u64 s = mod_P(hash_code(kmer), P, M);
    for (;;) {
        if (!(__ldg(nul32 + (s >> 5)) & flag_mask(s))) return SEED_MISS;
        const ulonglong2 nd = __ldg(img + s);
        if (nd.x == kmer) return nd.y;
        s = (s + 1 == P) ? 0 : s + 1;
    }
}

// sink of the extraction kernel: every read position with a k-mer gets its hit word -- the node's value with the `freq`
// bit (known to be 1) replaced by the strand of the READ's k-mer -- or SEED_MISS (absent, or not unique in the contigs)
struct SeedLookupSink {
    static constexpr int RUN = G;
    static constexpr int MIN_BLOCKS = MIN_CTAS;
    TableView t;              // P, M (and the counters k_build adds its statistics to)
    const ulonglong2 *img;
    const u32 *nul32;
    const u64 *offs;          // the block's read offsets
    u64 read_index0, first_base;
    u64 *hits;                // [total_bases]

    __device__ __forceinline__ void init(u32 *) {}
    __device__ __forceinline__ void finish() {}
    __device__ __forceinline__ void consume(const Occ (&o)[G], int nv)
    {
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (g < nv) {
                const u64 ord = o[g].ord >> 1;
                const u64 at = __ldg(offs + ((ord >> 16) - read_index0)) + (ord & 0xFFFFu) - first_base;
                const u64 v = seed_find(img, nul32, t.P, t.M, o[g].klo);
                hits[at] = (v == SEED_MISS || !((v >> 62) & 1)) ? SEED_MISS : ((v & ~(1ULL << 62)) | ((o[g].ord & 1) << 62));
            }
        }
    }
};

// get_align_seed (map_func.cpp:181-237), one warp per read: the first i in [start-1, len-K-S] whose hit words at i and i+S
// are both present, on the same contig, S apart.  out = {contig id, contig start, contig end, read start, read end, 'F'|'R'|'N'}
static __global__ void __launch_bounds__(256) k_seed_pick(const u64 *__restrict__ hits, const u64 *__restrict__ offs, u64 n_reads, u64 first_base,
                                                         const int32_t *__restrict__ starts, int K, int S, int32_t *out)
{
    const u64 r = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const u64 o0 = offs[r] - first_base;
    const long long len = (long long)(offs[r + 1] - offs[r]);
    const long long i_first = starts ? (long long)starts[r] - 1 : 0, i_last = len - K - S;
    int32_t res[6] = {-1, -1, -1, -1, -1, 'N'};
    for (long long i0 = i_first < 0 ? 0 : i_first; i0 <= i_last; i0 += 32) {
        const long long i = i0 + lane;
        bool ok = false;
        u64 a = 0, b = 0;
        if (i <= i_last) {
            a = __ldg(hits + o0 + i);
            if (a != SEED_MISS) {
                b = __ldg(hits + o0 + i + S);
                if (b != SEED_MISS && (u32)a == (u32)b) {
                    const int pa = (int)((a >> 32) & 0x3FFFFFFFu), pb = (int)((b >> 32) & 0x3FFFFFFFu);
                    ok = abs(pb - pa) == S;
                }
            }
        }
        const u32 bal = __ballot_sync(0xffffffffu, ok);
        if (bal) {
            const int w = __ffs(bal) - 1;
            a = __shfl_sync(0xffffffffu, a, w); b = __shfl_sync(0xffffffffu, b, w);
            const long long iw = i0 + w;
            const int pa = (int)((a >> 32) & 0x3FFFFFFFu), pb = (int)((b >> 32) & 0x3FFFFFFFu);
            const bool forward = ((a >> 62) & 1) == (a >> 63);          // strand of the read's k-mer == strand of the contig's
            res[0] = (int32_t)(u32)a;
            res[1] = (forward ? pa : pb) + 1;
            res[2] = (forward ? pb : pa) + K;
            res[3] = (int32_t)iw + 1;
            res[4] = (int32_t)iw + S + K;
            res[5] = forward ? 'F' : 'R';
            break;
        }
    }
    if (lane < 6) out[r * 6 + lane] = res[lane];
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct seedidx_ctx {
    int K = 0, device = 0;
    dbg_ctx *g = nullptr;
    dbg_stats st;
    std::vector<uint32_t> piece_id, piece_start;
    uint64_t next_contig = 0;
    bool finalized = false;
    uint64_t count = 0;
    cudaStream_t stream = nullptr;
    u32 *d_piece_id = nullptr, *d_piece_start = nullptr;
    u64 *d_misc = nullptr;            // [0] payload errors, [1] k-mer 0 placed, [2..] counters of the lookup launches
    // lookup staging
    char *d_bases = nullptr; u64 *d_offs = nullptr, *d_hits = nullptr, *d_chunk_first = nullptr;
    int32_t *d_starts = nullptr, *d_out = nullptr;
    uint64_t cap_bases = 0, cap_reads = 0, cap_chunks = 0;
    uint64_t launches = 0;
};

extern "C" int seedidx_create(seedidx_ctx **out, int32_t K, uint64_t init_slots, float load_factor, int32_t device)
{
    if (!out) return sset_err(DBG_ERR_INVALID, "seedidx_create: NULL argument");
    *out = nullptr;
    if (K < 1 || K > 31) return sset_err(DBG_ERR_INVALID, "K=%d outside 1..31 (the reference's node holds a 64-bit k-mer, kmerSet.h:54)", K);
    dbg_params p;
    memset(&p, 0, sizeof(p));
    p.K = K; p.max_read_len = (int32_t)(PIECE_KMERS + K - 1); p.init_slots = init_slots; p.load_factor = load_factor;
    p.device = device; p.track_order = 1; p.payload_mode = 1;
    seedidx_ctx *c = new seedidx_ctx();
    c->K = K; c->device = device;
    int rc = dbg_create(&c->g, &p);
    if (rc != DBG_OK) { delete c; return sset_err(rc, "dbg_create: %s", dbg_last_error()); }
    *out = c;
    SCU(cudaSetDevice(device));
    SCU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SCU(cudaMalloc(&c->d_misc, (2 + CNT_N) * sizeof(u64)));
    SCU(cudaMemset(c->d_misc, 0, (2 + CNT_N) * sizeof(u64)));
    SDBG(dbg_get_stats(c->g, &c->st));
    return DBG_OK;
}

extern "C" void seedidx_destroy(seedidx_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->g) dbg_destroy(c->g);
    cudaFree(c->d_piece_id); cudaFree(c->d_piece_start); cudaFree(c->d_misc);
    cudaFree(c->d_bases); cudaFree(c->d_offs); cudaFree(c->d_hits); cudaFree(c->d_chunk_first); cudaFree(c->d_starts); cudaFree(c->d_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// chop_contig_to_kmerset for contigs seqs[offs[i] .. offs[i+1]); ids continue from the previous call.  An empty sequence
// keeps its id (map_pair blanks contigs shorter than -l but keeps their index, map_pair.cpp:100-110).
extern "C" int seedidx_add_contigs(seedidx_ctx *c, const char *seqs, const uint64_t *offs, uint64_t n_contigs)
{
    if (!c || (!seqs && n_contigs) || (!offs && n_contigs)) return sset_err(DBG_ERR_INVALID, "seedidx_add_contigs: NULL argument");
    if (c->finalized) return sset_err(DBG_ERR_STATE, "seedidx_add_contigs after seedidx_finalize");
    if (c->next_contig + n_contigs > 0xFFFFFFFFull) return sset_err(DBG_ERR_INVALID, "more than 2^32 contigs (id:32, kmerSet.h:55)");
    const uint64_t K = (uint64_t)c->K;
    std::vector<char> pb;
    std::vector<uint64_t> po(1, 0);
    for (uint64_t q = 0; q < n_contigs; q++) {
        const char *sq = seqs + offs[q];
        const uint64_t len = offs[q + 1] - offs[q];
        if (len >= (1ull << 30)) return sset_err(DBG_ERR_INVALID, "contig %llu is longer than 2^30 (pos:30, kmerSet.h:56)", (unsigned long long)(c->next_contig + q));
        uint64_t i = 0;
        while (i < len) {
            while (i < len && sq[i] == 'N') i++;                   // scaffold_to_contig: only upper-case N separates blocks
            const uint64_t start = i;
            while (i < len && sq[i] != 'N') i++;
            const uint64_t blen = i - start;
            if (blen < K) continue;
            const uint64_t nk = blen - K + 1;
            for (uint64_t k0 = 0; k0 < nk; k0 += PIECE_KMERS) {
                const uint64_t kk = nk - k0 < PIECE_KMERS ? nk - k0 : PIECE_KMERS;
                pb.insert(pb.end(), sq + start + k0, sq + start + k0 + kk + K - 1);
                po.push_back(pb.size());
                c->piece_id.push_back((uint32_t)(c->next_contig + q));
                c->piece_start.push_back((uint32_t)(start + k0));
            }
        }
    }
    c->next_contig += n_contigs;
    if (po.size() > 1) SDBG(dbg_submit_reads(c->g, pb.data(), po.data(), po.size() - 1));
    return DBG_OK;
}

extern "C" int seedidx_finalize(seedidx_ctx *c, uint64_t *size, uint64_t *count, uint64_t *max_cutoff)
{
    if (!c) return sset_err(DBG_ERR_INVALID, "NULL ctx");
    if (!c->finalized) {
        SDBG(dbg_finalize(c->g, &c->st));
        SCU(cudaSetDevice(c->device));
        void *d_img = nullptr, *d_nul = nullptr, *d_nodes = nullptr;
        uint64_t n_local = 0, *d_polyA = nullptr;
        SDBG(dbg_device_image(c->g, &d_img, &d_nul));
        SDBG(dbg_device_build_table(c->g, &d_nodes, &n_local, &d_polyA));
        const size_t np = c->piece_id.size();
        SCU(cudaMalloc(&c->d_piece_id, (np + 1) * sizeof(u32)));
        SCU(cudaMalloc(&c->d_piece_start, (np + 1) * sizeof(u32)));
        if (np) {
            SCU(cudaMemcpyAsync(c->d_piece_id, c->piece_id.data(), np * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
            SCU(cudaMemcpyAsync(c->d_piece_start, c->piece_start.data(), np * sizeof(u32), cudaMemcpyHostToDevice, c->stream));
        }
        const u64 P = c->st.array_size, M = (u64)((((unsigned __int128)1) << 64) / P);
        int n_sms = 148;
        SCU(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, c->device));
        k_seed_payload<<<n_sms * 8, 256, 0, c->stream>>>((const NodeT<false> *)d_nodes, n_local, P, M, (ulonglong2 *)d_img, c->d_piece_id,
                                                          c->d_piece_start, c->d_misc);
        SCU(cudaGetLastError());
        k_seed_kmer0<<<1, 32, 0, c->stream>>>((ulonglong2 *)d_img, (u32 *)d_nul, P, M, (const u64 *)d_polyA, c->d_piece_id, c->d_piece_start, c->d_misc + 1);
        SCU(cudaGetLastError());
        c->launches += 2;
        u64 misc[2];
        SCU(cudaMemcpyAsync(misc, c->d_misc, sizeof(misc), cudaMemcpyDeviceToHost, c->stream));
        SCU(cudaStreamSynchronize(c->stream));
        if (misc[0]) return sset_err(DBG_ERR_STATE, "%llu nodes were not found in the laid-out table", (unsigned long long)misc[0]);
        c->count = c->st.count + misc[1];
        c->finalized = true;
        // add_kmerset enlarges as soon as count >= max when another k-mer arrives (kmerSet.cpp:172-175): refused
        if (c->count >= c->st.max_cutoff)
            return sset_err(DBG_ERR_TABLE_FULL, "%llu k-mers reach the table's limit of %llu (size %llu x load factor): the reference would enlarge; "
                            "size the table like map_pair does (3 x the contig length, load factor 0.5)",
                            (unsigned long long)c->count, (unsigned long long)c->st.max_cutoff, (unsigned long long)c->st.array_size);
    }
    if (size) *size = c->st.array_size;
    if (count) *count = c->count;
    if (max_cutoff) *max_cutoff = c->st.max_cutoff;
    return DBG_OK;
}

// the KmerSet map_pair / map_reads probe (kmerSet.h:64-75): array[size] of 16-byte nodes + nul_flag[size/8+1], MSB first
extern "C" int seedidx_export(seedidx_ctx *c, void *array, uint8_t *nul_flag)
{
    if (!c || !array || !nul_flag) return sset_err(DBG_ERR_INVALID, "NULL argument");
    if (!c->finalized) return sset_err(DBG_ERR_STATE, "seedidx_export needs seedidx_finalize");
    SDBG(dbg_export_kmerset(c->g, array, nul_flag));
    return DBG_OK;
}

static int lookup_device(seedidx_ctx *c, const char *d_bases, const u64 *d_offs, uint64_t n_reads, uint64_t first_base, uint64_t total_bases,
                         const int32_t *d_starts, int seed_kmer_num, int32_t *d_out)
{
    if (n_reads == 0) return DBG_OK;
    void *d_img = nullptr, *d_nul = nullptr;
    SDBG(dbg_device_image(c->g, &d_img, &d_nul));
    const u64 P = c->st.array_size, M = (u64)((((unsigned __int128)1) << 64) / P);
    if (total_bases) {
        uint64_t abase = first_base & ~15ull;
        if (((uintptr_t)(d_bases + abase) & 15) != 0) return sset_err(DBG_ERR_INVALID, "device base buffer must be 16-byte aligned");
        uint64_t n_chunks = (first_base + total_bases - abase + CB - 1) / CB;
        if (n_chunks > 0x7fffffffull) return sset_err(DBG_ERR_INVALID, "block too large");
        if (n_chunks + 1 > c->cap_chunks) {
            SCU(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_chunk_first); c->d_chunk_first = nullptr;
            c->cap_chunks = n_chunks + 1 + n_chunks / 4;
            SCU(cudaMalloc(&c->d_chunk_first, c->cap_chunks * sizeof(u64)));
        }
        SCU(cudaMemsetAsync(c->d_hits, 0xFF, total_bases * sizeof(u64), c->stream));
        k_chunk_first<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, c->stream>>>(d_offs, n_reads, abase, n_chunks, c->d_chunk_first);
        SCU(cudaGetLastError());
        BuildArgs a;
        a.bases = d_bases; a.offs = d_offs; a.n_reads = n_reads; a.abase = abase; a.end_base = first_base + total_bases;
        a.chunk_first = c->d_chunk_first; a.read_index0 = 0; a.K = c->K; a.R = 65535;
        a.stage_words = (uint32_t)(((CB + ((a.R + 15) / 16) * 16) / 16 + 8 + 1) & ~1);
        a.count_stats = 1; a.seed = 1;
        SeedLookupSink sk;
        sk.t.nodes = nullptr; sk.t.P = P; sk.t.M = M; sk.t.lo = 0; sk.t.n_local = 0; sk.t.counters = c->d_misc + 2; sk.t.polyA = nullptr;
        sk.img = (const ulonglong2 *)d_img; sk.nul32 = (const u32 *)d_nul; sk.offs = d_offs; sk.read_index0 = 0; sk.first_base = first_base;
        sk.hits = c->d_hits;
        size_t smem = ((size_t)a.stage_words + MAXR + MAXR + 2) * sizeof(u32);
        if (smem > 48 * 1024) SCU(cudaFuncSetAttribute(k_build<false, SeedLookupSink>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_build<false, SeedLookupSink><<<(unsigned)n_chunks, BLOCK, smem, c->stream>>>(a, sk);
        SCU(cudaGetLastError());
        c->launches += 2;
    }
    k_seed_pick<<<(unsigned)((n_reads * 32 + 255) / 256), 256, 0, c->stream>>>(c->d_hits, d_offs, n_reads, first_base, d_starts, c->K, seed_kmer_num, d_out);
    SCU(cudaGetLastError());
    c->launches++;
    return DBG_OK;
}

// get_align_seed(read, search_start, read.size()) for every read of a host batch: out[6 * i ..] = {contig_id_index,
// seed_contig_start, seed_contig_end, seed_read_start, seed_read_end, 'F'|'R'|'N'}; the first five are -1 when there is no
// seed (also for reads shorter than K + seed_kmer_num, which the callers skip: map_pair.cpp:284).  search_start: 1-based
// first read position per read, or NULL for 1.
extern "C" int seedidx_align_reads(seedidx_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads, const int32_t *search_start,
                                   int32_t seed_kmer_num, int32_t *out)
{
    if (!c || (!bases && n_reads) || (!offs && n_reads) || (!out && n_reads)) return sset_err(DBG_ERR_INVALID, "seedidx_align_reads: NULL argument");
    if (!c->finalized) return sset_err(DBG_ERR_STATE, "seedidx_align_reads needs seedidx_finalize");
    if (seed_kmer_num < 1) return sset_err(DBG_ERR_INVALID, "seed_kmer_num=%d", seed_kmer_num);
    SCU(cudaSetDevice(c->device));
    const uint64_t SUB_BASES = 128ull << 20, SUB_READS = 2ull << 20;
    uint64_t r0 = 0;
    while (r0 < n_reads) {
        uint64_t lim = r0 + SUB_READS < n_reads ? r0 + SUB_READS : n_reads, lo = r0 + 1, hi = lim;
        while (lo < hi) { uint64_t mid = (lo + hi + 1) / 2; if (offs[mid] - offs[r0] <= SUB_BASES) lo = mid; else hi = mid - 1; }
        const uint64_t r1 = lo, nb = offs[r1] - offs[r0], nr = r1 - r0;
        for (uint64_t r = r0; r < r1; r++)
            if (offs[r + 1] - offs[r] > 65535) return sset_err(DBG_ERR_INVALID, "read %llu is longer than 65535 bases", (unsigned long long)r);
        if (nb + 64 > c->cap_bases || nr + 2 > c->cap_reads) {
            SCU(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_bases); cudaFree(c->d_offs); cudaFree(c->d_hits); cudaFree(c->d_starts); cudaFree(c->d_out);
            c->d_bases = nullptr; c->d_offs = nullptr; c->d_hits = nullptr; c->d_starts = nullptr; c->d_out = nullptr;
            c->cap_bases = (nb > (16ull << 20) ? nb : (16ull << 20)) + 64; c->cap_reads = (nr > (1ull << 18) ? nr : (1ull << 18)) + 2;
            SCU(cudaMalloc(&c->d_bases, c->cap_bases));
            SCU(cudaMalloc(&c->d_hits, c->cap_bases * sizeof(u64)));
            SCU(cudaMalloc(&c->d_offs, c->cap_reads * sizeof(u64)));
            SCU(cudaMalloc(&c->d_starts, c->cap_reads * sizeof(int32_t)));
            SCU(cudaMalloc(&c->d_out, c->cap_reads * 6 * sizeof(int32_t)));
        }
        const uint64_t pad = offs[r0] & 15;
        if (nb) SCU(cudaMemcpyAsync(c->d_bases + pad, bases + offs[r0], nb, cudaMemcpyHostToDevice, c->stream));
        SCU(cudaMemcpyAsync(c->d_offs, offs + r0, (nr + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        if (search_start) SCU(cudaMemcpyAsync(c->d_starts, search_start + r0, nr * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        int rc = lookup_device(c, c->d_bases + pad - offs[r0], c->d_offs, nr, offs[r0], nb, search_start ? c->d_starts : nullptr, seed_kmer_num, c->d_out);
        if (rc) return rc;
        SCU(cudaMemcpyAsync(out + 6 * r0, c->d_out, nr * 6 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        SCU(cudaStreamSynchronize(c->stream));
        r0 = r1;
    }
    return DBG_OK;
}

extern "C" uint64_t seedidx_launch_count(const seedidx_ctx *c) { return c ? c->launches + (c->g ? dbg_launch_count(c->g) : 0) : 0; }
