// kfreq.cu -- the K-mer frequency table `correct_error` loads.
//
// The reference does not build this table: it only LOADS what the external program `kmerfreq` wrote
// (test/01.clean_correct/work.sh:18-25; correct_error/main_parallel_senior.cpp:273-408 for the 1-bit form,
// correct_error/main.cpp:161-220 for the 8-bit form).  Those loaders are the format contract:
//   <prefix>.kmer.freq.cz      concatenated zlib compress() streams, one per block of 8 Mi k-mers in index order
//                              (1-bit: 1 MiB raw per block, bit 7 - idx%8 of byte idx/8; 8-bit: one saturating
//                              byte per k-mer, 8 MiB raw per block)
//   <prefix>.kmer.freq.cz.len  the compressed byte length of every block, decimal, one per line
//   <prefix>.kmer.freq.stat    the k-mer spectrum (5 '#' lines, blank, 7-column header, rows 1..65535)
// Index of a k-mer = its canonical 2K-bit value (the loaders OR in the reverse complements themselves and
// only do so for idx <= rc(idx), i.e. they expect canonical entries).  Every read position counts; N counts as
// A (seqKmer.cpp alphabet[]).  "High frequency" = count > cutoff (correct_error/main.cpp:202).
// PARITY UNPINNED: kmerfreq (fanagislab/kmerfreq, unpinned, not vendored) is absent; semantics follow the
// consumers and the .stat artefacts in test/01.clean_correct.
//
// GPU side: a direct-index table of u32 counts (4^K x 4 B: 68.7 GB at K=17, fits one B200) bumped with
// fire-and-forget RED.ADD by the same fused extract kernel as the graph build (k_build<FreqSink>); spectrum,
// 1-bit and 8-bit images are produced by streaming kernels; zlib runs on host threads.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

#include "../../include/dbg_b200.h"
#include "dbg_kernels.cuh"

using namespace dbg;

static thread_local char k_err[512] = "";
static int kset_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(k_err, sizeof(k_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" const char *kfreq_last_error(void) { return k_err; }

#define KCU(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return kset_err(e_ == cudaErrorMemoryAllocation ? DBG_ERR_NOMEM : DBG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

static const uint64_t KF_BLOCK = 8ull * 1024 * 1024;      // k-mers per .cz block (SrcBlockSize, main_parallel_senior.cpp:71)

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
struct FreqSink {
    static constexpr int RUN = G;
    static constexpr int MIN_BLOCKS = MIN_CTAS;
    TableView t;           // counters only
    u32 *table;
    u64 lo, hi;            // owned index range [lo, hi)

    __device__ __forceinline__ void init(u32 *) {}
    __device__ __forceinline__ void finish() {}
    __device__ __forceinline__ void consume(const Occ (&o)[G], int nv)
    {
#pragma unroll
        for (int g = 0; g < G; g++)
            if (g < nv && o[g].klo >= lo && o[g].klo < hi) atomicAdd(table + (o[g].klo - lo), 1u);   // RED.ADD, no return
    }
};

// spectrum: hist[f] = number of k-mer species seen f times (f capped at 65535)
__global__ void __launch_bounds__(256) k_kfreq_hist(const u32 *__restrict__ table, u64 n, u64 *hist)
{
    __shared__ u32 sh[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 v = __ldg(table + i);
        if (v == 0) continue;
        if (v < 2048) atomicAdd(&sh[v], 1u);
        else atomicAdd(hist + (v > 65535u ? 65535u : v), 1ULL);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += 256)
        if (sh[i]) atomicAdd(hist + i, (u64)sh[i]);
}

// 1-bit image: bit (7 - idx%8) of byte idx/8 set iff count > cutoff; one warp -> one u32 of 32 k-mers
__global__ void __launch_bounds__(256) k_kfreq_bits(const u32 *__restrict__ table, u64 i_lo, u64 n, u32 cutoff, u32 *bits32)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
        const u64 i = i0 + threadIdx.x;
        bool on = i < n && __ldg(table + i_lo + i) > cutoff;
        u32 bal = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0 && i < n) bits32[i >> 5] = __byte_perm(__brev(bal), 0, 0x0123);
    }
}

// 8-bit image: min(255, count)
__global__ void __launch_bounds__(256) k_kfreq_bytes(const u32 *__restrict__ table, u64 i_lo, u64 n4, u32 *bytes32)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(table + i_lo) + q);
        bytes32[q] = min(v.x, 255u) | (min(v.y, 255u) << 8) | (min(v.z, 255u) << 16) | (min(v.w, 255u) << 24);
    }
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct kfreq_ctx {
    int K, device, max_read_len;
    uint64_t total;            // 4^K
    uint64_t lo, hi;           // owned index range (whole .cz blocks)
    u32 *d_table;
    u64 *d_counters, *d_chunk_first;
    uint64_t cap_chunks;
    char *d_bases;
    u64 *d_offs;
    uint64_t cap_bases, cap_reads;
    cudaStream_t stream;
    uint64_t reads, launches;
};

extern "C" void kfreq_destroy(kfreq_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->d_table); cudaFree(c->d_counters); cudaFree(c->d_chunk_first); cudaFree(c->d_bases); cudaFree(c->d_offs);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int kfreq_create(kfreq_ctx **out, int32_t K, int32_t device, int32_t block_rank, int32_t block_count)
{
    if (!out) return kset_err(DBG_ERR_INVALID, "kfreq_create: NULL");
    *out = nullptr;
    if (K < 1 || K > 17) return kset_err(DBG_ERR_INVALID, "K=%d outside 1..17 (direct-index table)", K);
    int n = block_count > 1 ? block_count : 1;
    if (block_rank < 0 || block_rank >= n) return kset_err(DBG_ERR_INVALID, "block_rank %d / %d", block_rank, n);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return kset_err(DBG_ERR_CUDA, "no CUDA device visible: libdbgb200 has no CPU fallback"); }
    if (device < 0 || device >= ndev) return kset_err(DBG_ERR_INVALID, "device %d of %d", device, ndev);
    KCU(cudaSetDevice(device));
    kfreq_ctx *c = new kfreq_ctx();
    memset(c, 0, sizeof(*c));
    c->K = K; c->device = device; c->max_read_len = 65535;
    c->total = 1ull << (2 * K);
    uint64_t n_blocks = (c->total + KF_BLOCK - 1) / KF_BLOCK;
    uint64_t per = (n_blocks + n - 1) / n;                    // contiguous runs of whole .cz blocks per rank
    c->lo = (uint64_t)block_rank * per * KF_BLOCK; if (c->lo > c->total) c->lo = c->total;
    c->hi = c->lo + per * KF_BLOCK < c->total ? c->lo + per * KF_BLOCK : c->total;
    *out = c;
    KCU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    KCU(cudaMalloc(&c->d_table, (c->hi - c->lo + 4) * sizeof(u32)));
    KCU(cudaMalloc(&c->d_counters, CNT_N * sizeof(u64)));
    KCU(cudaMemsetAsync(c->d_table, 0, (c->hi - c->lo + 4) * sizeof(u32), c->stream));
    KCU(cudaMemsetAsync(c->d_counters, 0, CNT_N * sizeof(u64), c->stream));
    KCU(cudaStreamSynchronize(c->stream));
    return DBG_OK;
}

// clear the table and the counters (a fresh count on the same context: benchmarks, repeated libraries)
extern "C" int kfreq_reset(kfreq_ctx *c)
{
    if (!c) return kset_err(DBG_ERR_INVALID, "NULL ctx");
    KCU(cudaSetDevice(c->device));
    KCU(cudaMemsetAsync(c->d_table, 0, (c->hi - c->lo + 4) * sizeof(u32), c->stream));
    KCU(cudaMemsetAsync(c->d_counters, 0, CNT_N * sizeof(u64), c->stream));
    KCU(cudaStreamSynchronize(c->stream));
    c->reads = 0;
    return DBG_OK;
}

static int kfreq_count_device(kfreq_ctx *c, const char *d_bases, const u64 *d_offs, uint64_t n_reads, uint64_t first_base, uint64_t total_bases)
{
    if (n_reads == 0 || total_bases == 0) return DBG_OK;
    uint64_t abase = first_base & ~15ull;
    if (((uintptr_t)(d_bases + abase) & 15) != 0) return kset_err(DBG_ERR_INVALID, "device base buffer must be 16-byte aligned");
    uint64_t n_chunks = (first_base + total_bases - abase + CB - 1) / CB;
    if (n_chunks > 0x7fffffffull) return kset_err(DBG_ERR_INVALID, "block too large");
    if (n_chunks + 1 > c->cap_chunks) {
        KCU(cudaDeviceSynchronize());
        cudaFree(c->d_chunk_first); c->d_chunk_first = nullptr;
        c->cap_chunks = n_chunks + 1 + n_chunks / 4;
        KCU(cudaMalloc(&c->d_chunk_first, c->cap_chunks * sizeof(u64)));
    }
    k_chunk_first<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, c->stream>>>(d_offs, n_reads, abase, n_chunks, c->d_chunk_first);
    KCU(cudaGetLastError());
    BuildArgs a;
    a.bases = d_bases; a.offs = d_offs; a.n_reads = n_reads; a.abase = abase; a.end_base = first_base + total_bases;
    a.chunk_first = c->d_chunk_first; a.read_index0 = 0; a.K = c->K; a.R = c->max_read_len;
    a.stage_words = (uint32_t)(((CB + ((a.R + 15) / 16) * 16) / 16 + 8 + 1) & ~1);
    a.count_stats = 1; a.seed = 0;
    FreqSink sk;
    sk.t.nodes = nullptr; sk.t.P = 1; sk.t.M = 0; sk.t.lo = 0; sk.t.n_local = 0; sk.t.counters = c->d_counters; sk.t.polyA = nullptr;
    sk.table = c->d_table; sk.lo = c->lo; sk.hi = c->hi;
    size_t smem = ((size_t)a.stage_words + MAXR + MAXR + 2) * sizeof(u32);
    if (smem > 48 * 1024) KCU(cudaFuncSetAttribute(k_build<false, FreqSink>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_build<false, FreqSink><<<(unsigned)n_chunks, BLOCK, smem, c->stream>>>(a, sk);
    KCU(cudaGetLastError());
    c->launches += 2;
    c->reads += n_reads;
    return DBG_OK;
}

extern "C" int kfreq_submit_reads_device(kfreq_ctx *c, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                         uint64_t first_base, uint64_t total_bases)
{
    if (!c || (!d_bases && n_reads) || (!d_offs && n_reads)) return kset_err(DBG_ERR_INVALID, "kfreq_submit_reads_device: NULL argument");
    KCU(cudaSetDevice(c->device));
    return kfreq_count_device(c, d_bases, (const u64 *)d_offs, n_reads, first_base, total_bases);
}

extern "C" int kfreq_submit_reads(kfreq_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads)
{
    if (!c || (!bases && n_reads) || (!offs && n_reads)) return kset_err(DBG_ERR_INVALID, "kfreq_submit_reads: NULL argument");
    KCU(cudaSetDevice(c->device));
    const uint64_t SUB_BASES = 256ull << 20, SUB_READS = 4ull << 20;
    uint64_t r0 = 0;
    while (r0 < n_reads) {
        uint64_t lim = r0 + SUB_READS < n_reads ? r0 + SUB_READS : n_reads, lo = r0 + 1, hi = lim;
        while (lo < hi) { uint64_t mid = (lo + hi + 1) / 2; if (offs[mid] - offs[r0] <= SUB_BASES) lo = mid; else hi = mid - 1; }
        uint64_t r1 = lo, nb = offs[r1] - offs[r0], nr = r1 - r0;
        if (nb + 64 > c->cap_bases || nr + 2 > c->cap_reads) {
            KCU(cudaDeviceSynchronize());
            cudaFree(c->d_bases); cudaFree(c->d_offs); c->d_bases = nullptr; c->d_offs = nullptr;
            c->cap_bases = (nb > SUB_BASES ? nb : SUB_BASES) + 64; c->cap_reads = (nr > SUB_READS ? nr : SUB_READS) + 2;
            KCU(cudaMalloc(&c->d_bases, c->cap_bases));
            KCU(cudaMalloc(&c->d_offs, c->cap_reads * sizeof(u64)));
        }
        KCU(cudaStreamSynchronize(c->stream));          // previous kernel done with the staging buffers
        uint64_t pad = offs[r0] & 15;
        if (nb) KCU(cudaMemcpyAsync(c->d_bases + pad, bases + offs[r0], nb, cudaMemcpyHostToDevice, c->stream));
        KCU(cudaMemcpyAsync(c->d_offs, offs + r0, (nr + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        int rc = kfreq_count_device(c, c->d_bases + pad - offs[r0], c->d_offs, nr, offs[r0], nb);
        if (rc) return rc;
        r0 = r1;
    }
    KCU(cudaStreamSynchronize(c->stream));
    return DBG_OK;
}

extern "C" int kfreq_finalize(kfreq_ctx *c, uint64_t *n_occurrences, uint64_t *n_reads)
{
    if (!c) return kset_err(DBG_ERR_INVALID, "NULL ctx");
    KCU(cudaSetDevice(c->device));
    KCU(cudaStreamSynchronize(c->stream));
    u64 cnt[CNT_N];
    KCU(cudaMemcpy(cnt, c->d_counters, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (n_occurrences) *n_occurrences = cnt[CNT_OCC];
    if (n_reads) *n_reads = c->reads;
    return DBG_OK;
}

extern "C" int kfreq_index_range(kfreq_ctx *c, uint64_t *lo, uint64_t *hi)
{
    if (!c) return kset_err(DBG_ERR_INVALID, "NULL ctx");
    if (lo) *lo = c->lo;
    if (hi) *hi = c->hi;
    return DBG_OK;
}

extern "C" int kfreq_histogram(kfreq_ctx *c, uint64_t hist[65536])
{
    if (!c || !hist) return kset_err(DBG_ERR_INVALID, "NULL argument");
    KCU(cudaSetDevice(c->device));
    u64 *d_h = nullptr;
    KCU(cudaMalloc(&d_h, 65536 * sizeof(u64)));
    cudaError_t e = cudaMemsetAsync(d_h, 0, 65536 * sizeof(u64), c->stream);
    if (e == cudaSuccess) {
        k_kfreq_hist<<<148 * 8, 256, 0, c->stream>>>(c->d_table, c->hi - c->lo, d_h);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(hist, d_h, 65536 * sizeof(u64), cudaMemcpyDeviceToHost);
    cudaFree(d_h);
    KCU(e);
    return DBG_OK;
}

// raw image of k-mers [i_lo, i_lo+n) of this context (indices relative to the owned range), into host memory
static int export_range(kfreq_ctx *c, int bits, uint32_t cutoff, uint64_t i_lo, uint64_t n, uint8_t *dst, u32 *d_tmp)
{
    if (bits == 1) {
        k_kfreq_bits<<<148 * 8, 256, 0, c->stream>>>(c->d_table, i_lo, n, cutoff, d_tmp);
        KCU(cudaGetLastError());
        KCU(cudaMemcpyAsync(dst, d_tmp, (n + 7) / 8, cudaMemcpyDeviceToHost, c->stream));
    } else {
        k_kfreq_bytes<<<148 * 8, 256, 0, c->stream>>>(c->d_table, i_lo, (n + 3) / 4, d_tmp);
        KCU(cudaGetLastError());
        KCU(cudaMemcpyAsync(dst, d_tmp, n, cudaMemcpyDeviceToHost, c->stream));
    }
    c->launches++;
    KCU(cudaStreamSynchronize(c->stream));
    return DBG_OK;
}

extern "C" int kfreq_export(kfreq_ctx *c, int32_t bits, int32_t cutoff, uint8_t *out)
{
    if (!c || !out || (bits != 1 && bits != 8)) return kset_err(DBG_ERR_INVALID, "kfreq_export: bad argument");
    KCU(cudaSetDevice(c->device));
    const uint64_t n = c->hi - c->lo, step = 256ull << 20;     // k-mers per pass
    u32 *d_tmp = nullptr;
    KCU(cudaMalloc(&d_tmp, (bits == 1 ? step / 8 : step) + 64));
    int rc = DBG_OK;
    for (uint64_t i = 0; i < n && rc == DBG_OK; i += step) {
        uint64_t m = n - i < step ? n - i : step;
        rc = export_range(c, bits, (uint32_t)(cutoff < 0 ? 0 : cutoff), i, m, out + (bits == 1 ? i / 8 : i), d_tmp);
    }
    cudaFree(d_tmp);
    return rc;
}

// <prefix>.kmer.freq.cz + .cz.len (+ .stat when this context owns the whole index space)
extern "C" int kfreq_write_cz(kfreq_ctx *c, const char *prefix, int32_t bits, int32_t cutoff)
{
    if (!c || !prefix || (bits != 1 && bits != 8)) return kset_err(DBG_ERR_INVALID, "kfreq_write_cz: bad argument");
    KCU(cudaSetDevice(c->device));
    const uint64_t n = c->hi - c->lo;
    const uint64_t n_blocks = (n + KF_BLOCK - 1) / KF_BLOCK;
    const uint64_t raw_block = bits == 1 ? KF_BLOCK / 8 : KF_BLOCK;
    std::string cz = std::string(prefix) + ".kmer.freq.cz", len = cz + ".len";
    FILE *fz = fopen(cz.c_str(), "wb"), *fl = fopen(len.c_str(), "w");
    if (!fz || !fl) { if (fz) fclose(fz); if (fl) fclose(fl); return kset_err(DBG_ERR_INVALID, "cannot open %s for writing", cz.c_str()); }
    // groups of blocks: device -> host raw image, zlib compress() on host threads, append in index order
    const uint64_t group = bits == 1 ? 64 : 16;
    std::vector<uint8_t> raw(group * raw_block);
    std::vector<std::vector<uint8_t>> comp(group);
    u32 *d_tmp = nullptr;
    cudaError_t e = cudaMalloc(&d_tmp, group * raw_block + 64);
    if (e != cudaSuccess) { fclose(fz); fclose(fl); KCU(e); }
    int rc = DBG_OK;
    unsigned nt = std::thread::hardware_concurrency(); if (nt == 0) nt = 1; if (nt > 32) nt = 32;
    for (uint64_t b0 = 0; b0 < n_blocks && rc == DBG_OK; b0 += group) {
        uint64_t nb = n_blocks - b0 < group ? n_blocks - b0 : group;
        uint64_t i_lo = b0 * KF_BLOCK, m = (i_lo + nb * KF_BLOCK <= n) ? nb * KF_BLOCK : n - i_lo;
        rc = export_range(c, bits, (uint32_t)(cutoff < 0 ? 0 : cutoff), i_lo, m, raw.data(), d_tmp);
        if (rc) break;
        std::vector<std::thread> th;
        std::vector<int> zrc(nb, Z_OK);
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                for (uint64_t j = t; j < nb; j += nt) {
                    uint64_t kmers = (i_lo + (j + 1) * KF_BLOCK <= n) ? KF_BLOCK : n - (i_lo + j * KF_BLOCK);
                    uLong src_len = (uLong)(bits == 1 ? (kmers + 7) / 8 : kmers);
                    uLongf dst_len = compressBound(src_len);
                    comp[j].resize(dst_len);
                    zrc[j] = compress(comp[j].data(), &dst_len, raw.data() + j * raw_block, src_len);
                    comp[j].resize(dst_len);
                }
            });
        for (auto &t : th) t.join();
        for (uint64_t j = 0; j < nb; j++) {
            if (zrc[j] != Z_OK) { rc = kset_err(DBG_ERR_INVALID, "zlib compress failed (%d)", zrc[j]); break; }
            fwrite(comp[j].data(), 1, comp[j].size(), fz);
            fprintf(fl, "%llu\n", (unsigned long long)comp[j].size());
        }
    }
    cudaFree(d_tmp);
    fclose(fz); fclose(fl);
    if (rc) return rc;
    if (c->lo == 0 && c->hi == c->total) {
        // the spectrum file, same layout as test/01.clean_correct/*.kmer.freq.stat
        std::vector<uint64_t> h(65536);
        rc = kfreq_histogram(c, h.data());
        if (rc) return rc;
        std::string st = std::string(prefix) + ".kmer.freq.stat";
        FILE *fs = fopen(st.c_str(), "w");
        if (!fs) return kset_err(DBG_ERR_INVALID, "cannot open %s", st.c_str());
        // species/individuals.  NB counts above 65535 are folded into the last row, like a 16-bit saturating counter
        double species = 0, indiv = 0;
        for (uint32_t f = 1; f < 65536; f++) { species += (double)h[f]; indiv += (double)h[f] * f; }
        fprintf(fs, "#Kmer size: %d\n#Maximum Kmer frequency: 65535\n#Kmer indivdual number: %.0f\n#Kmer species number: %.0f\n", c->K, indiv, species);
        fprintf(fs, "#Theoretic space of Kmer species: %llu  occupied ratio: %g\n\n", (unsigned long long)c->total, species / (double)c->total);
        fprintf(fs, "#Kmer_Frequency\tKmer_Species_Number\tKmer_Species_Ratio\tKmer_Species_accumulate_Ratio\tKmer_Individual_Number\tKmer_Individual_Ratio\tKmer_Individual_accumulate_ratio\n");
        double as = 0, ai = 0;
        for (uint32_t f = 1; f < 65536; f++) {
            double s = (double)h[f], i = (double)h[f] * f;
            as += s; ai += i;
            fprintf(fs, "%u\t%llu\t%g\t%g\t%.0f\t%g\t%g\n", f, (unsigned long long)h[f], species > 0 ? s / species : 0.0, species > 0 ? as / species : 0.0,
                    i, indiv > 0 ? i / indiv : 0.0, indiv > 0 ? ai / indiv : 0.0);
        }
        fclose(fs);
    }
    return DBG_OK;
}
