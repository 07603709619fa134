// export_pipe.cu -- pipelined hand-over of the KmerSet image (dbg_export_kmerset; the seam of the reference is the global
// `KmerSet *kset`, DBGgraph.h:31, filled by build_debruijn_graph and probed by contig.cpp).
//
// The image is P x 16 B (32 B on the wide path) and about half of it is empty slots; one PCIe link moves ~50 GB/s, so the
// plain copy costs more than the whole build.  Here:
//   device : occupied nodes are compacted in slot order, chunk by chunk (popcount of the occupancy words -> tile counts ->
//            offsets -> ordered gather; three small streaming kernels, ~1 ms for 200 M slots, in the shadow of the copies);
//   link   : a chunk travels either COMPACT (its nodes only, into a pinned ring slot) or PLAIN (image bytes straight into
//            the caller's array when that is pinned memory);
//   host   : worker threads expand compact chunks into the array (export_expand.cpp) while later chunks are on the link.
// The choice is made chunk by chunk when the copy is enqueued: compact while a ring slot is free, plain otherwise -- so the
// split follows whatever the host threads can absorb, and the result is byte-identical to the plain copy either way.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "export_pipe.h"

namespace dbg {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int XT_WORDS = 256;                 // occupancy words per tile (one CTA): 8192 slots
constexpr int XT_SLOTS = XT_WORDS * 32;

// occupancy word as the layout kernels write it (MSB-first bytes) -> bit j = slot j of the word
__device__ __forceinline__ u32 slot_bits(u32 w) { return __brev(__byte_perm(w, 0, 0x0123)); }

static __global__ void __launch_bounds__(XT_WORDS) k_exp_tilecount(const u32 *__restrict__ nul32, u64 n_words, u32 *__restrict__ tile_cnt)
{
    __shared__ u32 s_w[XT_WORDS / 32];
    const u64 w = (u64)blockIdx.x * XT_WORDS + threadIdx.x;
    u32 c = w < n_words ? __popc(nul32[w]) : 0u;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { u32 t = 0; for (int i = 0; i < XT_WORDS / 32; i++) t += s_w[i]; tile_cnt[blockIdx.x] = t; }
}

// one CTA per chunk: exclusive scan of its tiles' counts (offset of each tile inside the chunk) and the chunk total
static __global__ void __launch_bounds__(256) k_exp_chunkscan(const u32 *__restrict__ tile_cnt, u64 n_tiles, u32 tiles_per_chunk,
                                                               u32 *__restrict__ tile_off, u64 *__restrict__ chunk_cnt)
{
    __shared__ u32 s_w[8];
    __shared__ u32 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const u64 t0 = (u64)blockIdx.x * tiles_per_chunk;
    for (u32 base = 0; base < tiles_per_chunk; base += 256) {
        const u64 t = t0 + base + threadIdx.x;
        const bool in = base + threadIdx.x < tiles_per_chunk && t < n_tiles;
        const u32 c = in ? tile_cnt[t] : 0u;
        u32 incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= (u32)d) incl += v; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        u32 off = s_carry + incl - c;
        for (u32 w = 0; w < (threadIdx.x >> 5); w++) off += s_w[w];
        if (in) tile_off[t] = off;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = off + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_cnt[blockIdx.x] = s_carry;
}

// ordered gather of one range of tiles: node of occupied slot s -> out[chunk_goff[chunk] + tile_off[tile] + rank in tile]
template <int NB>   // node bytes: 16 or 32
static __global__ void __launch_bounds__(XT_WORDS) k_exp_gather(const void *__restrict__ img, const u32 *__restrict__ nul32, u64 n_words, u64 P,
                                                                 const u32 *__restrict__ tile_off, const u64 *__restrict__ chunk_goff,
                                                                 u32 tiles_per_chunk, u64 tile0, void *__restrict__ out)
{
    __shared__ u32 s_bits[XT_WORDS], s_off[XT_WORDS], s_w[XT_WORDS / 32];
    const u64 tile = tile0 + blockIdx.x;
    const u64 w = tile * XT_WORDS + threadIdx.x;
    const u32 bits = w < n_words ? slot_bits(nul32[w]) : 0u;
    const u32 c = __popc(bits);
    u32 incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= (u32)d) incl += v; }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 off = incl - c;
    for (u32 q = 0; q < (threadIdx.x >> 5); q++) off += s_w[q];
    s_bits[threadIdx.x] = bits; s_off[threadIdx.x] = off;
    __syncthreads();
    const u64 base = chunk_goff[tile / tiles_per_chunk] + tile_off[tile];
    const u64 slot0 = tile * XT_SLOTS;
    const ulonglong2 *src = static_cast<const ulonglong2 *>(img);
    ulonglong2 *dst = static_cast<ulonglong2 *>(out);
#pragma unroll 4
    for (int it = 0; it < XT_SLOTS / XT_WORDS; it++) {
        const u32 sl = it * XT_WORDS + threadIdx.x;         // a warp covers exactly one occupancy word
        const u32 b = s_bits[sl >> 5];
        const u32 j = sl & 31;
        if ((b >> j) & 1u) {
            const u64 d = base + s_off[sl >> 5] + __popc(b & ((1u << j) - 1u));
            const u64 s = slot0 + sl;
            if (s < P) {
                if (NB == 16) dst[d] = __ldcs(src + s);
                else { dst[2 * d] = __ldcs(src + 2 * s); dst[2 * d + 1] = __ldcs(src + 2 * s + 1); }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
struct Job { uint64_t chunk; int slot; };

struct ExportPipe {
    int device = 0;
    cudaStream_t copy_stream = nullptr;
    // device scratch (grow-only)
    u32 *d_tile_cnt = nullptr, *d_tile_off = nullptr;
    u64 *d_chunk_cnt = nullptr;         // [n_chunks] counts, then [n_chunks + 1] global offsets
    uint64_t cap_tiles = 0, cap_chunks = 0;
    void *d_compact = nullptr;
    uint64_t cap_compact = 0;           // bytes
    u64 *h_chunk = nullptr;             // pinned: counts + offsets
    uint64_t cap_h_chunk = 0;
    // pinned ring + bitmap
    std::vector<char *> ring;
    uint64_t ring_slot_bytes = 0;
    uint8_t *h_bits = nullptr;
    uint64_t cap_bits = 0;
    std::vector<cudaEvent_t> ev_gather, ev_copy;
    // workers
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> jobs;
    std::vector<int> free_slots;
    std::atomic<uint64_t> done{0};
    bool stop = false;
    // the export in flight (read by the workers)
    uint64_t P = 0, chunk_slots = 0;
    int nb = 16;
    char *array = nullptr;
    uint8_t *nul_flag = nullptr;
    const u64 *cnt = nullptr;
};

static void worker_main(ExportPipe *p)
{
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv.wait(lk, [&] { return p->stop || !p->jobs.empty(); });
            if (p->jobs.empty()) return;       // stop
            j = p->jobs.front(); p->jobs.pop_front();
        }
        const uint64_t s0 = j.chunk * p->chunk_slots;
        const uint64_t n = s0 + p->chunk_slots <= p->P ? p->chunk_slots : p->P - s0;
        const uint64_t b0 = s0 / 8, nbytes = (n + 7) / 8;
        memcpy(p->nul_flag + b0, p->h_bits + b0, nbytes);
        expand_nodes(p->h_bits + b0, n, p->ring[j.slot], p->array + s0 * (uint64_t)p->nb, p->nb);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->free_slots.push_back(j.slot);
        }
        p->done.fetch_add(1, std::memory_order_release);
    }
}

void export_pipe_destroy(ExportPipe *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv.notify_all();
    for (auto &t : p->workers) t.join();
    cudaSetDevice(p->device);
    for (char *r : p->ring) cudaFreeHost(r);
    if (p->h_bits) cudaFreeHost(p->h_bits);
    if (p->h_chunk) cudaFreeHost(p->h_chunk);
    cudaFree(p->d_tile_cnt); cudaFree(p->d_tile_off); cudaFree(p->d_chunk_cnt); cudaFree(p->d_compact);
    for (cudaEvent_t e : p->ev_gather) cudaEventDestroy(e);
    for (cudaEvent_t e : p->ev_copy) cudaEventDestroy(e);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    delete p;
}

#define XP_TRY(call)                                                                                         \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) { snprintf(err, err_len, "%s: %s", #call, cudaGetErrorString(e_)); return -1; } \
    } while (0)
// allocation failures make the pipe step aside (plain copy), they are not errors
#define XP_ALLOC(call)                                                \
    do {                                                              \
        if ((call) != cudaSuccess) { cudaGetLastError(); return 1; }  \
    } while (0)

static uint64_t env_u64(const char *name, uint64_t dflt)
{
    const char *e = getenv(name);
    return e && *e ? strtoull(e, nullptr, 10) : dflt;
}

int export_pipe_run(ExportPipe **pp, int device, cudaStream_t stream, const void *d_img, const uint32_t *d_nul32, uint64_t nul_words,
                    uint64_t P, int nb, void *array, uint8_t *nul_flag, float *ms, uint64_t stats[4], char *err, int err_len)
{
    // knobs (read per call: tests vary them inside one process)
    uint64_t chunk_slots = env_u64("DBG_B200_EXPORT_CHUNK", 1ull << 20);
    chunk_slots = (chunk_slots + XT_SLOTS - 1) / XT_SLOTS * XT_SLOTS;
    unsigned hw = std::thread::hardware_concurrency();
    int n_threads = (int)env_u64("DBG_B200_EXPORT_THREADS", hw > 1 ? (hw - 1 < 16 ? hw - 1 : 16) : 0);
    if (n_threads <= 0 || P < 4 * chunk_slots || (reinterpret_cast<uintptr_t>(array) & 15) != 0) return 1;
    if (n_threads > 64) n_threads = 64;
    const uint64_t plain_pct = env_u64("DBG_B200_EXPORT_PLAIN_PCT", 0);      // experiments: this share of the chunks goes plain, evenly spread
    int n_slots_ring = (int)env_u64("DBG_B200_EXPORT_SLOTS", (uint64_t)n_threads + 4);
    if (n_slots_ring < 1) n_slots_ring = 1;
    const uint64_t n_chunks = (P + chunk_slots - 1) / chunk_slots;
    const u32 tiles_per_chunk = (u32)(chunk_slots / XT_SLOTS);
    const uint64_t n_tiles = n_chunks * tiles_per_chunk;
    const uint64_t slot_bytes = chunk_slots * (uint64_t)nb + 64;
    const uint64_t bits_bytes = P / 8 + 1;

    const auto t_begin = std::chrono::steady_clock::now();
    ExportPipe *p = *pp;
    if (!p) {
        p = new ExportPipe();
        p->device = device;
        *pp = p;
        XP_TRY(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    }
    // is the destination pinned (cudaHostAlloc / cudaHostRegister)?  Only then may the DMA engine write it directly.
    bool array_pinned = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, array) == cudaSuccess) array_pinned = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
        if (getenv("DBG_B200_EXPORT_NO_DIRECT")) array_pinned = false;
    }
    // ---- scratch (grow-only; the first export of a context pays for the pinned ring) ----
    if (n_tiles > p->cap_tiles) {
        cudaFree(p->d_tile_cnt); cudaFree(p->d_tile_off); p->d_tile_cnt = p->d_tile_off = nullptr; p->cap_tiles = 0;
        XP_ALLOC(cudaMalloc(&p->d_tile_cnt, n_tiles * sizeof(u32)));
        XP_ALLOC(cudaMalloc(&p->d_tile_off, n_tiles * sizeof(u32)));
        p->cap_tiles = n_tiles;
    }
    if (n_chunks > p->cap_chunks) {
        cudaFree(p->d_chunk_cnt); p->d_chunk_cnt = nullptr; p->cap_chunks = 0;
        if (p->h_chunk) { cudaFreeHost(p->h_chunk); p->h_chunk = nullptr; }
        XP_ALLOC(cudaMalloc(&p->d_chunk_cnt, (2 * n_chunks + 2) * sizeof(u64)));
        XP_ALLOC(cudaMallocHost(&p->h_chunk, (2 * n_chunks + 2) * sizeof(u64)));
        p->cap_chunks = n_chunks;
    }
    if (bits_bytes > p->cap_bits) {
        if (p->h_bits) { cudaFreeHost(p->h_bits); p->h_bits = nullptr; } p->cap_bits = 0;
        XP_ALLOC(cudaMallocHost(&p->h_bits, bits_bytes + 64));
        p->cap_bits = bits_bytes;
    }
    if (slot_bytes != p->ring_slot_bytes || (int)p->ring.size() < n_slots_ring) {
        if (slot_bytes != p->ring_slot_bytes) { for (char *r : p->ring) cudaFreeHost(r); p->ring.clear(); }
        p->ring_slot_bytes = slot_bytes;
        while ((int)p->ring.size() < n_slots_ring) {
            char *r = nullptr;
            if (cudaMallocHost(&r, slot_bytes) != cudaSuccess) { cudaGetLastError(); break; }
            p->ring.push_back(r);
        }
        if (p->ring.empty()) return 1;
    }
    const int ring_n = (int)p->ring.size() < n_slots_ring ? (int)p->ring.size() : n_slots_ring;
    while (p->ev_gather.size() < n_chunks) { cudaEvent_t e; XP_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_gather.push_back(e); }
    while (p->ev_copy.size() < n_chunks) { cudaEvent_t e; XP_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_copy.push_back(e); }
    while ((int)p->workers.size() < n_threads) p->workers.emplace_back(worker_main, p);

    // ---- counts and offsets; the bitmap travels meanwhile ----
    cudaEvent_t ev0 = p->ev_copy[0];
    XP_TRY(cudaEventRecord(ev0, stream));
    XP_TRY(cudaStreamWaitEvent(p->copy_stream, ev0, 0));                  // the image is complete on `stream`
    XP_TRY(cudaMemcpyAsync(p->h_bits, d_nul32, bits_bytes, cudaMemcpyDeviceToHost, p->copy_stream));
    k_exp_tilecount<<<(unsigned)n_tiles, XT_WORDS, 0, stream>>>(d_nul32, nul_words, p->d_tile_cnt);
    XP_TRY(cudaGetLastError());
    k_exp_chunkscan<<<(unsigned)n_chunks, 256, 0, stream>>>(p->d_tile_cnt, n_tiles, tiles_per_chunk, p->d_tile_off, p->d_chunk_cnt);
    XP_TRY(cudaGetLastError());
    XP_TRY(cudaMemcpyAsync(p->h_chunk, p->d_chunk_cnt, n_chunks * sizeof(u64), cudaMemcpyDeviceToHost, stream));
    XP_TRY(cudaStreamSynchronize(stream));
    u64 *cnt = p->h_chunk, *goff = p->h_chunk + n_chunks;
    goff[0] = 0;
    for (uint64_t c = 0; c < n_chunks; c++) goff[c + 1] = goff[c] + cnt[c];
    const uint64_t total = goff[n_chunks];
    if (total * (uint64_t)nb + 64 > p->cap_compact) {
        cudaFree(p->d_compact); p->d_compact = nullptr; p->cap_compact = 0;
        const uint64_t want = total * (uint64_t)nb + 64 + (total * (uint64_t)nb) / 16;
        XP_ALLOC(cudaMalloc(&p->d_compact, want));
        p->cap_compact = want;
    }
    XP_TRY(cudaMemcpyAsync(p->d_chunk_cnt + n_chunks, goff, (n_chunks + 1) * sizeof(u64), cudaMemcpyHostToDevice, stream));
    // gather, chunk by chunk, each followed by an event the copies wait on
    for (uint64_t c = 0; c < n_chunks; c++) {
        if (nb == 16) k_exp_gather<16><<<tiles_per_chunk, XT_WORDS, 0, stream>>>(d_img, d_nul32, nul_words, P, p->d_tile_off, p->d_chunk_cnt + n_chunks,
                                                                               tiles_per_chunk, c * tiles_per_chunk, p->d_compact);
        else k_exp_gather<32><<<tiles_per_chunk, XT_WORDS, 0, stream>>>(d_img, d_nul32, nul_words, P, p->d_tile_off, p->d_chunk_cnt + n_chunks,
                                                                        tiles_per_chunk, c * tiles_per_chunk, p->d_compact);
        XP_TRY(cudaGetLastError());
        XP_TRY(cudaEventRecord(p->ev_gather[c], stream));
    }

    // ---- the pipeline ----
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->P = P; p->chunk_slots = chunk_slots; p->nb = nb; p->array = static_cast<char *>(array); p->nul_flag = nul_flag; p->cnt = cnt;
        p->free_slots.clear();
        for (int i = 0; i < ring_n; i++) p->free_slots.push_back(i);
        p->done.store(0);
    }
    struct Flight { uint64_t chunk; int slot; };      // slot < 0: plain copy
    std::deque<Flight> flight;
    uint64_t next = 0, n_done_plain = 0, n_compact = 0, n_plain = 0, n_pushed = 0, bytes_link = bits_bytes;
    const int MAX_FLIGHT = 3;
    int rc = 0;
    while (n_done_plain + p->done.load(std::memory_order_acquire) < n_chunks) {
        bool progressed = false;
        while (next < n_chunks && (int)flight.size() < MAX_FLIGHT) {
            const uint64_t s0 = next * chunk_slots;
            const uint64_t n = s0 + chunk_slots <= P ? chunk_slots : P - s0;
            int slot = -1;
            const bool force_plain = array_pinned && plain_pct && ((next + 1) * plain_pct / 100 != next * plain_pct / 100);
            if (cnt[next] < n && !force_plain) {          // (a completely full chunk gains nothing from the detour)
                std::lock_guard<std::mutex> lk(p->mu);
                if (!p->free_slots.empty()) { slot = p->free_slots.back(); p->free_slots.pop_back(); }
            }
            if (slot < 0 && !array_pinned && cnt[next] < n) break;        // wait for a ring slot
            cudaError_t e;
            if (slot >= 0) {
                e = cudaStreamWaitEvent(p->copy_stream, p->ev_gather[next], 0);
                if (e == cudaSuccess && cnt[next])
                    e = cudaMemcpyAsync(p->ring[slot], static_cast<const char *>(p->d_compact) + goff[next] * (uint64_t)nb, cnt[next] * (uint64_t)nb,
                                        cudaMemcpyDeviceToHost, p->copy_stream);
                bytes_link += cnt[next] * (uint64_t)nb; n_compact++;
            } else {
                e = cudaMemcpyAsync(static_cast<char *>(array) + s0 * (uint64_t)nb, static_cast<const char *>(d_img) + s0 * (uint64_t)nb, n * (uint64_t)nb,
                                    cudaMemcpyDeviceToHost, p->copy_stream);
                bytes_link += n * (uint64_t)nb; n_plain++;
            }
            if (e == cudaSuccess) e = cudaEventRecord(p->ev_copy[next], p->copy_stream);
            if (e != cudaSuccess) { snprintf(err, err_len, "export pipe copy: %s", cudaGetErrorString(e)); rc = -1; break; }
            flight.push_back({next, slot});
            next++;
            progressed = true;
        }
        if (rc) break;
        while (!flight.empty()) {
            cudaError_t q = cudaEventQuery(p->ev_copy[flight.front().chunk]);
            if (q == cudaErrorNotReady) { cudaGetLastError(); break; }
            if (q != cudaSuccess) { snprintf(err, err_len, "export pipe: %s", cudaGetErrorString(q)); rc = -1; break; }
            const Flight f = flight.front(); flight.pop_front();
            // (the bitmap copy precedes every chunk copy on the copy stream: h_bits is complete here)
            if (f.slot >= 0) {
                { std::lock_guard<std::mutex> lk(p->mu); p->jobs.push_back({f.chunk, f.slot}); }
                p->cv.notify_one();
                n_pushed++;
            } else {
                const uint64_t s0 = f.chunk * chunk_slots;
                const uint64_t n = s0 + chunk_slots <= P ? chunk_slots : P - s0;
                memcpy(nul_flag + s0 / 8, p->h_bits + s0 / 8, (n + 7) / 8);
                n_done_plain++;
            }
            progressed = true;
        }
        if (rc) break;
        if (!progressed) std::this_thread::yield();
    }
    if (rc) {
        // drain: nothing may still reference the caller's buffers when we return
        cudaStreamSynchronize(p->copy_stream);
        while (p->done.load(std::memory_order_acquire) < n_pushed) std::this_thread::yield();
        return rc;
    }
    // the byte that holds slot P (nul_flag has P/8 + 1 bytes) and any partial last byte were copied with the last chunk
    nul_flag[P / 8] = p->h_bits[P / 8];
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    if (ms) *ms = (float)(dt * 1e3);
    if (stats) { stats[0] = n_compact; stats[1] = n_plain; stats[2] = bytes_link; stats[3] = total; }
    return 0;
}

}  // namespace dbg
