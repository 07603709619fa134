// synth.cu -- deterministic synthetic reads: a counter-based generator that gives the
// same bytes on the host and on the device, so CPU oracle and GPU path see identical inputs and the
// human-scale config can be generated directly in HBM.
//
//   genome base p      = 2 bits of splitmix64(seed, p / 32)
//   read i (paired)    : pair = i/2 draws a fragment start and strand; mate 0 reads the fragment's first
//                        read_len bases, mate 1 the reverse complement of its last read_len bases
//   substitutions / N  : per (read, position) hash against err_per_2p24 / n_per_2p24
#include <cstdint>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/dbg_b200.h"

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ inline uint32_t genome_base(uint64_t seed, uint64_t p)
{
    uint64_t w = mix64(seed * 0xD1342543DE82EF95ULL + (p >> 5));
    return (uint32_t)(w >> (2 * (p & 31))) & 3u;
}

__host__ __device__ inline uint64_t mulhi64(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

__host__ __device__ inline char synth_base(const dbg_synth_params &p, uint64_t read, uint32_t t)
{
    const char letters[4] = {'A', 'C', 'G', 'T'};
    uint64_t pair = read >> 1;
    uint32_t mate = (uint32_t)(read & 1);
    uint64_t h = mix64(p.seed ^ mix64(pair + 0x5851F42D4C957F2DULL));
    uint32_t ins = p.insert < p.read_len ? p.read_len : p.insert;
    uint64_t span = p.genome_len - ins + 1;
    uint64_t f = mulhi64(mix64(h), span);
    uint32_t flip = (uint32_t)(h & 1);
    // which end of the fragment, which strand
    bool rc = (mate ^ flip) != 0;
    uint64_t start = rc ? f + ins - p.read_len : f;
    uint32_t b = rc ? 3u - genome_base(p.seed, start + (p.read_len - 1 - t)) : genome_base(p.seed, start + t);
    uint64_t e = mix64((p.seed + 0x632BE59BD9B4E019ULL) ^ (read * 65536ULL + t));
    if ((uint32_t)(e >> 40) < p.err_per_2p24) b = (b + 1 + (uint32_t)((e >> 8) % 3)) & 3u;
    if ((uint32_t)((e >> 16) & 0xFFFFFFu) < p.n_per_2p24) return 'N';
    return letters[b];
}

__global__ void k_synth(dbg_synth_params p, uint64_t first_read, uint64_t n_reads, char *out)
{
    uint64_t total = n_reads * p.read_len;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        uint64_t r = i / p.read_len;
        uint32_t t = (uint32_t)(i - r * p.read_len);
        out[i] = synth_base(p, first_read + r, t);
    }
}

bool params_ok(const dbg_synth_params *p)
{
    if (!p || p->read_len == 0 || p->read_len > 65535) return false;
    uint32_t ins = p->insert < p->read_len ? p->read_len : p->insert;
    return p->genome_len >= ins;
}

}  // namespace

extern "C" int dbg_synth_reads_host(const dbg_synth_params *p, uint64_t first_read, uint64_t n_reads, char *out)
{
    if (!params_ok(p) || (!out && n_reads)) return DBG_ERR_INVALID;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    if (n_reads < 4096) nt = 1;
    const dbg_synth_params prm = *p;
    auto work = [&](uint64_t r0, uint64_t r1) {
        for (uint64_t r = r0; r < r1; r++)
            for (uint32_t t = 0; t < prm.read_len; t++) out[r * prm.read_len + t] = synth_base(prm, first_read + r, t);
    };
    std::vector<std::thread> th;
    uint64_t per = (n_reads + nt - 1) / nt;
    for (unsigned i = 1; i < nt; i++) {
        uint64_t r0 = i * per, r1 = r0 + per < n_reads ? r0 + per : n_reads;
        if (r0 < r1) th.emplace_back(work, r0, r1);
    }
    work(0, per < n_reads ? per : n_reads);
    for (auto &t : th) t.join();
    return DBG_OK;
}

extern "C" int dbg_synth_reads_device(const dbg_synth_params *p, uint64_t first_read, uint64_t n_reads, char *d_out,
                                      int32_t device, void *stream)
{
    if (!params_ok(p) || (!d_out && n_reads)) return DBG_ERR_INVALID;
    if (n_reads == 0) return DBG_OK;
    if (cudaSetDevice(device) != cudaSuccess) return DBG_ERR_CUDA;
    k_synth<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(*p, first_read, n_reads, d_out);
    return cudaGetLastError() == cudaSuccess ? DBG_OK : DBG_ERR_CUDA;
}
