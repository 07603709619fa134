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

#include "synth_core.h"
static_assert(sizeof(synth_params_t) == sizeof(dbg_synth_params), "synth_params_t mirrors dbg_synth_params");

namespace {

using synth::synth_base;
inline const synth_params_t &core(const dbg_synth_params &p) { return reinterpret_cast<const synth_params_t &>(p); }

__global__ void k_synth(synth_params_t p, uint64_t first_read, uint64_t n_reads, char *out)
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
    const synth_params_t prm = core(*p);
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
    k_synth<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(core(*p), first_read, n_reads, d_out);
    return cudaGetLastError() == cudaSuccess ? DBG_OK : DBG_ERR_CUDA;
}
