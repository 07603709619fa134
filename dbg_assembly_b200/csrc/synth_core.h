// synth_core.h -- the counter-based synthetic read generator, usable from CUDA and from plain C++:
// the same bytes on the host and on the device, so the CPU reference and the GPU path see identical inputs.
//
//   genome base p      = 2 bits of splitmix64(seed, p / 32)
//   read i (paired)    : pair = i/2 draws a fragment start and strand; mate 0 reads the fragment's first
//                        read_len bases, mate 1 the reverse complement of its last read_len bases
//   substitutions / N  : per (read, position) hash against err_per_2p24 / n_per_2p24
// Included by csrc/synth.cu (libdbgb200: dbg_synth_reads_host/_device) and by oracle/tools/synth_fasta.cpp, the
// stand-alone FASTA writer the reference arm of bench.py uses (so that arm loads nothing from the product library).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SYNTH_HD __host__ __device__
#else
#define SYNTH_HD
#endif

struct synth_params_t {      // == dbg_synth_params (include/dbg_b200.h)
    uint64_t seed;
    uint64_t genome_len;
    uint32_t read_len;
    uint32_t insert;
    uint32_t err_per_2p24;
    uint32_t n_per_2p24;
};

namespace synth {


SYNTH_HD inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

SYNTH_HD inline uint32_t genome_base(uint64_t seed, uint64_t p)
{
    uint64_t w = mix64(seed * 0xD1342543DE82EF95ULL + (p >> 5));
    return (uint32_t)(w >> (2 * (p & 31))) & 3u;
}

SYNTH_HD inline uint64_t mulhi64(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

SYNTH_HD inline char synth_base(const synth_params_t &p, uint64_t read, uint32_t t)
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

}   // namespace synth
