// export_expand.cpp -- host side of the pipelined KmerSet export: expand slot-ordered compact nodes into the P-slot
// table image the reference's traversal probes (kmerSet.h:88-99: array[P] of KmerNode + nul_flag, MSB first).
// Plain C++ (g++ through nvcc): AVX-512 expand-loads where the CPU has them, a branch-free SSE2 loop otherwise; the output
// is written with non-temporal stores (it is 2x the input and is not read again by these threads).
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace dbg {

uint64_t expand_nodes(const uint8_t *bits, uint64_t n_slots, const void *src, void *dst, int node_bytes);

namespace {

// qword mask of the 4 slots of a bitmap nibble (MSB first): node_bytes 16 -> 2 qwords per slot
struct Luts {
    uint8_t q2[16];        // 16-B nodes: nibble -> 8-bit qword mask (4 slots x 2 qwords)
    uint8_t q4[4];         // 32-B nodes: 2 bitmap bits -> 8-bit qword mask (2 slots x 4 qwords)
    Luts()
    {
        for (int n = 0; n < 16; n++) { int m = 0; for (int s = 0; s < 4; s++) if (n & (8 >> s)) m |= 3 << (2 * s); q2[n] = (uint8_t)m; }
        for (int n = 0; n < 4; n++) { int m = 0; for (int s = 0; s < 2; s++) if (n & (2 >> s)) m |= 15 << (4 * s); q4[n] = (uint8_t)m; }
    }
};
const Luts g_luts;

template <int Q>   // qwords per node
uint64_t expand_sse2(const uint8_t *bits, uint64_t n_slots, const uint64_t *src, uint64_t *dst)
{
    const uint64_t *src0 = src;
    for (uint64_t s = 0; s < n_slots; s++) {
        const unsigned o = (bits[s >> 3] >> (7 - (s & 7))) & 1u;
        const __m128i m = _mm_set1_epi32(-(int)o);
        for (int q = 0; q < Q; q += 2) {
            // (the source buffer is padded: a load behind the last node is harmless)
            const __m128i v = _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i *>(src + q)), m);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + q), v);
        }
        src += o * Q; dst += Q;
    }
    return (uint64_t)(src - src0) / Q;
}

__attribute__((target("avx512f"))) uint64_t expand_avx512_16(const uint8_t *bits, uint64_t n_bytes, const uint64_t *src, uint64_t *dst, bool aligned)
{
    const uint64_t *src0 = src;
    for (uint64_t j = 0; j < n_bytes; j++) {
        const unsigned B = bits[j];
        const __mmask8 m0 = g_luts.q2[B >> 4], m1 = g_luts.q2[B & 15];
        const __m512i a = _mm512_maskz_expandloadu_epi64(m0, src); src += __builtin_popcount(m0);
        const __m512i b = _mm512_maskz_expandloadu_epi64(m1, src); src += __builtin_popcount(m1);
        if (aligned) { _mm512_stream_si512(reinterpret_cast<__m512i *>(dst), a); _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + 8), b); }
        else { _mm512_storeu_si512(dst, a); _mm512_storeu_si512(dst + 8, b); }
        dst += 16;
    }
    return (uint64_t)(src - src0) / 2;
}

__attribute__((target("avx512f"))) uint64_t expand_avx512_32(const uint8_t *bits, uint64_t n_bytes, const uint64_t *src, uint64_t *dst, bool aligned)
{
    const uint64_t *src0 = src;
    for (uint64_t j = 0; j < n_bytes; j++) {
        const unsigned B = bits[j];
#pragma GCC unroll 4
        for (int p = 3; p >= 0; p--) {
            const __mmask8 m = g_luts.q4[(B >> (2 * p)) & 3];
            const __m512i a = _mm512_maskz_expandloadu_epi64(m, src); src += __builtin_popcount(m);
            if (aligned) _mm512_stream_si512(reinterpret_cast<__m512i *>(dst), a); else _mm512_storeu_si512(dst, a);
            dst += 8;
        }
    }
    return (uint64_t)(src - src0) / 4;
}

bool have_avx512()
{
    static const bool v = __builtin_cpu_supports("avx512f");
    return v;
}

}  // namespace

uint64_t expand_nodes(const uint8_t *bits, uint64_t n_slots, const void *src_, void *dst_, int node_bytes)
{
    const uint64_t *src = static_cast<const uint64_t *>(src_);
    uint64_t *dst = static_cast<uint64_t *>(dst_);
    const int Q = node_bytes / 8;
    uint64_t used = 0;
    const uint64_t whole = n_slots / 8;        // bitmap bytes covered completely
    if (have_avx512() && whole) {
        const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 63) == 0;
        used = Q == 2 ? expand_avx512_16(bits, whole, src, dst, aligned) : expand_avx512_32(bits, whole, src, dst, aligned);
        src += used * Q; dst += whole * 8 * Q; bits += whole; n_slots -= whole * 8;
    }
    used += Q == 2 ? expand_sse2<2>(bits, n_slots, src, dst) : expand_sse2<4>(bits, n_slots, src, dst);
    _mm_sfence();
    return used;
}

}  // namespace dbg
