// dbg_core.cuh -- device/host helpers shared by the build kernels (sm_100a).
//
// Reference semantics restated here (paths relative to fanagislab/DBG_assembly, DBG_contig/):
//   base codes          seqKmer.cpp:9-19      A,a,N,n->0  C,c->1  G,g->2  T,t->3
//   hash_code           kmerSet.h:105-116     Jenkins/Wang 64-bit mix
//   reverse complement  seqKmer.cpp:89-97
//   link lanes          kmerSet.cpp:56,341    lane of base b = bits (3-b)*8 .. +7, saturating at 255
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

namespace dbg {

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- device table node ------------------------------------------------------------------------------
// K <= 31: 32 B = exactly ONE DRAM sector   { klo | nord | 8 x f16 counts }
// K <= 63: 64 B = two sectors of one burst  { klo khi nord pad | 8 x f16 counts, 16 B pad }
//   klo/khi: canonical k-mer; (0,0) = empty, like the reference, whose all-A k-mer is kept in a side node
//            (DBGgraph.cpp:153-164).
//   nord   : ~ordinal of the earliest occurrence seen so far (0 = none yet); ordinal = read_index<<16 | j.
//            Only used to reproduce the reference's slot layout at export.
//   counts : occurrences per neighbour base, l lanes A,C,G,T then r lanes A,C,G,T, as IEEE half floats, so
//            that one occurrence is ONE fire-and-forget 128-bit vector reduction
//            (red.global.add.noftz.v4.f16x2 -> SASS REDG.E.ADD.F16x8) adding 1.0 to its two lanes: no CAS
//            loop, no return trip, one L2 request.  Halves count integers exactly up to 2048 and then stay
//            at 2048 (2048 + 1 rounds back to 2048), so a lane can never overflow; the reference's 8-bit
//            saturating lanes (kmerSet.cpp:56,341) are min(255, count) at export.
template <bool WIDE> struct NodeT;
template <> struct __align__(32) NodeT<false> { u64 klo; u64 nord; u64 c0; u64 c1; };
template <> struct __align__(64) NodeT<true> { u64 klo; u64 khi; u64 nord; u64 pad; u64 c0; u64 c1; u64 pad2[2]; };
static_assert(sizeof(NodeT<false>) == 32 && sizeof(NodeT<true>) == 64, "node sizes");
constexpr u32 HALF_ONE = 0x3C00u;     // 1.0
constexpr u32 HALF_255 = 0x5BF8u;     // 255.0; positive halves order like their bit patterns

struct TableView {
    void *nodes;      // n_local NodeT<WIDE> slots (shard range + overflow margin)
    u64 P;            // reference table size (find_next_prime)
    u64 M;            // floor(2^64 / P) for the Barrett reduction of hash % P
    u64 lo;           // first home slot owned by this shard
    u64 n_local;      // physical slots
    u64 *counters;    // [0] new nodes, [1] probe conflicts, [2] occurrences, [3] kmers_logged, [4] error flag, [5] reads
    u64 *polyA;       // 8 x u64 occurrence counts for the k-mer-0 side node: l lanes A,C,G,T then r lanes
    u64 guard_budget; // long-probe budget of the build so far (insert_probe): 2^20 + occurrences submitted
};

// [6] k-mer-0 link words (k_polyA_insert), [7] tile counter of k_insert_tuples, [8] 1024-slot crossings of long probes
enum { CNT_NEW = 0, CNT_CONFLICT = 1, CNT_OCC = 2, CNT_LOGGED = 3, CNT_ERROR = 4, CNT_READS = 5, CNT_GUARD = 8, CNT_N = 16 };

// ---- hashing ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 hash_code(u64 kmer)
{
    kmer += ~(kmer << 32);
    kmer ^= (kmer >> 22);
    kmer += ~(kmer << 13);
    kmer ^= (kmer >> 8);
    kmer += (kmer << 3);
    kmer ^= (kmer >> 15);
    kmer += ~(kmer << 27);
    kmer ^= (kmer >> 31);
    return kmer;
}

// 128-bit keys: no reference exists; reduces to hash_code(lo) when hi == 0.
__host__ __device__ __forceinline__ u64 hash_code_wide(u64 lo, u64 hi)
{
    u64 f = hash_code(hi) - 0x6A396CD39C352659ULL;   // hash_code(0) = 7654268697807496793
    f = (f << 32) | (f >> 32);
    return hash_code(lo ^ f);
}

__device__ __forceinline__ u64 mod_P(u64 h, u64 P, u64 M)
{
    u64 q = __umul64hi(h, M);
    u64 r = h - q * P;
    return r >= P ? r - P : r;
}

// ---- sequence helpers ---------------------------------------------------------------------------------
// 4 ASCII bases (little-endian word, byte 0 = first base) -> 8 bits, first base in the top 2 bits.
__device__ __forceinline__ u32 pack4(u32 w)
{
    u32 x = (w >> 1) & 0x03030303u;            // A0 C1 T2 G3 (and N,n -> 3)
    u32 y = x ^ ((x >> 1) & 0x01010101u);      // A0 C1 G2 T3
    u32 n = (w >> 3) & 0x01010101u;            // bit 3 is set only for N / n inside [ACGTNacgtn]
    y &= ~(n * 3u);                            // N,n -> 0 like alphabet[] (seqKmer.cpp:15,17)
    return (y * 0x40100401u) >> 24;
}

__device__ __forceinline__ u32 pack16(uint4 v)
{
    return (pack4(v.x) << 24) | (pack4(v.y) << 16) | (pack4(v.z) << 8) | pack4(v.w);
}

// 2-bit code of base q in a packed stream (16 bases per u32, first base in the top bits)
__device__ __forceinline__ u32 code_at(const u32 *pk, u32 q)
{
    return (pk[q >> 4] >> (30 - 2 * (q & 15))) & 3u;
}

// reverse complement of a right-aligned 2K-bit word (seqKmer.cpp:89-97 computes the same value)
__device__ __forceinline__ u64 revcomp64(u64 x, int K)
{
    x = __brevll(~x);
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    return x >> (64 - 2 * K);
}

// window of K <= 31 bases starting at base p of the packed stream
__device__ __forceinline__ u64 window64(const u32 *pk, u32 p, int K)
{
    u32 wi = p >> 4, sh = (p & 15) * 2;
    u32 w0 = pk[wi], w1 = pk[wi + 1], w2 = pk[wi + 2];
    u32 hi = __funnelshift_l(w1, w0, sh);
    u32 lo = __funnelshift_l(w2, w1, sh);
    return (((u64)hi << 32) | lo) >> (64 - 2 * K);
}

// 128-bit helpers for 31 < K <= 63 (value = hi:lo)
struct U128 { u64 lo, hi; };

__device__ __forceinline__ U128 shr128(U128 v, int s)   // 0 <= s < 128
{
    U128 r;
    if (s == 0) return v;
    if (s >= 64) { r.lo = v.hi >> (s - 64); r.hi = 0; }
    else { r.lo = (v.lo >> s) | (v.hi << (64 - s)); r.hi = v.hi >> s; }
    return r;
}

__device__ __forceinline__ U128 window128(const u32 *pk, u32 p, int K)
{
    u32 wi = p >> 4, sh = (p & 15) * 2;
    u32 w0 = pk[wi], w1 = pk[wi + 1], w2 = pk[wi + 2], w3 = pk[wi + 3], w4 = pk[wi + 4];
    u32 a = __funnelshift_l(w1, w0, sh), b = __funnelshift_l(w2, w1, sh);
    u32 c = __funnelshift_l(w3, w2, sh), d = __funnelshift_l(w4, w3, sh);
    U128 v; v.hi = ((u64)a << 32) | b; v.lo = ((u64)c << 32) | d;
    return shr128(v, 128 - 2 * K);
}

__device__ __forceinline__ U128 revcomp128(U128 v, int K)
{
    // reverse all 64 two-bit groups: swap halves, reverse each half
    u64 a = __brevll(~v.lo), b = __brevll(~v.hi);
    a = ((a >> 1) & 0x5555555555555555ULL) | ((a & 0x5555555555555555ULL) << 1);
    b = ((b >> 1) & 0x5555555555555555ULL) | ((b & 0x5555555555555555ULL) << 1);
    U128 r; r.hi = a; r.lo = b;
    return shr128(r, 128 - 2 * K);
}

// ---- memory helpers -------------------------------------------------------------------------------
// one 256-bit load (LDG.E.256 on sm_100): a whole 32-B sector in ONE L2 request
__device__ __forceinline__ void ld256_cg(const void *p, u64 &a, u64 &b, u64 &c, u64 &d)
{
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void ld256_cs(const void *p, u64 &a, u64 &b, u64 &c, u64 &d)
{
    asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

struct NodeRegs {
    u64 klo, khi, nord;
    u64 c0, c1;        // l lanes / r lanes, 4 halves each
};

// .cg: the table is only ever coherent at L2 (atomics live there)
__device__ __forceinline__ void load_node(const NodeT<false> *p, NodeRegs &n)
{
    ld256_cg(p, n.klo, n.nord, n.c0, n.c1);    // the whole node: one request
    n.khi = 0;
}
__device__ __forceinline__ void load_node(const NodeT<true> *p, NodeRegs &n)
{
    u64 pad;
    ld256_cg(p, n.klo, n.khi, n.nord, pad);
    ulonglong2 c = __ldcg(reinterpret_cast<const ulonglong2 *>(p) + 2);
    n.c0 = c.x; n.c1 = c.y;
}

// what the insert path keeps of a probed node: key, ordinal and the two lanes this occurrence would bump
struct ProbeRegs {
    u64 klo, khi, nord;
    u32 lanes;         // half bits of l lane lb | r lane rb << 16 (0 when the occurrence has no such neighbour)
};

template <bool WIDE>
__device__ __forceinline__ void probe_node(const NodeT<WIDE> *p, u32 lb, u32 rb, ProbeRegs &q)
{
    NodeRegs n;
    load_node(p, n);
    q.klo = n.klo; q.khi = n.khi; q.nord = n.nord;
    u32 l = lb < 4 ? ((u32)(n.c0 >> (16 * lb)) & 0xFFFFu) : 0xFFFFu;      // 0xFFFF: "nothing to add on this side"
    u32 r = rb < 4 ? ((u32)(n.c1 >> (16 * rb)) & 0xFFFFu) : 0xFFFFu;
    q.lanes = l | (r << 16);
}

// one occurrence: +1.0 on lane lb of the l counts and lane rb of the r counts (skipping lanes that are
// absent or already known to be >= 255), as ONE 128-bit vector reduction
template <bool WIDE>
__device__ __forceinline__ void bump_counts(NodeT<WIDE> *p, u32 lanes, u32 lb, u32 rb)
{
    u64 al = 0, ar = 0;
    if ((lanes & 0xFFFFu) < HALF_255) al = (u64)HALF_ONE << (16 * (lb & 3));
    if ((lanes >> 16) < HALF_255) ar = (u64)HALF_ONE << (16 * (rb & 3));
    if ((al | ar) == 0) return;
    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(&p->c0), "r"((u32)al), "r"((u32)(al >> 32)),
                 "r"((u32)ar), "r"((u32)(ar >> 32))
                 : "memory");
}

// the reference's link word from four half counts: lane of base b = bits (3-b)*8.., saturating at 255
// (kmerSet.cpp:56,341).  Two packed-half instructions per pair of lanes: clamp to 255.0, add 1024.0 -- in
// [1024, 2048) a half's ulp is 1, so the low byte of the sum IS the count -- then one byte permute gathers the four
// counts with lane A on top.
__device__ __forceinline__ u32 pack_link(u64 c)
{
    const __half2 cap = __floats2half2_rn(255.f, 255.f), magic = __floats2half2_rn(1024.f, 1024.f);
    u32 lo = (u32)c, hi = (u32)(c >> 32);                    // lo = lanes A,C ; hi = lanes G,T (16 bits each)
    __half2 a = __hadd2(__hmin2(*reinterpret_cast<__half2 *>(&lo), cap), magic);
    __half2 b = __hadd2(__hmin2(*reinterpret_cast<__half2 *>(&hi), cap), magic);
    return __byte_perm(*reinterpret_cast<u32 *>(&a), *reinterpret_cast<u32 *>(&b), 0x0246);   // A<<24 | C<<16 | G<<8 | T
}

__device__ __forceinline__ bool cas128(void *addr, u64 new_lo, u64 new_hi, u64 &old_lo, u64 &old_hi)
{
    // compare value is (0,0): claim an empty 128-bit key.  atom.cas.b128 needs sm_90+ (PTX ISA 8.3).
    u64 z = 0;
    asm volatile(
        "{\n\t"
        ".reg .b128 c, n, o;\n\t"
        "mov.b128 c, {%2, %2};\n\t"
        "mov.b128 n, {%3, %4};\n\t"
        "atom.global.cas.b128 o, [%5], c, n;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}\n"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(z), "l"(new_lo), "l"(new_hi), "l"(addr)
        : "memory");
    return (old_lo | old_hi) == 0;
}

// saturating +1 on the lane of `base` (0..3; 4 = no neighbour), kmerSet.cpp:56 / DBGgraph.cpp:185-193
__device__ __forceinline__ u32 lane_inc(u32 link, u32 base)
{
    if (base < 4) {
        u32 sh = 24 - 8 * base;
        if (((link >> sh) & 0xFFu) < 255u) link += 1u << sh;
    }
    return link;
}

// nul_flag / del_flag bit of slot idx inside a little-endian u32 view of the MSB-first byte bitmap
// (kmerSet.h:144-169: byte idx/8, bit 7 - idx%8)
__host__ __device__ __forceinline__ u32 flag_mask(u64 idx)
{
    return 1u << (8 * ((idx >> 3) & 3) + 7 - (idx & 7));
}

}  // namespace dbg
