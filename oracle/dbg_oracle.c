/* TEST INFRASTRUCTURE ONLY -- see dbg_oracle.h.  Plain-C, sequential restatement of the reference's
 * De Bruijn graph build (equals the reference run with -t 1).  Paths cited are relative to
 * /root/reference/DBG_contig/.  Nothing here is reachable from the product path. */
#include "dbg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- a-1: seqKmer.cpp:9-19 ------- */
int orc_base_code(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': case 'N': case 'n': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4; /* outside the domain: UB in the reference (DBGgraph.cpp:71-73) */
    }
}

/* ---------------------------------------------------------------- a-2: seqKmer.cpp:34-41 ------ */
uint64_t orc_seq2bit(const char *seq, int k)
{
    uint64_t kbit = 0;
    for (int i = 0; i < k; i++) kbit = (kbit << 2) | (uint64_t)orc_base_code((unsigned char)seq[i]);
    return kbit;
}

/* ---------------------------------------------------------------- a-3: seqKmer.cpp:89-97 ------ */
uint64_t orc_rev_com_kbit(uint64_t kbit, int k)
{
    kbit = ~kbit;
    kbit = ((kbit & 0x3333333333333333ULL) << 2) | ((kbit & 0xCCCCCCCCCCCCCCCCULL) >> 2);
    kbit = ((kbit & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((kbit & 0xF0F0F0F0F0F0F0F0ULL) >> 4);
    kbit = ((kbit & 0x00FF00FF00FF00FFULL) << 8) | ((kbit & 0xFF00FF00FF00FF00ULL) >> 8);
    kbit = ((kbit & 0x0000FFFF0000FFFFULL) << 16) | ((kbit & 0xFFFF0000FFFF0000ULL) >> 16);
    kbit = ((kbit & 0x00000000FFFFFFFFULL) << 32) | ((kbit & 0xFFFFFFFF00000000ULL) >> 32);
    return kbit >> (64 - (k << 1));
}

/* 128-bit twin of the above: complement, reverse the 2-bit groups, right-align to 2K bits */
static u128 rev_com_wide(u128 kbit, int k)
{
    uint64_t lo = (uint64_t)kbit, hi = (uint64_t)(kbit >> 64);
    /* reversing 64 groups of 2 bits = swap halves, reverse each half's 32 groups */
    uint64_t rlo = orc_rev_com_kbit(hi, 32); /* k=32 -> shift by 0 */
    uint64_t rhi = orc_rev_com_kbit(lo, 32);
    u128 r = ((u128)rhi << 64) | rlo;
    return r >> (128 - 2 * k);
}

void orc_rev_com_wide(uint64_t lo, uint64_t hi, int k, uint64_t *rlo, uint64_t *rhi)
{
    u128 r = rev_com_wide(((u128)hi << 64) | lo, k);
    *rlo = (uint64_t)r;
    *rhi = (uint64_t)(r >> 64);
}

/* ---------------------------------------------------------------- a-8: kmerSet.h:105-116 ------ */
uint64_t orc_hash_code(uint64_t kmer)
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

/* No reference exists for K > 31.  Fold the high word into the low one so that hi == 0 reproduces
 * hash_code(lo) exactly (the only available pin). */
uint64_t orc_hash_code_wide(uint64_t lo, uint64_t hi)
{
    uint64_t f = orc_hash_code(hi) - orc_hash_code(0);
    f = (f << 32) | (f >> 32);
    return orc_hash_code(lo ^ f);
}

/* ---------------------------------------------------------------- a-9: kmerSet.cpp:72-95 ------ */
int orc_is_prime(uint64_t num)
{
    uint64_t i, max;
    if (num < 4) return 1;
    if (num % 2 == 0) return 0;
    /* the reference calls sqrt((float)num): float overload, float rounding, and then `i < max`
     * (exclusive) -- so squares of primes etc. can pass.  Restated literally. */
    max = (uint64_t)sqrtf((float)num);
    for (i = 3; i < max; i += 2)
        if (num % i == 0) return 0;
    return 1;
}

uint64_t orc_find_next_prime(uint64_t num)
{
    if (num % 2 == 0) num++;
    for (;;) {
        if (orc_is_prime(num)) return num;
        num += 2;
    }
}

/* ---------------------------------------------------------------- a-4..a-6: DBGgraph.cpp:49-104 */
int orc64_parse_read(const char *read, uint64_t len, int K, int max_read_len,
                     uint64_t *kmer, uint8_t *left, uint8_t *right)
{
    if (len < (uint64_t)K) return 0;                                   /* :51-53 */
    /* DBGgraph.cpp:371-376 */
    uint64_t mask = (K == 32) ? ~0ULL : ((1ULL << (2 * K)) - 1);
    uint64_t rc_or[4];
    rc_or[3] = 0;
    rc_or[1] = 1ULL << (2 * K - 1);
    rc_or[2] = 1ULL << (2 * K - 2);
    rc_or[0] = rc_or[1] + rc_or[2];

    int readlen = (len > (uint64_t)max_read_len) ? max_read_len : (int)len;   /* :63 */
    uint64_t kbit = 0, rc = 0;
    int n = 0;
    for (int j = 0; j < readlen - K + 1; j++) {
        if (j == 0) {                                                  /* :66-69 */
            kbit = orc_seq2bit(read, K);
            rc = orc_rev_com_kbit(kbit, K);
        } else {                                                       /* :71-73 */
            int b = orc_base_code((unsigned char)read[j + K - 1]);
            kbit = ((kbit << 2) | (uint64_t)b) & mask;
            rc = (rc >> 2) | rc_or[b];
        }
        uint8_t lb = 4, rb = 4;
        uint64_t km;
        if (kbit <= rc) {                                              /* :80-83 */
            km = kbit;
            if (j > 0) lb = (uint8_t)orc_base_code((unsigned char)read[j - 1]);
            if (j < readlen - K) rb = (uint8_t)orc_base_code((unsigned char)read[j + K]);
        } else {                                                       /* :85-89 */
            km = rc;
            if (j > 0) rb = (uint8_t)(3 - orc_base_code((unsigned char)read[j - 1]));
            if (j < readlen - K) lb = (uint8_t)(3 - orc_base_code((unsigned char)read[j + K]));
        }
        kmer[n] = km; left[n] = lb; right[n] = rb; n++;                /* :93-96 */
    }
    return n;
}

int orc128_parse_read(const char *read, uint64_t len, int K, int max_read_len,
                      uint64_t *kmer_lo, uint64_t *kmer_hi, uint8_t *left, uint8_t *right)
{
    if (len < (uint64_t)K) return 0;
    u128 mask = (K == 64) ? ~(u128)0 : (((u128)1 << (2 * K)) - 1);
    int readlen = (len > (uint64_t)max_read_len) ? max_read_len : (int)len;
    u128 kbit = 0, rc = 0;
    int n = 0;
    for (int j = 0; j < readlen - K + 1; j++) {
        if (j == 0) {
            for (int i = 0; i < K; i++) kbit = (kbit << 2) | (u128)orc_base_code((unsigned char)read[i]);
            rc = rev_com_wide(kbit, K);
        } else {
            int b = orc_base_code((unsigned char)read[j + K - 1]);
            kbit = ((kbit << 2) | (u128)b) & mask;
            rc = (rc >> 2) | ((u128)(3 - b) << (2 * (K - 1)));
        }
        uint8_t lb = 4, rb = 4;
        u128 km;
        if (kbit <= rc) {
            km = kbit;
            if (j > 0) lb = (uint8_t)orc_base_code((unsigned char)read[j - 1]);
            if (j < readlen - K) rb = (uint8_t)orc_base_code((unsigned char)read[j + K]);
        } else {
            km = rc;
            if (j > 0) rb = (uint8_t)(3 - orc_base_code((unsigned char)read[j - 1]));
            if (j < readlen - K) lb = (uint8_t)(3 - orc_base_code((unsigned char)read[j + K]));
        }
        kmer_lo[n] = (uint64_t)km; kmer_hi[n] = (uint64_t)(km >> 64);
        left[n] = lb; right[n] = rb; n++;
    }
    return n;
}

/* ---------------------------------------------------------------- a-7: kmerSet.h:70-99 -------- */
struct orc_set {
    int K, max_read_len, wide;
    uint64_t buffer_reads, max_double_times, double_times;
    uint64_t size, count, count_conflict, max;
    float load_factor;
    uint64_t *klo, *khi;           /* khi only when wide */
    uint32_t *l, *r;
    uint8_t *nul_flag;
    uint32_t polyA_l, polyA_r;     /* side node for k-mer 0, DBGgraph.cpp:153-164,399-402 */
    uint64_t total_reads, kmers_logged, occurrences;
    /* per-read scratch */
    uint64_t *s_lo, *s_hi; uint8_t *s_left, *s_right;
    void *export_buf;
};

static const uint32_t BitAddVal[4] = {0x1000000u, 0x10000u, 0x100u, 0x1u};  /* kmerSet.cpp:56 */

static int is_null(const uint8_t *f, uint64_t i) { return 1 - ((f[i / 8] >> (7 - i % 8)) & 1u); }   /* kmerSet.h:144 */
static void set_fill(uint8_t *f, uint64_t i) { f[i / 8] |= (uint8_t)(128u >> (i % 8)); }            /* kmerSet.h:151 */
static uint8_t lane(uint32_t link, int base) { return (link >> ((3 - base) * 8)) & 0xFFu; }         /* kmerSet.cpp:341 */

static uint64_t hash_of(const orc_set *s, uint64_t lo, uint64_t hi)
{
    return s->wide ? orc_hash_code_wide(lo, hi) : orc_hash_code(lo);
}

/* init_kmerset_parallel, kmerSet.cpp:98-127 */
orc_set *orc_create(int K, int max_read_len, uint64_t init_size, float load_factor,
                    uint64_t max_double_times, uint64_t buffer_reads, int wide)
{
    if (K < 1 || (!wide && K > 31) || (wide && K > 63) || max_read_len < 1 || buffer_reads < 1) return NULL;
    orc_set *s = (orc_set *)calloc(1, sizeof(orc_set));
    s->K = K; s->max_read_len = max_read_len; s->wide = wide;
    s->buffer_reads = buffer_reads; s->max_double_times = max_double_times;
    if (init_size < 3) init_size = 3; else init_size = orc_find_next_prime(init_size);
    s->size = init_size;
    if (load_factor <= 0) load_factor = 0.25f; else if (load_factor >= 1) load_factor = 0.75f;
    s->load_factor = load_factor;
    s->max = (uint64_t)(s->size * load_factor);          /* uint64 * float -> float, as in the reference */
    s->klo = (uint64_t *)calloc(s->size, 8);
    s->khi = wide ? (uint64_t *)calloc(s->size, 8) : NULL;
    s->l = (uint32_t *)calloc(s->size, 4);
    s->r = (uint32_t *)calloc(s->size, 4);
    s->nul_flag = (uint8_t *)calloc(s->size / 8 + 1, 1);
    int kn = max_read_len - K + 1; if (kn < 1) kn = 1;
    s->s_lo = (uint64_t *)malloc(8 * (size_t)kn);
    s->s_hi = (uint64_t *)malloc(8 * (size_t)kn);
    s->s_left = (uint8_t *)malloc((size_t)kn);
    s->s_right = (uint8_t *)malloc((size_t)kn);
    return s;
}

void orc_destroy(orc_set *s)
{
    if (!s) return;
    free(s->klo); free(s->khi); free(s->l); free(s->r); free(s->nul_flag);
    free(s->s_lo); free(s->s_hi); free(s->s_left); free(s->s_right); free(s->export_buf);
    free(s);
}

/* enlarge_kmerset_parallel, kmerSet.cpp:132-189 (in-place x2 rehash in old-slot order, swap chain) */
static void enlarge(orc_set *s, uint64_t num)
{
    uint64_t old_size = s->size, new_size = s->size;
    do { new_size = orc_find_next_prime(new_size * 2); } while (new_size * s->load_factor < s->count + num);
    s->size = new_size;
    s->klo = (uint64_t *)realloc(s->klo, new_size * 8);
    memset(s->klo + old_size, 0, (new_size - old_size) * 8);
    if (s->wide) { s->khi = (uint64_t *)realloc(s->khi, new_size * 8); memset(s->khi + old_size, 0, (new_size - old_size) * 8); }
    s->l = (uint32_t *)realloc(s->l, new_size * 4); memset(s->l + old_size, 0, (new_size - old_size) * 4);
    s->r = (uint32_t *)realloc(s->r, new_size * 4); memset(s->r + old_size, 0, (new_size - old_size) * 4);
    s->max = (uint64_t)(new_size * s->load_factor);

    uint8_t *old_nul = s->nul_flag;
    uint8_t *old_del = (uint8_t *)calloc(old_size / 8 + 1, 1);   /* build-time del_flag is all zero */
    s->nul_flag = (uint8_t *)calloc(new_size / 8 + 1, 1);

    for (uint64_t i = 0; i < old_size; i++) {
        if (is_null(old_nul, i) || !is_null(old_del, i)) continue;
        uint64_t t_lo = s->klo[i], t_hi = s->wide ? s->khi[i] : 0; uint32_t t_l = s->l[i], t_r = s->r[i];
        s->klo[i] = 0; if (s->wide) s->khi[i] = 0; s->l[i] = 0; s->r[i] = 0;
        set_fill(old_del, i);
        for (;;) {
            uint64_t hc = hash_of(s, t_lo, t_hi) % s->size;
            while (!is_null(s->nul_flag, hc)) hc = (hc + 1) % s->size;
            set_fill(s->nul_flag, hc);
            if (hc < old_size && !is_null(old_nul, hc) && is_null(old_del, hc)) {
                uint64_t x_lo = s->klo[hc], x_hi = s->wide ? s->khi[hc] : 0; uint32_t x_l = s->l[hc], x_r = s->r[hc];
                s->klo[hc] = t_lo; if (s->wide) s->khi[hc] = t_hi; s->l[hc] = t_l; s->r[hc] = t_r;
                t_lo = x_lo; t_hi = x_hi; t_l = x_l; t_r = x_r;
                set_fill(old_del, hc);
            } else {
                s->klo[hc] = t_lo; if (s->wide) s->khi[hc] = t_hi; s->l[hc] = t_l; s->r[hc] = t_r;
                break;
            }
        }
    }
    free(old_nul); free(old_del);
}

/* thread_updatekmers for one occurrence, DBGgraph.cpp:141-205 (threadNum = 1) */
static void update_one(orc_set *s, uint64_t lo, uint64_t hi, uint8_t lb, uint8_t rb)
{
    s->occurrences++;
    if (lo == 0 && hi == 0) {                                          /* :153-164 */
        if (lb != 4 && lane(s->polyA_l, lb) < 255) s->polyA_l += BitAddVal[lb];
        if (rb != 4 && lane(s->polyA_r, rb) < 255) s->polyA_r += BitAddVal[rb];
        return;
    }
    uint64_t hc = hash_of(s, lo, hi) % s->size;                        /* :167 */
    for (;;) {
        int empty = (s->klo[hc] == 0) && (!s->wide || s->khi[hc] == 0);
        if (empty) {                                                   /* :174-182 */
            s->klo[hc] = lo; if (s->wide) s->khi[hc] = hi;
            s->l[hc] = (lb != 4) ? BitAddVal[lb] : 0;
            s->r[hc] = (rb != 4) ? BitAddVal[rb] : 0;
            set_fill(s->nul_flag, hc);
            s->count++;
            return;
        }
        if (s->klo[hc] == lo && (!s->wide || s->khi[hc] == hi)) {      /* :185-196 */
            if (lb != 4 && lane(s->l[hc], lb) < 255) s->l[hc] += BitAddVal[lb];
            if (rb != 4 && lane(s->r[hc], rb) < 255) s->r[hc] += BitAddVal[rb];
            return;
        }
        s->count_conflict++;                                           /* :201-204 */
        hc = (hc + 1 == s->size) ? 0 : hc + 1;
    }
}

/* parse_one_reads_file, DBGgraph.cpp:217-359 */
uint64_t orc_add_file(orc_set *s, const char *bases, const uint64_t *offs, uint64_t n_reads)
{
    uint64_t next = 0;
    for (;;) {
        uint64_t nblk = n_reads - next; if (nblk > s->buffer_reads) nblk = s->buffer_reads;
        for (uint64_t i = next; i < next + nblk; i++) {
            const char *rd = bases + offs[i]; uint64_t len = offs[i + 1] - offs[i];
            int n;
            if (s->wide) n = orc128_parse_read(rd, len, s->K, s->max_read_len, s->s_lo, s->s_hi, s->s_left, s->s_right);
            else         n = orc64_parse_read(rd, len, s->K, s->max_read_len, s->s_lo, s->s_left, s->s_right);
            if (len >= (uint64_t)s->K) s->kmers_logged += len - (uint64_t)s->K + 1;  /* :101 untrimmed */
            for (int j = 0; j < n; j++) update_one(s, s->s_lo[j], s->wide ? s->s_hi[j] : 0, s->s_left[j], s->s_right[j]);
        }
        next += nblk; s->total_reads += nblk;
        if (nblk < s->buffer_reads) break;                             /* :329-331 final block: no grow check */
        if (s->count > s->max) {                                       /* :337-351 */
            if (s->double_times < s->max_double_times) { enlarge(s, 1); s->double_times++; }
            else break;   /* "Memory reach the maximum allowed": rest of this file ignored */
        }
    }
    return next;
}

/* add_node_to_kmerset(kset, PolyA), kmerSet.cpp:253-273 / DBGgraph.cpp:418 */
void orc_finish(orc_set *s)
{
    uint64_t hc = hash_of(s, 0, 0) % s->size;
    for (;;) {
        if (is_null(s->nul_flag, hc)) {
            s->klo[hc] = 0; if (s->wide) s->khi[hc] = 0;
            s->l[hc] = s->polyA_l; s->r[hc] = s->polyA_r;
            set_fill(s->nul_flag, hc);
            s->count++;
            return;
        }
        s->count_conflict++;
        hc = (hc + 1 == s->size) ? 0 : hc + 1;
    }
}

uint64_t orc_size(const orc_set *s) { return s->size; }
uint64_t orc_count(const orc_set *s) { return s->count; }
uint64_t orc_conflict(const orc_set *s) { return s->count_conflict; }
uint64_t orc_max(const orc_set *s) { return s->max; }
uint64_t orc_doublings(const orc_set *s) { return s->double_times; }
uint64_t orc_total_reads(const orc_set *s) { return s->total_reads; }
uint64_t orc_kmers_logged(const orc_set *s) { return s->kmers_logged; }
uint64_t orc_occurrences(const orc_set *s) { return s->occurrences; }
int orc_node_bytes(const orc_set *s) { return s->wide ? 32 : 16; }
const uint8_t *orc_nul_flag(const orc_set *s) { return s->nul_flag; }

const void *orc_array(const orc_set *s)
{
    orc_set *m = (orc_set *)s;
    free(m->export_buf);
    if (!s->wide) {
        struct n16 { uint64_t kmer; uint32_t l, r; } *a = (struct n16 *)malloc(s->size * 16);
        for (uint64_t i = 0; i < s->size; i++) { a[i].kmer = s->klo[i]; a[i].l = s->l[i]; a[i].r = s->r[i]; }
        m->export_buf = a;
    } else {
        struct n32 { uint64_t lo, hi; uint32_t l, r; uint64_t z; } *a = (struct n32 *)malloc(s->size * 32);
        for (uint64_t i = 0; i < s->size; i++) { a[i].lo = s->klo[i]; a[i].hi = s->khi[i]; a[i].l = s->l[i]; a[i].r = s->r[i]; a[i].z = 0; }
        m->export_buf = a;
    }
    return m->export_buf;
}

uint64_t orc_dump(const orc_set *s, uint64_t *slot, uint64_t *kmer_lo, uint64_t *kmer_hi, uint32_t *l, uint32_t *r)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i < s->size; i++) {
        if (is_null(s->nul_flag, i)) continue;
        if (slot) slot[n] = i;
        if (kmer_lo) kmer_lo[n] = s->klo[i];
        if (kmer_hi) kmer_hi[n] = s->wide ? s->khi[i] : 0;
        if (l) l[n] = s->l[i];
        if (r) r[n] = s->r[i];
        n++;
    }
    return n;
}

/* ---------------------------------------------------------------- a-13: contig.cpp:107-205 ---- */
void orc_calculate_kmer_links(const orc_set *s, int freq_cutoff, uint8_t *klink, uint8_t *del_flag,
                              int64_t *depth_stat, uint64_t *tips, uint64_t *n_tips,
                              uint64_t *branches, uint64_t *n_branches, int64_t *stats)
{
    int64_t total = 0, deleted = 0, linear = 0;
    uint64_t nt = 0, nb = 0;
    for (int i = 0; i < 256; i++) depth_stat[i] = 0;
    for (uint64_t i = 0; i < s->size; i++) {
        if (is_null(s->nul_flag, i)) continue;
        int l_num = 0, l_base = 0, r_num = 0, r_base = 0;
        int max_depth = 0;
        for (int j = 0; j < 4; j++) {                                 /* :128-143 */
            int d = lane(s->l[i], j);
            depth_stat[d]++;
            if (d > freq_cutoff) {
                if (l_num < 3) l_num++;
                if (max_depth < d) { max_depth = d; l_base = j; }
            }
        }
        max_depth = 0;
        for (int j = 0; j < 4; j++) {                                 /* :147-162 */
            int d = lane(s->r[i], j);
            depth_stat[d]++;
            if (d > freq_cutoff) {
                if (r_num < 3) r_num++;
                if (max_depth < d) { max_depth = d; r_base = j; }
            }
        }
        total++;
        uint8_t b0 = (uint8_t)(l_num | (l_base << 2) | (r_num << 4) | (r_base << 6));
        uint8_t b1 = 0;
        if (l_num == 0 && r_num == 0) { del_flag[i / 8] |= (uint8_t)(128u >> (i % 8)); deleted++; }   /* :166-169 */
        if (l_num == 1 && r_num == 1) { b1 |= 1; linear++; }                                           /* :171-174 */
        if (l_num + r_num == 1) tips[nt++] = i;                                                        /* :175-177 */
        if (l_num > 1 || r_num > 1) branches[nb++] = i;                                                /* :178-180 */
        klink[2 * i] = b0; klink[2 * i + 1] = b1;
    }
    *n_tips = nt; *n_branches = nb;
    stats[0] = total; stats[1] = deleted; stats[2] = linear;
}

/* ---------------------------------------------------------------- a-14 / a-15: K-mer frequency table -----------
 * PARITY UNPINNED: the producer (`kmerfreq`, fanagislab/kmerfreq, unpinned) is not in the reference tree.  This
 * restates what its consumers and artefacts imply: canonical k-mer of EVERY read position (N counts
 * as A, correct_error/seqKmer.cpp alphabet[]), direct index = the canonical 2K-bit value.  The in-tree mini builder
 * construct_ref_kmer_table (correct_error/simulate_lowfreq_kmer.cpp:189-260) walks positions the same way. */
void orc_kfreq_count(const char *bases, const uint64_t *offs, uint64_t n_reads, int K, uint32_t *counts)
{
    uint64_t mask = (1ULL << (2 * K)) - 1;
    for (uint64_t i = 0; i < n_reads; i++) {
        const char *rd = bases + offs[i];
        uint64_t len = offs[i + 1] - offs[i];
        if (len < (uint64_t)K) continue;
        uint64_t kbit = 0, rc = 0;
        for (uint64_t j = 0; j + K <= len; j++) {
            if (j == 0) { kbit = orc_seq2bit(rd, K); rc = orc_rev_com_kbit(kbit, K); }
            else {
                uint64_t b = (uint64_t)orc_base_code((unsigned char)rd[j + K - 1]);
                kbit = ((kbit << 2) | b) & mask;
                rc = (rc >> 2) | ((3 - b) << (2 * (K - 1)));
            }
            uint64_t km = kbit <= rc ? kbit : rc;
            if (counts[km] != 0xFFFFFFFFu) counts[km]++;
        }
    }
}
