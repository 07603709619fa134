/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the contig SEED INDEX of link_scaffold:
 * the k-mer -> (contig id, position, unique?, strand) hash that map_pair / map_reads build from the contigs and probe
 * with read k-mers.  Only tests/ may load it; the product (dbg_assembly_b200/csrc/seedidx.cu) never does.
 *
 * PINNED: tests/test_seedidx.py compares every function here with the reference's own code compiled in place
 * (oracle/_ref/ref_seed_driver = link_scaffold/{kmerSet,map_func,seqKmer,gzstream}.cpp + oracle/ref_seed_driver.cpp):
 * table size, count, every node and its slot, and the seeds get_align_seed finds for a batch of reads; golden copies of
 * those outputs are committed under tests/golden/ (seed_*.npz) for the box without /root/reference.
 *
 * Citations are link_scaffold/<file>:<line> under /root/reference.  Sequential like the original.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dbg_oracle.h"

/* kmerSet.h:53-60 -- struct { uint64_t kmer; uint64_t id:32, pos:30, freq:1, direct:1; } with GCC's bit-field layout
 * (first field in the low bits): value = id | pos << 32 | freq << 62 | direct << 63 */
typedef struct { uint64_t kmer, value; } seed_node;

struct orc_seed {
    int K;
    uint64_t size, count, conflict, max;
    float load_factor;
    seed_node *array;
    uint8_t *nul;
    int would_enlarge;
};

static int seed_is_null(const uint8_t *nul, uint64_t idx) { return 1 - ((nul[idx / 8] >> (7 - idx % 8)) & 1u); }   /* kmerSet.h:115-118 */
static void seed_set_fill(uint8_t *nul, uint64_t idx) { nul[idx / 8] |= (uint8_t)(128u >> (idx % 8)); }              /* kmerSet.h:121-124, BitOrVal */

/* init_kmerset, kmerSet.cpp:82-107.  max = (uint64_t)(size * load_factor) is a FLOAT product (size converted to float). */
orc_seed *orc_seed_create(int K, uint64_t init_size, float load_factor)
{
    orc_seed *s = (orc_seed *)calloc(1, sizeof(orc_seed));
    if (init_size < 3) init_size = 3;
    else init_size = orc_find_next_prime(init_size);            /* same is_prime / find_next_prime as DBG_contig (kmerSet.cpp:56-79) */
    s->K = K;
    s->size = init_size;
    if (load_factor <= 0) load_factor = 0.25f;
    else if (load_factor >= 1) load_factor = 0.75f;
    s->load_factor = load_factor;
    s->max = (uint64_t)((float)s->size * load_factor);
    s->array = (seed_node *)calloc(s->size, sizeof(seed_node));   /* the reference leaves it uninitialised; empty slots are never read */
    s->nul = (uint8_t *)calloc(s->size / 8 + 1, 1);
    return s;
}

void orc_seed_destroy(orc_seed *s)
{
    if (!s) return;
    free(s->array); free(s->nul); free(s);
}

/* add_kmerset, kmerSet.cpp:168-210.  The reference enlarges when count >= max BEFORE looking at the k-mer; map_pair and
 * map_reads size the table at 3 x the contig length with load factor 0.5 (map_pair.cpp:122-124), so that never happens
 * there; the oracle (like the product) reports it instead of restating enlarge_kmerset. */
static void seed_add(orc_seed *s, uint64_t kmer, uint32_t id, uint32_t pos, int direct)
{
    if (s->count >= s->max) { s->would_enlarge = 1; return; }
    uint64_t hc = orc_hash_code(kmer) % s->size;
    for (;;) {
        if (seed_is_null(s->nul, hc)) {
            s->array[hc].kmer = kmer;
            s->array[hc].value = (uint64_t)id | ((uint64_t)(pos & 0x3FFFFFFFu) << 32) | (1ULL << 62) | ((uint64_t)(direct & 1) << 63);
            seed_set_fill(s->nul, hc);
            s->count++;
            return;
        }
        if (s->array[hc].kmer == kmer) { s->array[hc].value &= ~(1ULL << 62); return; }     /* freq = 0: seen again */
        s->conflict++;
        hc = (hc + 1 == s->size) ? 0 : hc + 1;
    }
}

/* chop_contig_to_kmerset, map_func.cpp:119-172, for the sequences seqs[offs[i] .. offs[i+1]), ids id0 + i.
 * scaffold_to_contig (map_func.cpp:303-326) cuts a sequence at runs of 'N' (upper case only; 'n' is base code 0 like in
 * alphabet[], seqKmer.cpp:15-24).  Blocks shorter than K: the reference's loop bound `contig_str.size()-KmerSize+1` is
 * unsigned and wraps -- undefined behaviour there; skipped here (and in the product).  Returns 0, or -1 if the reference
 * would have enlarged its table. */
int orc_seed_add_contigs(orc_seed *s, const char *seqs, const uint64_t *offs, uint64_t n, uint64_t id0)
{
    const int K = s->K;
    const uint64_t mask = (K == 32) ? ~0ULL : ((1ULL << (2 * K)) - 1);     /* KmerHeadMaskVal, :121 */
    uint64_t rc_or[4];                                                      /* KmerRCOrVal, :123-126 */
    rc_or[3] = 0; rc_or[1] = 1ULL << (2 * K - 1); rc_or[2] = 1ULL << (2 * K - 2); rc_or[0] = rc_or[1] + rc_or[2];
    for (uint64_t c = 0; c < n; c++) {
        const char *sq = seqs + offs[c];
        const uint64_t len = offs[c + 1] - offs[c];
        uint64_t i = 0;
        while (i < len) {
            while (i < len && sq[i] == 'N') i++;
            const uint64_t start = i;
            while (i < len && sq[i] != 'N') i++;
            const uint64_t blen = i - start;
            if (blen < (uint64_t)K) continue;
            uint64_t kbit = 0, rc = 0;
            for (uint64_t j = 0; j + K <= blen; j++) {
                if (j == 0) { kbit = orc_seq2bit(sq + start, K); rc = orc_rev_com_kbit(kbit, K); }
                else {
                    const uint64_t b = (uint64_t)orc_base_code((unsigned char)sq[start + j + K - 1]);
                    kbit = ((kbit << 2) | b) & mask;
                    rc = (rc >> 2) | rc_or[b & 3];
                }
                if (kbit < rc) seed_add(s, kbit, (uint32_t)(id0 + c), (uint32_t)(start + j), 1);     /* :154-162: strict <, tie -> direct 0 */
                else seed_add(s, rc, (uint32_t)(id0 + c), (uint32_t)(start + j), 0);
                if (s->would_enlarge) return -1;
            }
        }
    }
    return 0;
}

uint64_t orc_seed_size(const orc_seed *s) { return s->size; }
uint64_t orc_seed_count(const orc_seed *s) { return s->count; }
uint64_t orc_seed_max(const orc_seed *s) { return s->max; }
uint64_t orc_seed_conflict(const orc_seed *s) { return s->conflict; }
const void *orc_seed_array(const orc_seed *s) { return s->array; }
const uint8_t *orc_seed_nul_flag(const orc_seed *s) { return s->nul; }

/* exist_kmerset, kmerSet.cpp:216-238 (del_flag is never set on this table) */
static uint64_t seed_find(const orc_seed *s, uint64_t kmer)
{
    uint64_t hc = orc_hash_code(kmer) % s->size;
    for (;;) {
        if (seed_is_null(s->nul, hc)) return s->size;
        if (s->array[hc].kmer == kmer) return hc;
        hc = (hc + 1 == s->size) ? 0 : hc + 1;
    }
}

/* get_align_seed, map_func.cpp:181-237.  out = {contig_id_index, seed_contig_start, seed_contig_end, seed_read_start,
 * seed_read_end, direct ('F' / 'R' / 'N')}; the first five stay -1 when no seed is found. */
void orc_seed_align(const orc_seed *s, const char *read, int len, int search_start, int search_end, int seed_kmer_num, int32_t out[6])
{
    const int K = s->K;
    out[0] = out[1] = out[2] = out[3] = out[4] = -1; out[5] = 'N';
    (void)len;
    for (int i = search_start - 1; i <= search_end - K - seed_kmer_num; i++) {
        uint64_t kbit = orc_seq2bit(read + i, K), rc = orc_rev_com_kbit(kbit, K);
        uint64_t kmer = kbit < rc ? kbit : rc;
        int direct = kbit < rc ? 1 : 0;
        uint64_t idx = seed_find(s, kmer);
        if (idx == s->size || !((s->array[idx].value >> 62) & 1)) continue;
        uint64_t kbit2 = orc_seq2bit(read + i + seed_kmer_num, K), rc2 = orc_rev_com_kbit(kbit2, K);
        uint64_t kmer2 = kbit2 < rc2 ? kbit2 : rc2;
        uint64_t idx2 = seed_find(s, kmer2);
        if (idx2 == s->size) continue;
        const uint64_t v = s->array[idx].value, v2 = s->array[idx2].value;
        const int32_t pos = (int32_t)((v >> 32) & 0x3FFFFFFFu), pos2 = (int32_t)((v2 >> 32) & 0x3FFFFFFFu);
        if (((v2 >> 62) & 1) && (uint32_t)v2 == (uint32_t)v && abs(pos2 - pos) == seed_kmer_num) {
            if (direct == (int)(v >> 63)) { out[1] = pos + 1; out[2] = pos2 + K; out[5] = 'F'; }
            else { out[1] = pos2 + 1; out[2] = pos + K; out[5] = 'R'; }
            out[3] = i + 1; out[4] = i + seed_kmer_num + K;
            out[0] = (int32_t)(uint32_t)v;
            return;
        }
    }
}
