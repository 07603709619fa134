// dbg_mock.cpp -- TEST INFRASTRUCTURE: a CPU stand-in for the handful of libdbgb200 entry points that
// integration/DBGgraph_b200.cpp calls, built on the oracle (oracle/liboracle.so).  It lets tests/test_frontend_cpu.py
// run the reference's front end + the binding end to end WITHOUT a GPU, so that the host-side logic of the binding
// (reader threads, block submission, device-table retry, growth replay, KmerSet hand-over) is exercised on the CPU
// box.  It is linked IN FRONT of the real library (-ldbgmock -ldbgb200): everything it does not define
// (dbg_replay_growth, dbg_find_next_prime, dbg_strerror ...) is the shipped code.  Never part of the product.
//
// Behaviour mirrored from the real library: the table never grows (one oracle block per submit with an unlimited block
// size, so the oracle's own grow check never fires); DBG_ERR_TABLE_FULL at finalize / get_stats when the nodes do not
// fit the table; dbg_dump_shard returns every node but the k-mer-0 one with its first-occurrence ordinal.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "dbg_b200.h"
#include "dbg_oracle.h"

struct dbg_ctx {
    dbg_params prm;
    orc_set *set = nullptr;
    uint64_t P = 0;
    std::unordered_map<uint64_t, uint64_t> first;   // canonical k-mer -> ordinal of its first occurrence
    uint64_t next_read = 0;
    bool full = false, finished = false;
    std::vector<uint64_t> klo; std::vector<uint8_t> lb, rb;
};

extern "C" {

int dbg_host_alloc(void **p, uint64_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? DBG_OK : DBG_ERR_NOMEM; }
int dbg_host_free(void *p) { free(p); return DBG_OK; }

int dbg_create(dbg_ctx **out, const dbg_params *p)
{
    dbg_ctx *c = new dbg_ctx();
    c->prm = *p;
    c->set = orc_create(p->K, p->max_read_len, p->init_slots, p->load_factor, 0, ~0ull, 0);
    c->P = orc_size(c->set);
    c->klo.resize(70000); c->lb.resize(70000); c->rb.resize(70000);
    *out = c;
    return DBG_OK;
}

void dbg_destroy(dbg_ctx *c)
{
    if (!c) return;
    orc_destroy(c->set);
    delete c;
}

int dbg_submit_reads(dbg_ctx *c, const char *bases, const uint64_t *offs, uint64_t n_reads)
{
    // first occurrences (what the device tracks as ordinals), and the "does it still fit" rule of the device table
    for (uint64_t i = 0; i < n_reads && !c->full; i++) {
        const uint64_t len = offs[i + 1] - offs[i];
        int n = orc64_parse_read(bases + offs[i], len, c->prm.K, c->prm.max_read_len, c->klo.data(), c->lb.data(), c->rb.data());
        for (int j = 0; j < n; j++)
            if (c->klo[j] != 0 && c->first.emplace(c->klo[j], ((c->next_read + i) << 16) | (uint64_t)j).second && c->first.size() + 1 > c->P)
                c->full = true;
    }
    c->next_read += n_reads;
    if (!c->full) orc_add_file(c->set, bases, offs, n_reads);     // one block (never a grow check: block size unlimited)
    return DBG_OK;                                                 // like the device: the overflow shows at the next stats read
}

static void fill(dbg_ctx *c, dbg_stats *st)
{
    memset(st, 0, sizeof(*st));
    st->array_size = orc_size(c->set); st->max_cutoff = orc_max(c->set); st->count = orc_count(c->set);
    st->conflict = orc_conflict(c->set); st->reads = c->next_read; st->kmers_logged = orc_kmers_logged(c->set);
    st->occurrences = orc_occurrences(c->set);
    float lf = c->prm.load_factor; if (lf <= 0) lf = 0.25f; else if (lf >= 1) lf = 0.75f;
    st->load_factor = lf; st->shard_hi = st->array_size;
}

int dbg_get_stats(dbg_ctx *c, dbg_stats *st)
{
    if (c->full) return DBG_ERR_TABLE_FULL;
    fill(c, st);
    return DBG_OK;
}

int dbg_finalize(dbg_ctx *c, dbg_stats *st)
{
    if (c->full) return DBG_ERR_TABLE_FULL;
    if (!c->finished) { orc_finish(c->set); c->finished = true; }
    fill(c, st);
    // the side node's link words
    const uint64_t n = orc_count(c->set);
    std::vector<uint64_t> slot(n), k(n); std::vector<uint32_t> l(n), r(n);
    orc_dump(c->set, slot.data(), k.data(), nullptr, l.data(), r.data());
    for (uint64_t i = 0; i < n; i++) if (k[i] == 0) { st->polyA_l = l[i]; st->polyA_r = r[i]; }
    return DBG_OK;
}

int dbg_export_kmerset(dbg_ctx *c, void *array, uint8_t *nul_flag)
{
    const uint64_t P = orc_size(c->set);
    memcpy(array, orc_array(c->set), P * 16);
    memcpy(nul_flag, orc_nul_flag(c->set), P / 8 + 1);
    return DBG_OK;
}

int dbg_dump_shard(dbg_ctx *c, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *l_link, uint32_t *r_link, uint64_t *first_ordinal,
                   uint64_t *n)
{
    const uint64_t total = orc_count(c->set), m = total - (c->finished ? 1 : 0), cap = *n;
    *n = m;
    if (!kmers_lo && !kmers_hi && !l_link && !r_link && !first_ordinal) return DBG_OK;
    if (cap < m) return DBG_ERR_BUFFER;
    std::vector<uint64_t> slot(total), k(total); std::vector<uint32_t> l(total), r(total);
    orc_dump(c->set, slot.data(), k.data(), nullptr, l.data(), r.data());
    uint64_t w = 0;
    for (uint64_t i = total; i-- > 0;) {            // reverse slot order: the real dump is unordered, do not rely on order
        if (k[i] == 0) continue;
        if (kmers_lo) kmers_lo[w] = k[i];
        if (kmers_hi) kmers_hi[w] = 0;
        if (l_link) l_link[w] = l[i];
        if (r_link) r_link[w] = r[i];
        if (first_ordinal) first_ordinal[w] = c->first.at(k[i]);
        w++;
    }
    return DBG_OK;
}

}   // extern "C"
