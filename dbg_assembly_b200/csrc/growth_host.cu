// growth_host.cu -- HOST-side replay of the reference's table growth ("emulate enlarge on host").
//
// The GPU build sizes its table once from -i.  The reference instead checks `count > max` after every non-final
// block of -b reads and, while -e allows, doubles the table with an in-place rehash (DBGgraph.cpp:329-351,
// kmerSet.cpp:132-189).  Node CONTENTS do not depend on that, but the slot layout does, and the slot layout drives
// the order and strand of everything the host traversal prints.  This file replays that history on
// the host from what the GPU already has -- every node with the ordinal (read index, offset) of its first
// occurrence -- and writes the table the reference would have ended with:
//
//   1. keys sorted by first occurrence are fed, block of -b reads by block, into a slot-index array with the
//      reference's linear probing;
//   2. at every block boundary where the reference would grow, the in-place rehash is replayed exactly: old slots in
//      ascending order, a node landing on a slot that still holds an un-moved node takes the slot and the evicted
//      node is inserted next (the swap chain of kmerSet.cpp:163-183);
//   3. the k-mer-0 node goes in last (DBGgraph.cpp:418).
//
// It is sequential like the code it mirrors (a few hundred ns per node: hashing and cache misses, again per doubling)
// and only needed when a run outgrows -i; the front end
// calls it instead of dbg_export_kmerset in that case.  If -e is exhausted the reference stops reading the current
// file ("Memory reach the maximum allowed"): contents then differ from a full build, which a replay cannot undo --
// the plan reports `truncated` and no layout is produced.
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

#include "dbg_core.cuh"
#include "../../include/dbg_b200.h"

using namespace dbg;

namespace {

struct Replay {
    bool wide;
    const uint64_t *klo, *khi;
    uint64_t size = 0;
    std::vector<int64_t> slot;     // node index per slot, -1 = empty

    uint64_t home(int64_t i) const
    {
        const uint64_t h = wide ? hash_code_wide(klo[i], khi ? khi[i] : 0) : hash_code(klo[i]);
        return h % size;
    }
    void insert(int64_t i)
    {
        uint64_t hc = home(i);
        while (slot[hc] != -1) hc = (hc + 1 == size) ? 0 : hc + 1;      // DBGgraph.cpp:201-204
        slot[hc] = i;
    }
    // enlarge_kmerset_parallel, kmerSet.cpp:132-189
    void enlarge(uint64_t new_size)
    {
        const uint64_t old_size = size;
        std::vector<uint8_t> unmoved(old_size);                         // filled at the start and not yet re-inserted
        for (uint64_t i = 0; i < old_size; i++) unmoved[i] = slot[i] != -1;
        std::vector<uint8_t> filled(new_size, 0);                       // the new nul_flag: re-inserted nodes only
        slot.resize(new_size, -1);
        size = new_size;
        for (uint64_t i = 0; i < old_size; i++) {
            if (!unmoved[i]) continue;
            int64_t t = slot[i];
            slot[i] = -1; unmoved[i] = 0;
            for (;;) {
                uint64_t hc = home(t);
                while (filled[hc]) hc = (hc + 1) % size;
                filled[hc] = 1;
                if (hc < old_size && unmoved[hc]) { std::swap(t, slot[hc]); unmoved[hc] = 0; }   // evicted node goes next
                else { slot[hc] = t; break; }
            }
        }
    }
};

}   // namespace

extern "C" int dbg_replay_growth(const dbg_growth_params *g, const uint64_t *reads_per_file, uint32_t n_files,
                                 const uint64_t *kmer_lo, const uint64_t *kmer_hi, const uint32_t *l_link, const uint32_t *r_link,
                                 const uint64_t *first_ordinal, uint64_t n_nodes, uint32_t polyA_l, uint32_t polyA_r,
                                 dbg_growth_result *res, void *array, uint8_t *nul_flag)
{
    if (!g || !res || (n_files && !reads_per_file) || (n_nodes && (!kmer_lo || !first_ordinal))) return DBG_ERR_INVALID;
    if (g->wide && n_nodes && !kmer_hi) return DBG_ERR_INVALID;
    if (array && n_nodes && (!l_link || !r_link)) return DBG_ERR_INVALID;
    memset(res, 0, sizeof(*res));
    // init_kmerset_parallel, kmerSet.cpp:98-115
    uint64_t size = g->init_slots < 3 ? 3 : dbg_find_next_prime(g->init_slots);
    float lf = g->load_factor;
    if (lf <= 0) lf = 0.25f; else if (lf >= 1) lf = 0.75f;
    uint64_t max = (uint64_t)(size * lf);
    const uint64_t B = g->buffer_reads ? g->buffer_reads : 10000;

    // the blocks of -b reads, file by file, in reading order: block q ends before global read index block_end[q];
    // a grow check follows every FULL block (parse_one_reads_file, DBGgraph.cpp:217-359: the final, short -- possibly
    // empty -- block of a file has none)
    std::vector<uint64_t> file_base(n_files + 1, 0), file_block0(n_files + 1, 0);
    for (uint32_t f = 0; f < n_files; f++) {
        file_base[f + 1] = file_base[f] + reads_per_file[f];
        file_block0[f + 1] = file_block0[f] + reads_per_file[f] / B + 1;            // full blocks + the final short one
    }
    const uint64_t n_blocks = file_block0[n_files];
    // new nodes per block (no sort needed for the plan)
    std::vector<uint64_t> new_in_block(n_blocks + 1, 0);
    for (uint64_t i = 0; i < n_nodes; i++) {
        const uint64_t read = first_ordinal[i] >> 16;
        if (read >= file_base[n_files]) return DBG_ERR_INVALID;                     // ordinal beyond the declared reads
        const uint32_t f = (uint32_t)(std::upper_bound(file_base.begin(), file_base.end(), read) - file_base.begin() - 1);
        const uint64_t q = file_block0[f] + (read - file_base[f]) / B;
        new_in_block[q]++;
    }

    const bool layout = array != nullptr;
    Replay rp;
    rp.wide = g->wide != 0; rp.klo = kmer_lo; rp.khi = kmer_hi; rp.size = size;
    // layout: nodes in first-occurrence order (sort (ordinal, index) pairs: sequential compares, no indirection)
    std::vector<int64_t> order;
    if (layout) {
        rp.slot.assign(size, -1);
        std::vector<std::pair<uint64_t, int64_t> > keyed(n_nodes);
        for (uint64_t i = 0; i < n_nodes; i++) keyed[i] = std::make_pair(first_ordinal[i], (int64_t)i);
        std::sort(keyed.begin(), keyed.end());
        order.resize(n_nodes);
        for (uint64_t i = 0; i < n_nodes; i++) {
            if (i && keyed[i].first == keyed[i - 1].first) return DBG_ERR_INVALID;  // ordinals must be unique
            order[i] = keyed[i].second;
        }
    }

    uint64_t pos = 0, count = 0, doublings = 0;
    for (uint32_t f = 0; f < n_files && !res->truncated; f++) {
        const uint64_t nf = reads_per_file[f];
        const uint64_t full_blocks = nf / B;
        for (uint64_t qb = 0; qb <= full_blocks; qb++) {
            const uint64_t q = file_block0[f] + qb;
            // a block that fills the table makes the reference probe forever (no free slot): report it instead
            if (count + new_in_block[q] >= size && new_in_block[q]) return DBG_ERR_TABLE_FULL;
            if (layout)
                for (uint64_t e = pos + new_in_block[q]; pos < e; pos++) rp.insert(order[pos]);
            count += new_in_block[q];
            if (qb == full_blocks) break;                                 // final block of the file: no grow check (:329-331)
            if (count > max) {                                            // :337-351
                if (doublings < g->max_double_times) {
                    uint64_t new_size = size;
                    do { new_size = dbg_find_next_prime(new_size * 2); } while (new_size * lf < count + 1);   // kmerSet.cpp:137-139
                    if (layout) rp.enlarge(new_size);
                    size = new_size;
                    max = (uint64_t)(size * lf);
                    doublings++;
                } else {
                    // the reference ignores the rest of this file: a node first seen there must not exist
                    // (and occurrences there must not have been counted): any ignored read invalidates a full build
                    const uint64_t next = (qb + 1) * B;
                    if (next < nf) { res->truncated = 1; res->truncated_file = f; res->truncated_first_read = file_base[f] + next; }
                    break;
                }
            }
        }
    }
    res->final_size = size; res->final_max = max; res->doublings = doublings; res->count = count + 1;
    if (res->truncated) return DBG_OK;
    if (!layout) return DBG_OK;

    // the k-mer-0 node, last and always (add_node_to_kmerset, kmerSet.cpp:253-273 / DBGgraph.cpp:418)
    uint64_t hz = (rp.wide ? hash_code_wide(0, 0) : hash_code(0)) % size;
    while (rp.slot[hz] != -1) hz = (hz + 1 == size) ? 0 : hz + 1;
    const int64_t POLYA = (int64_t)n_nodes;
    rp.slot[hz] = POLYA;

    const size_t nb = rp.wide ? sizeof(dbg_node32) : sizeof(dbg_node16);
    memset(array, 0, size * nb);
    if (nul_flag) memset(nul_flag, 0, size / 8 + 1);
    for (uint64_t s = 0; s < size; s++) {
        const int64_t i = rp.slot[s];
        if (i < 0) continue;
        const uint64_t lo = i == POLYA ? 0 : kmer_lo[i], hi = (i == POLYA || !kmer_hi) ? 0 : kmer_hi[i];
        const uint32_t l = i == POLYA ? polyA_l : l_link[i], r = i == POLYA ? polyA_r : r_link[i];
        if (rp.wide) { dbg_node32 *n = static_cast<dbg_node32 *>(array) + s; n->kmer_lo = lo; n->kmer_hi = hi; n->l_link = l; n->r_link = r; n->pad = 0; }
        else { dbg_node16 *n = static_cast<dbg_node16 *>(array) + s; n->kmer = lo; n->l_link = l; n->r_link = r; }
        if (nul_flag) nul_flag[s >> 3] |= (uint8_t)(0x80u >> (s & 7));                          // MSB first, kmerSet.h:144-155
    }
    return DBG_OK;
}
