// checkpoint_host.cu -- on-disk checkpoint of the finished graph: the KmerSet that
// build_debruijn_graph hands to the traversal (kmerSet.h:88-99), as a compact file of its filled slots in slot order
// {slot, kmer[, kmer_hi], l_link, r_link}.  The reference has no restart point between the build and the traversal
// (main.cpp:204-207 keeps everything in memory); with this a front end can re-run the traversal with other cut-offs
// (-D -T -I -P ...) without touching the reads again.  Host code only: plain file I/O around the table image.
//
//   file = header (64 B) | count records (24 B, or 32 B with 128-bit keys) | u64 FNV-1a of the records
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/dbg_b200.h"

namespace {

const uint64_t CKPT_MAGIC = 0x4332303042474244ULL;     // "DBGB200C"

struct Rec24 { uint64_t slot, kmer; uint32_t l, r; };
struct Rec32 { uint64_t slot, kmer, kmer_hi; uint32_t l, r; };

uint64_t fnv(uint64_t h, const void *p, size_t n)
{
    const unsigned char *q = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; i++) { h ^= q[i]; h *= 1099511628211ull; }
    return h;
}

bool filled(const uint8_t *nul, uint64_t s) { return (nul[s >> 3] >> (7 - (s & 7))) & 1; }    // kmerSet.h:144-147

}   // namespace

extern "C" int dbg_checkpoint_write(const char *path, const dbg_checkpoint_header *hdr, const void *array, const uint8_t *nul_flag)
{
    if (!path || !hdr || !array || !nul_flag) return DBG_ERR_INVALID;
    FILE *fp = fopen(path, "wb");
    if (!fp) return DBG_ERR_INVALID;
    dbg_checkpoint_header h = *hdr;
    h.magic = CKPT_MAGIC; h.version = 1;
    const bool wide = h.wide != 0;
    const size_t rb = wide ? sizeof(Rec32) : sizeof(Rec24);
    // filled slots, gathered by a few threads over contiguous slot ranges, written in slot order
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    if (h.size < (1u << 20)) nt = 1;
    std::vector<std::vector<unsigned char> > part(nt);
    auto work = [&](unsigned t) {
        const uint64_t s0 = h.size * t / nt, s1 = h.size * (t + 1) / nt;
        std::vector<unsigned char> &out = part[t];
        for (uint64_t s = s0; s < s1; s++) {
            if (!filled(nul_flag, s)) continue;
            const size_t at = out.size();
            out.resize(at + rb);
            if (wide) { const dbg_node32 &nd = static_cast<const dbg_node32 *>(array)[s]; Rec32 r = {s, nd.kmer_lo, nd.kmer_hi, nd.l_link, nd.r_link}; memcpy(&out[at], &r, rb); }
            else { const dbg_node16 &nd = static_cast<const dbg_node16 *>(array)[s]; Rec24 r = {s, nd.kmer, nd.l_link, nd.r_link}; memcpy(&out[at], &r, rb); }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    uint64_t n = 0;
    for (auto &p : part) n += p.size() / rb;
    h.records = n;
    bool ok = fwrite(&h, sizeof(h), 1, fp) == 1;
    uint64_t sum = 1469598103934665603ull;
    for (auto &p : part) {
        if (p.empty()) continue;
        sum = fnv(sum, p.data(), p.size());
        ok = ok && fwrite(p.data(), 1, p.size(), fp) == p.size();
    }
    ok = ok && fwrite(&sum, sizeof(sum), 1, fp) == 1;
    ok = (fclose(fp) == 0) && ok;
    return ok ? DBG_OK : DBG_ERR_BUFFER;
}

extern "C" int dbg_checkpoint_read_header(const char *path, dbg_checkpoint_header *hdr)
{
    if (!path || !hdr) return DBG_ERR_INVALID;
    FILE *fp = fopen(path, "rb");
    if (!fp) return DBG_ERR_INVALID;
    const bool ok = fread(hdr, sizeof(*hdr), 1, fp) == 1 && hdr->magic == CKPT_MAGIC && hdr->version == 1;
    fclose(fp);
    return ok ? DBG_OK : DBG_ERR_INVALID;
}

// array[size] (16-B or 32-B nodes) and nul_flag[size/8+1] are caller-allocated from the header's size; both are zeroed
// here, then the records are put back into their slots
extern "C" int dbg_checkpoint_read(const char *path, void *array, uint8_t *nul_flag)
{
    if (!path || !array || !nul_flag) return DBG_ERR_INVALID;
    dbg_checkpoint_header h;
    int rc = dbg_checkpoint_read_header(path, &h);
    if (rc) return rc;
    FILE *fp = fopen(path, "rb");
    if (!fp) return DBG_ERR_INVALID;
    fseek(fp, (long)sizeof(h), SEEK_SET);
    const bool wide = h.wide != 0;
    const size_t rb = wide ? sizeof(Rec32) : sizeof(Rec24), nb = wide ? sizeof(dbg_node32) : sizeof(dbg_node16);
    memset(array, 0, h.size * nb);
    memset(nul_flag, 0, h.size / 8 + 1);
    std::vector<unsigned char> buf((size_t)(1u << 16) * rb);
    uint64_t left = h.records, sum = 1469598103934665603ull, prev = 0;
    bool ok = true, first = true;
    while (left && ok) {
        const size_t m = left < (1u << 16) ? (size_t)left : (size_t)(1u << 16);
        ok = fread(buf.data(), rb, m, fp) == m;
        if (!ok) break;
        sum = fnv(sum, buf.data(), m * rb);
        for (size_t i = 0; i < m && ok; i++) {
            uint64_t s;
            if (wide) { Rec32 r; memcpy(&r, &buf[i * rb], rb); s = r.slot; if (s < h.size) { dbg_node32 &nd = static_cast<dbg_node32 *>(array)[s]; nd.kmer_lo = r.kmer; nd.kmer_hi = r.kmer_hi; nd.l_link = r.l; nd.r_link = r.r; nd.pad = 0; } }
            else { Rec24 r; memcpy(&r, &buf[i * rb], rb); s = r.slot; if (s < h.size) { dbg_node16 &nd = static_cast<dbg_node16 *>(array)[s]; nd.kmer = r.kmer; nd.l_link = r.l; nd.r_link = r.r; } }
            ok = s < h.size && (first || s > prev);              // slot order, no duplicates
            if (ok) nul_flag[s >> 3] |= (uint8_t)(0x80u >> (s & 7));
            prev = s; first = false;
        }
        left -= m;
    }
    uint64_t want = 0;
    ok = ok && fread(&want, sizeof(want), 1, fp) == 1 && want == sum;
    fclose(fp);
    return ok ? DBG_OK : DBG_ERR_INVALID;
}
