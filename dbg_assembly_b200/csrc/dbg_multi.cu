// dbg_multi.cu -- ONE host process driving several GPUs: the multi-GPU build behind the same five calls a front end
// makes for one GPU (create / submit_reads / finalize / export_kmerset / destroy), so that the reference's
// single-process front end (debruijn_contig: main.cpp:204-207 hands `kset` from the build to the traversal) can use a
// whole B200 box.  It is built from the per-GPU pieces of the C ABI (include/dbg_b200.h):
//
//   reads of a block are dealt to the GPUs in contiguous parts (H2D per GPU, all PCIe links in parallel);
//   every GPU extracts its part once and stores the tuples straight into fixed regions of the owners' receive buffers
//   over NVLink peer mappings (dbg_exchange_scatter_opt_device; owner = slot range, the reference's `kmer % threadNum`
//   split lifted to GPUs, DBGgraph.cpp:148); a region that would overflow doubles the regions and redoes the round
//   (nothing was inserted yet); every owner partitions + inserts what it received (dbg_insert_tuple_regions_device);
//   finalize: k-mer-0 counters summed, boundary clusters handed around the ring (dbg_shard_tail_export/_import), every
//   GPU lays out its slice of the reference's table; export: the slices are copied into the caller's ONE table image
//   in parallel, the shared nul_flag bytes are fixed and the k-mer-0 node goes in last (DBGgraph.cpp:418).
//   If a boundary cluster does not fit the hand-off (tables of a few thousand slots, pathologically dense ones) the
//   table is rebuilt on the host from the shard dumps by replaying the keys in first-occurrence order (same layout).
//
// One worker thread per GPU runs the per-GPU calls of a phase; phases are separated by joins.
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dbg_b200.h"

namespace {

thread_local char g_mg_err[512] = "";

int mg_fail(int code, const char *what)
{
    // a worker thread may already have left its (thread-local) message here: keep it, add the call's name
    char prev[sizeof(g_mg_err)];
    snprintf(prev, sizeof(prev), "%s", g_mg_err[0] && strncmp(g_mg_err, "GPU ", 4) == 0 ? g_mg_err : dbg_last_error());
    snprintf(g_mg_err, sizeof(g_mg_err), "%.60s: %.440s", what, prev);
    return code;
}

}   // namespace

struct dbg_mg {
    int n = 0;
    std::vector<int> dev;
    std::vector<dbg_ctx *> ctx;
    dbg_params prm;
    int tuple_bytes = 16;
    // per GPU: staging of its part of a block, receive buffer (n regions of cap_pair tuples), pointer table, fill counters
    std::vector<char *> d_bases;
    std::vector<uint64_t *> d_offs;
    uint64_t cap_bases = 0, cap_reads = 0;
    std::vector<void *> recv;
    std::vector<void **> d_ptrs;
    std::vector<uint32_t *> d_fill;
    uint64_t cap_pair = 0;
    uint64_t next_read = 0;
    bool finalized = false, fallback = false;
    std::vector<std::vector<uint8_t> > blobs;
    dbg_stats st;
    uint64_t rounds = 0, regrows = 0;
};

static void run_per_gpu(dbg_mg *m, const std::function<int(int)> &f, std::vector<int> &rc)
{
    rc.assign(m->n, 0);
    std::vector<std::string> msg(m->n);                      // dbg_last_error() is thread-local: fetch it in the worker
    std::vector<std::thread> th;
    for (int r = 1; r < m->n; r++) th.emplace_back([&, r]() { cudaSetDevice(m->dev[r]); rc[r] = f(r); if (rc[r]) msg[r] = dbg_last_error(); });
    cudaSetDevice(m->dev[0]);
    rc[0] = f(0);
    for (auto &t : th) t.join();
    for (int r = 1; r < m->n; r++)
        if (rc[r] && rc[0] == DBG_OK) { snprintf(g_mg_err, sizeof(g_mg_err), "GPU %d: %s", m->dev[r], msg[r].c_str()); break; }
}

static int first_error(const std::vector<int> &rc)
{
    for (int v : rc) if (v != DBG_OK && v != DBG_ERR_STATE) return v;
    for (int v : rc) if (v != DBG_OK) return v;
    return DBG_OK;
}

extern "C" const char *dbg_mg_last_error(void) { return g_mg_err; }

extern "C" void dbg_mg_destroy(dbg_mg *m)
{
    if (!m) return;
    for (int r = 0; r < m->n; r++) {
        cudaSetDevice(m->dev[r]);
        cudaDeviceSynchronize();
        if (r < (int)m->d_bases.size()) cudaFree(m->d_bases[r]);
        if (r < (int)m->d_offs.size()) cudaFree(m->d_offs[r]);
        if (r < (int)m->recv.size()) cudaFree(m->recv[r]);
        if (r < (int)m->d_ptrs.size()) cudaFree(m->d_ptrs[r]);
        if (r < (int)m->d_fill.size()) cudaFree(m->d_fill[r]);
        if (r < (int)m->ctx.size() && m->ctx[r]) dbg_destroy(m->ctx[r]);
    }
    delete m;
}

static int alloc_recv(dbg_mg *m, uint64_t cap_pair)
{
    for (int r = 0; r < m->n; r++) {
        cudaSetDevice(m->dev[r]);
        cudaDeviceSynchronize();
        if (m->recv[r]) { cudaFree(m->recv[r]); m->recv[r] = nullptr; }
        if (cudaMalloc(&m->recv[r], (size_t)m->n * cap_pair * m->tuple_bytes + 256) != cudaSuccess) { cudaGetLastError(); return DBG_ERR_NOMEM; }
    }
    for (int r = 0; r < m->n; r++) {
        cudaSetDevice(m->dev[r]);
        if (cudaMemcpy(m->d_ptrs[r], m->recv.data(), m->n * sizeof(void *), cudaMemcpyHostToDevice) != cudaSuccess) return DBG_ERR_CUDA;
    }
    m->cap_pair = cap_pair;
    return DBG_OK;
}

extern "C" int dbg_mg_create(dbg_mg **out, const dbg_params *p, int32_t n_gpus, const int32_t *devices)
{
    if (!out || !p || n_gpus < 1 || n_gpus > 64) { snprintf(g_mg_err, sizeof(g_mg_err), "dbg_mg_create: bad argument"); return DBG_ERR_INVALID; }
    *out = nullptr;
    const int ndev = dbg_device_count();
    if (ndev == 0) { snprintf(g_mg_err, sizeof(g_mg_err), "no CUDA device visible: libdbgb200 has no CPU fallback"); return DBG_ERR_CUDA; }
    dbg_mg *m = new dbg_mg();
    m->n = n_gpus;
    m->prm = *p;
    for (int r = 0; r < n_gpus; r++) {
        const int d = devices ? devices[r] : r;
        if (d < 0 || d >= ndev) { delete m; snprintf(g_mg_err, sizeof(g_mg_err), "device %d of %d", d, ndev); return DBG_ERR_INVALID; }
        m->dev.push_back(d);
    }
    *out = m;      // from here on the caller can dbg_mg_destroy() after a failure
    // peer access between every pair of distinct devices (stores into the owners' receive buffers go over NVLink)
    for (int i = 0; i < n_gpus; i++)
        for (int j = 0; j < n_gpus; j++) {
            if (m->dev[i] == m->dev[j]) continue;
            cudaSetDevice(m->dev[i]);
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]);
            if (!can) { snprintf(g_mg_err, sizeof(g_mg_err), "device %d cannot access device %d (no peer access)", m->dev[i], m->dev[j]); return DBG_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { snprintf(g_mg_err, sizeof(g_mg_err), "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return DBG_ERR_CUDA; }
            cudaGetLastError();
        }
    m->ctx.assign(n_gpus, nullptr);
    m->d_bases.assign(n_gpus, nullptr); m->d_offs.assign(n_gpus, nullptr); m->recv.assign(n_gpus, nullptr);
    m->d_ptrs.assign(n_gpus, nullptr); m->d_fill.assign(n_gpus, nullptr);
    for (int r = 0; r < n_gpus; r++) {
        dbg_params q = *p;
        q.device = m->dev[r]; q.shard_rank = r; q.shard_count = n_gpus;
        int rc = dbg_create(&m->ctx[r], &q);
        if (rc) return mg_fail(rc, "dbg_create");
        cudaSetDevice(m->dev[r]);
        if (cudaMalloc(&m->d_ptrs[r], n_gpus * sizeof(void *)) != cudaSuccess || cudaMalloc(&m->d_fill[r], (n_gpus + 1) * sizeof(uint32_t)) != cudaSuccess) {
            snprintf(g_mg_err, sizeof(g_mg_err), "out of device memory");
            return DBG_ERR_NOMEM;
        }
    }
    m->tuple_bytes = dbg_tuple_bytes(m->ctx[0]);
    // one round = up to 64 Mi bases per GPU
    m->cap_bases = 64ull << 20;
    m->cap_reads = 4ull << 20;
    if (const char *e = getenv("DBG_B200_MG_ROUND_BASES")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 1024) m->cap_bases = v; }
    for (int r = 0; r < n_gpus; r++) {
        cudaSetDevice(m->dev[r]);
        if (cudaMalloc(&m->d_bases[r], m->cap_bases + 64) != cudaSuccess || cudaMalloc(&m->d_offs[r], (m->cap_reads + 2) * sizeof(uint64_t)) != cudaSuccess) {
            snprintf(g_mg_err, sizeof(g_mg_err), "out of device memory");
            return DBG_ERR_NOMEM;
        }
    }
    uint64_t cap_pair = (uint64_t)((double)m->cap_bases / n_gpus * 1.25) + 4096;
    if (const char *e = getenv("DBG_B200_MG_CAP_PAIR")) { uint64_t v = strtoull(e, nullptr, 10); if (v >= 64) cap_pair = v; }     // tests: provoke the regrow
    int rc = alloc_recv(m, cap_pair);
    if (rc) { snprintf(g_mg_err, sizeof(g_mg_err), "receive buffers: out of device memory"); return rc; }
    return DBG_OK;
}

// one exchange round over reads [r0, r1) of the caller's block
static int mg_round(dbg_mg *m, const char *bases, const uint64_t *offs, uint64_t r0, uint64_t r1)
{
    const int n = m->n;
    // contiguous parts with about the same number of bases
    std::vector<uint64_t> cut(n + 1, r1);
    cut[0] = r0;
    const uint64_t total = offs[r1] - offs[r0];
    for (int q = 1; q < n; q++) {
        const uint64_t want = offs[r0] + total * q / n;
        uint64_t lo = cut[q - 1], hi = r1;
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (offs[mid] < want) lo = mid + 1; else hi = mid; }
        cut[q] = lo;
    }
    std::vector<int> rc;
    std::vector<std::vector<uint32_t> > fill(n, std::vector<uint32_t>(n + 1, 0));
    // H2D of the parts (all links in parallel), then the fused extract + scatter into the owners' regions
    for (;;) {
        run_per_gpu(m, [&](int r) -> int {
            const uint64_t a = cut[r], b = cut[r + 1];
            if (a == b) return DBG_OK;
            if (b - a > m->cap_reads) return DBG_ERR_BUFFER;
            const uint64_t fb = offs[a], nb = offs[b] - offs[a], shift = fb & 15;
            if (nb && cudaMemcpy(m->d_bases[r] + shift, bases + fb, nb, cudaMemcpyHostToDevice) != cudaSuccess) return DBG_ERR_CUDA;
            if (cudaMemcpy(m->d_offs[r], offs + a, (b - a + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice) != cudaSuccess) return DBG_ERR_CUDA;
            if (cudaDeviceSynchronize() != cudaSuccess) return DBG_ERR_CUDA;      // pageable sources: the DMA has landed before the kernels start
            // d_base such that d_base[off] is base `off` of the block's offset space
            const char *d_base = m->d_bases[r] + shift - fb;
            int e = dbg_exchange_scatter_opt_device(m->ctx[r], d_base, m->d_offs[r], b - a, fb, nb, m->next_read + (a - r0), n, m->d_ptrs[r],
                                                    (uint64_t)r * m->cap_pair, (uint32_t)m->cap_pair, m->d_fill[r], nullptr);
            if (e) return e;
            if (cudaDeviceSynchronize() != cudaSuccess) return DBG_ERR_CUDA;
            if (cudaMemcpy(fill[r].data(), m->d_fill[r], (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return DBG_ERR_CUDA;
            return DBG_OK;
        }, rc);
        int e = first_error(rc);
        if (e) return mg_fail(e, "dbg_exchange_scatter_opt_device");
        bool overflow = false;
        for (int r = 0; r < n; r++) overflow |= fill[r][n] != 0;
        if (!overflow) break;
        // a region was too small (skewed input): nothing was inserted; take back the side counters, double the regions, redo
        run_per_gpu(m, [&](int r) -> int {
            if (cut[r] == cut[r + 1]) return DBG_OK;
            int e2 = dbg_exchange_scatter_undo(m->ctx[r], nullptr);
            if (e2) return e2;
            return cudaDeviceSynchronize() == cudaSuccess ? DBG_OK : DBG_ERR_CUDA;
        }, rc);
        if ((e = first_error(rc))) return mg_fail(e, "dbg_exchange_scatter_undo");
        if (m->cap_pair > (1ull << 31)) { snprintf(g_mg_err, sizeof(g_mg_err), "exchange regions cannot grow further"); return DBG_ERR_NOMEM; }
        if ((e = alloc_recv(m, m->cap_pair * 2))) { snprintf(g_mg_err, sizeof(g_mg_err), "receive buffers: out of device memory"); return e; }
        m->regrows++;
    }
    // owners: partition by table slice + bucketed insert of what they received
    run_per_gpu(m, [&](int q) -> int {
        std::vector<uint64_t> counts(n);
        uint64_t tot = 0;
        for (int r = 0; r < n; r++) { counts[r] = fill[r][q]; tot += counts[r]; }
        if (tot == 0) return DBG_OK;
        return dbg_insert_tuple_regions_device(m->ctx[q], m->recv[q], (uint32_t)n, m->cap_pair, counts.data(), nullptr);
    }, rc);
    int e = first_error(rc);
    if (e) return mg_fail(e, "dbg_insert_tuple_regions_device");
    m->next_read += r1 - r0;
    m->rounds++;
    return DBG_OK;
}

extern "C" int dbg_mg_submit_reads(dbg_mg *m, const char *bases, const uint64_t *offs, uint64_t n_reads)
{
    if (!m || (!bases && n_reads) || (!offs && n_reads)) { snprintf(g_mg_err, sizeof(g_mg_err), "dbg_mg_submit_reads: NULL argument"); return DBG_ERR_INVALID; }
    if (m->finalized) { snprintf(g_mg_err, sizeof(g_mg_err), "submit after finalize"); return DBG_ERR_STATE; }
    uint64_t r0 = 0;
    while (r0 < n_reads) {
        // a round: as many reads as fit n x (cap_bases, cap_reads); a single read longer than cap_bases is cut to -r by the
        // per-GPU build anyway, but must still fit the staging: refuse instead of truncating silently
        const uint64_t max_bases = m->cap_bases * m->n / 2, max_reads = m->cap_reads;
        uint64_t lo = r0 + 1, hi = r0 + max_reads < n_reads ? r0 + max_reads : n_reads;
        while (lo < hi) { uint64_t mid = (lo + hi + 1) / 2; if (offs[mid] - offs[r0] <= max_bases) lo = mid; else hi = mid - 1; }
        const uint64_t r1 = lo;
        if (offs[r1] - offs[r0] > m->cap_bases) {
            // parts are cut by bases, so a part can exceed cap_bases only through one huge read
            uint64_t longest = 0;
            for (uint64_t i = r0; i < r1; i++) if (offs[i + 1] - offs[i] > longest) longest = offs[i + 1] - offs[i];
            if (longest > m->cap_bases / 2) { snprintf(g_mg_err, sizeof(g_mg_err), "a read of %llu bases does not fit the per-GPU staging", (unsigned long long)longest); return DBG_ERR_INVALID; }
        }
        int rc = mg_round(m, bases, offs, r0, r1);
        if (rc) return rc;
        r0 = r1;
    }
    return DBG_OK;
}

// decode the overflow (margin) nodes of a tail blob: raw build nodes -> (kmer, links, first ordinal)
static float half_to_float(uint16_t h)
{
    const uint32_t e = (h >> 10) & 31, f = h & 1023;
    if (e == 0) return (float)f / 16777216.0f;      // subnormal: f * 2^-24
    float v = 1.0f + (float)f / 1024.0f;
    int ex = (int)e - 15;
    while (ex > 0) { v *= 2.0f; ex--; }
    while (ex < 0) { v *= 0.5f; ex++; }
    return v;
}

static void blob_margin(const std::vector<uint8_t> &blob, bool wide, std::vector<uint64_t> &klo, std::vector<uint64_t> &khi, std::vector<uint32_t> &l,
                        std::vector<uint32_t> &r, std::vector<uint64_t> &ord)
{
    if (blob.size() < 32) return;
    uint64_t hdr[4];
    memcpy(hdr, blob.data(), 32);
    const uint64_t a = hdr[0], mt = hdr[1], nbb = hdr[2];
    for (uint64_t i = 0; i < mt; i++) {
        const uint8_t *nd = blob.data() + 32 + (a + i) * nbb;
        uint64_t w[8];
        memcpy(w, nd, nbb < 64 ? nbb : 64);
        const uint64_t lo = w[0], hi = wide ? w[1] : 0, nord = wide ? w[2] : w[1];
        const uint16_t *c = reinterpret_cast<const uint16_t *>(nd + (wide ? 32 : 16));
        uint32_t lk[2] = {0, 0};
        for (int s = 0; s < 2; s++)
            for (int b = 0; b < 4; b++) {
                float v = half_to_float(c[4 * s + b]);
                uint32_t q = v > 255.f ? 255u : (uint32_t)v;
                lk[s] |= q << (24 - 8 * b);
            }
        klo.push_back(lo); khi.push_back(hi); l.push_back(lk[0]); r.push_back(lk[1]); ord.push_back(~nord);
    }
}

extern "C" int dbg_mg_finalize(dbg_mg *m, dbg_stats *stats)
{
    if (!m) { snprintf(g_mg_err, sizeof(g_mg_err), "NULL"); return DBG_ERR_INVALID; }
    const int n = m->n;
    std::vector<int> rc;
    if (!m->finalized) {
        // k-mer-0 side counters: sum over the GPUs (DBGgraph.cpp:153-164 accumulates them in one node)
        uint64_t sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < n; r++) {
            uint64_t c[8];
            int e = dbg_get_polyA_counts(m->ctx[r], c);
            if (e) return mg_fail(e, "dbg_get_polyA_counts");
            for (int i = 0; i < 8; i++) sum[i] += c[i];
        }
        for (int r = 0; r < n; r++) { int e = dbg_set_polyA_counts(m->ctx[r], sum); if (e) return mg_fail(e, "dbg_set_polyA_counts"); }
        // boundary clusters around the ring, then every GPU lays out its window
        m->blobs.assign(n, std::vector<uint8_t>());
        m->fallback = false;
        for (int r = 0; r < n && !m->fallback; r++) {
            uint64_t nb = 0;
            int e = dbg_shard_tail_export(m->ctx[r], nullptr, 0, &nb);
            if (e == DBG_ERR_STATE) { m->fallback = true; break; }
            if (e) return mg_fail(e, "dbg_shard_tail_export");
            m->blobs[r].resize(nb);
            e = dbg_shard_tail_export(m->ctx[r], m->blobs[r].data(), nb, &nb);
            if (e == DBG_ERR_STATE) { m->fallback = true; break; }
            if (e) return mg_fail(e, "dbg_shard_tail_export");
        }
        for (int r = 0; r < n && !m->fallback; r++) {
            const std::vector<uint8_t> &b = m->blobs[(r + n - 1) % n];
            int e = dbg_shard_tail_import(m->ctx[r], b.data(), b.size());
            if (e == DBG_ERR_STATE) { m->fallback = true; break; }
            if (e) return mg_fail(e, "dbg_shard_tail_import");
        }
        std::vector<dbg_stats> st(n);
        run_per_gpu(m, [&](int r) -> int { return dbg_finalize(m->ctx[r], &st[r]); }, rc);
        int e = first_error(rc);
        if (e == DBG_ERR_STATE) { m->fallback = true; e = DBG_OK; for (int r = 0; r < n; r++) if (rc[r] == DBG_ERR_STATE) { int e2 = dbg_get_stats(m->ctx[r], &st[r]); if (e2) return mg_fail(e2, "dbg_get_stats"); } }
        if (e) return mg_fail(e, "dbg_finalize");
        dbg_stats g = st[0];
        g.count = 1; g.conflict = 0; g.reads = 0; g.kmers_logged = 0; g.occurrences = 0;
        for (int r = 0; r < n; r++) { g.count += st[r].count; g.conflict += st[r].conflict; g.reads += st[r].reads; g.kmers_logged += st[r].kmers_logged; g.occurrences += st[r].occurrences; }
        g.shard_lo = 0; g.shard_hi = g.array_size;
        // link words of the k-mer-0 node from the summed counters
        uint32_t pl = 0, pr = 0;
        for (int b = 0; b < 4; b++) { pl |= (uint32_t)(sum[b] > 255 ? 255 : sum[b]) << (24 - 8 * b); pr |= (uint32_t)(sum[4 + b] > 255 ? 255 : sum[4 + b]) << (24 - 8 * b); }
        g.polyA_l = pl; g.polyA_r = pr;
        if (g.count > g.array_size) { snprintf(g_mg_err, sizeof(g_mg_err), "%llu nodes do not fit %llu slots", (unsigned long long)g.count, (unsigned long long)g.array_size); return DBG_ERR_TABLE_FULL; }
        m->st = g;
        m->finalized = true;
    }
    if (stats) *stats = m->st;
    return DBG_OK;
}

extern "C" int dbg_mg_get_stats(dbg_mg *m, dbg_stats *stats)
{
    if (!m || !stats) { snprintf(g_mg_err, sizeof(g_mg_err), "NULL"); return DBG_ERR_INVALID; }
    if (m->finalized) { *stats = m->st; return DBG_OK; }
    dbg_stats g;
    memset(&g, 0, sizeof(g));
    for (int r = 0; r < m->n; r++) {
        dbg_stats s;
        int e = dbg_get_stats(m->ctx[r], &s);
        if (e) return mg_fail(e, "dbg_get_stats");
        if (r == 0) g = s;
        else { g.count += s.count; g.conflict += s.conflict; g.reads += s.reads; g.kmers_logged += s.kmers_logged; g.occurrences += s.occurrences; }
    }
    g.shard_lo = 0; g.shard_hi = g.array_size;
    if (g.count + 1 > g.array_size) { snprintf(g_mg_err, sizeof(g_mg_err), "%llu nodes so far do not fit %llu slots", (unsigned long long)g.count, (unsigned long long)g.array_size); return DBG_ERR_TABLE_FULL; }
    *stats = g;
    return DBG_OK;
}

// every node of the build (no k-mer-0 node) with its first-occurrence ordinal, for dbg_replay_growth; *n = capacity in,
// count out; all-NULL outputs = size query
extern "C" int dbg_mg_dump_nodes(dbg_mg *m, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *l_link, uint32_t *r_link, uint64_t *first_ordinal, uint64_t *n)
{
    if (!m || !n) { snprintf(g_mg_err, sizeof(g_mg_err), "NULL"); return DBG_ERR_INVALID; }
    const bool query = !kmers_lo && !kmers_hi && !l_link && !r_link && !first_ordinal;
    std::vector<uint64_t> cnt(m->n, 0);
    uint64_t total = 0;
    for (int r = 0; r < m->n; r++) {
        int e = dbg_dump_shard(m->ctx[r], nullptr, nullptr, nullptr, nullptr, nullptr, &cnt[r]);
        if (e) return mg_fail(e, "dbg_dump_shard");
        total += cnt[r];
    }
    // overflow nodes a rank handed over but its neighbour could not adopt (fallback case) are listed from the blobs
    std::vector<uint64_t> xlo, xhi, xord; std::vector<uint32_t> xl, xr;
    if (m->fallback) for (auto &b : m->blobs) blob_margin(b, m->st.wide != 0, xlo, xhi, xl, xr, xord);
    const uint64_t cap = *n;
    *n = total + xlo.size();
    if (query) return DBG_OK;
    if (cap < total + xlo.size()) { snprintf(g_mg_err, sizeof(g_mg_err), "dump capacity too small"); return DBG_ERR_BUFFER; }
    uint64_t off = 0;
    for (int r = 0; r < m->n; r++) {
        uint64_t c = cnt[r];
        int e = dbg_dump_shard(m->ctx[r], kmers_lo ? kmers_lo + off : nullptr, kmers_hi ? kmers_hi + off : nullptr, l_link ? l_link + off : nullptr,
                               r_link ? r_link + off : nullptr, first_ordinal ? first_ordinal + off : nullptr, &c);
        if (e) return mg_fail(e, "dbg_dump_shard");
        off += c;
    }
    // append the blob nodes that are not already present (a partly adopted hand-off lists them twice)
    uint64_t kept = off;
    for (size_t i = 0; i < xlo.size(); i++) {
        bool dup = false;
        if (kmers_lo) for (uint64_t j = 0; j < off && !dup; j++) dup = kmers_lo[j] == xlo[i] && (!kmers_hi || kmers_hi[j] == xhi[i]);
        if (dup) continue;
        if (kmers_lo) kmers_lo[kept] = xlo[i];
        if (kmers_hi) kmers_hi[kept] = xhi[i];
        if (l_link) l_link[kept] = xl[i];
        if (r_link) r_link[kept] = xr[i];
        if (first_ordinal) first_ordinal[kept] = xord[i];
        kept++;
    }
    *n = kept;
    return DBG_OK;
}

extern "C" int dbg_mg_export_kmerset(dbg_mg *m, void *array, uint8_t *nul_flag)
{
    if (!m || !array || !nul_flag) { snprintf(g_mg_err, sizeof(g_mg_err), "NULL"); return DBG_ERR_INVALID; }
    if (!m->finalized) { snprintf(g_mg_err, sizeof(g_mg_err), "dbg_mg_export_kmerset needs dbg_mg_finalize"); return DBG_ERR_STATE; }
    const uint64_t P = m->st.array_size;
    const int wide = m->st.wide;
    if (!m->fallback) {
        std::vector<int> rc;
        std::vector<uint64_t> edges(4 * m->n, UINT64_MAX);
        run_per_gpu(m, [&](int r) -> int { return dbg_export_shard_slice(m->ctx[r], array, nul_flag, edges.data() + 4 * r); }, rc);
        int e = first_error(rc);
        if (e) return mg_fail(e, "dbg_export_shard_slice");
        dbg_host_fix_nul_bytes(array, nul_flag, P, wide, edges.data(), edges.size());
        return dbg_host_polyA_insert(array, nul_flag, P, wide, (uint32_t)m->st.polyA_l, (uint32_t)m->st.polyA_r, nullptr);
    }
    // fallback: replay the keys in first-occurrence order on the host (no growth: one unchecked block)
    uint64_t cap = 0;
    int e = dbg_mg_dump_nodes(m, nullptr, nullptr, nullptr, nullptr, nullptr, &cap);
    if (e) return e;
    std::vector<uint64_t> lo(cap + 1), hi(cap + 1), ord(cap + 1);
    std::vector<uint32_t> l(cap + 1), r(cap + 1);
    uint64_t got = cap;
    e = dbg_mg_dump_nodes(m, lo.data(), hi.data(), l.data(), r.data(), ord.data(), &got);
    if (e) return e;
    uint64_t max_read = 0;
    for (uint64_t i = 0; i < got; i++) if ((ord[i] >> 16) > max_read) max_read = ord[i] >> 16;
    dbg_growth_params gp;
    memset(&gp, 0, sizeof(gp));
    gp.init_slots = m->prm.init_slots; gp.load_factor = m->prm.load_factor; gp.wide = wide; gp.max_double_times = 0; gp.buffer_reads = 1ull << 62;
    const uint64_t rpf[1] = {max_read + 1};
    dbg_growth_result res;
    e = dbg_replay_growth(&gp, rpf, 1, lo.data(), hi.data(), l.data(), r.data(), ord.data(), got, (uint32_t)m->st.polyA_l, (uint32_t)m->st.polyA_r, &res, array, nul_flag);
    if (e) { snprintf(g_mg_err, sizeof(g_mg_err), "dbg_replay_growth (dump merge) failed: %d", e); return e; }
    return DBG_OK;
}

extern "C" int dbg_mg_info(const dbg_mg *m, uint64_t info[4])
{
    if (!m || !info) return DBG_ERR_INVALID;
    info[0] = m->rounds; info[1] = m->regrows; info[2] = m->fallback ? 1 : 0; info[3] = m->cap_pair;
    return DBG_OK;
}
