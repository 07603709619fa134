/* dbg_b200.h -- C ABI of libdbgb200.so: the B200 (sm_100a) implementation of DBG_assembly's
 * De Bruijn graph BUILD hot path.  Plain pointers and sizes only; no C++/torch types cross this line.
 *
 * What it replaces in the reference (paths relative to fanagislab/DBG_assembly, DBG_contig/):
 *   the body of build_debruijn_graph()            DBGgraph.cpp:364-430
 *     - thread_parseBlock   (2-bit encode, rolling fwd/rc k-mer, canonical pick, neighbour bases)
 *                                                  DBGgraph.cpp:38-120, seqKmer.cpp:9-41,89-97
 *     - thread_updatekmers  (hash_code % P, linear probing, CAS claim, 8-bit saturating link lanes,
 *                            poly-A side node)     DBGgraph.cpp:126-213, kmerSet.h:105-116
 *     - init_kmerset_parallel / find_next_prime / add_node_to_kmerset
 *                                                  kmerSet.cpp:72-127,253-273
 *   and, optionally, the first pass of build_contig_sequence(): calculate_kmer_links()
 *                                                  contig.cpp:107-205
 * The reference has no FFI layer; its seam is the global `KmerSet *kset` (DBGgraph.h:31) handed from
 * build_debruijn_graph (main.cpp:204) to build_contig_sequence (main.cpp:207).  A front end keeps its
 * getopt/main, calls dbg_create .. dbg_export_kmerset from its build_debruijn_graph(), and gets the
 * KmerSet image in the reference's own slot layout (same as `debruijn_contig -t 1`), so that
 * contig.cpp runs unchanged on it.  integration/DBGgraph_b200.cpp is that binding; INTEGRATION.md
 * explains it.
 *
 * Conventions: every function returns 0 (DBG_OK) or a negative DBG_ERR_* code; no exceptions, no
 * aborts.  Host buffers are caller-owned.  A context is bound to one CUDA device and must be driven
 * from one host thread at a time (the reference enters its build from main() only, main.cpp:204).
 * There is NO CPU fallback: without a CUDA device dbg_create fails with DBG_ERR_CUDA.
 */
#ifndef DBG_B200_H_
#define DBG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBG_OK                 0
#define DBG_ERR_INVALID       -1   /* bad argument (K out of range, NULL pointer, ...)                 */
#define DBG_ERR_CUDA          -2   /* CUDA runtime error; dbg_last_error() has the text                */
#define DBG_ERR_NOMEM         -3   /* device or pinned-host allocation failed                          */
#define DBG_ERR_TABLE_FULL    -4   /* distinct k-mers exceeded the table (reference would enlarge)     */
#define DBG_ERR_STATE         -5   /* call order violated (e.g. submit after finalize)                 */
#define DBG_ERR_BUFFER        -6   /* caller-provided output buffer too small                          */

/* reference node, kmerSet.h:70-75 (16 B, K <= 31) */
typedef struct { uint64_t kmer; uint32_t l_link; uint32_t r_link; } dbg_node16;
/* 128-bit twin for 31 < K <= 63 (no reference counterpart) */
typedef struct { uint64_t kmer_lo; uint64_t kmer_hi; uint32_t l_link; uint32_t r_link; uint64_t pad; } dbg_node32;

typedef struct {
    int32_t  K;              /* -k  k-mer size; 1..31 -> 64-bit path, 32..63 -> 128-bit path (main.cpp:168)  */
    int32_t  max_read_len;   /* -r  reads are trimmed to this length (DBGgraph.cpp:63); <= 65535            */
    uint64_t init_slots;     /* (uint64)(-i * 1e9): raw slot request; the library applies the reference's
                                find_next_prime (kmerSet.cpp:98-105)                                        */
    float    load_factor;    /* -l  (clamped like kmerSet.cpp:110-111); max = (uint64)(P * load_factor)      */
    int32_t  device;         /* CUDA device ordinal                                                          */
    int32_t  track_order;    /* 1: record first-occurrence ordinals so dbg_export_kmerset reproduces the
                                slot layout of the reference run with -t 1; 0: layout is
                                a valid linear-probing layout but cluster-internal order is arbitrary        */
    int32_t  shard_rank;     /* multi-GPU: this context owns home slots [rank*ceil(P/n), ...) of the P-slot  */
    int32_t  shard_count;    /* table; 0/1 = unsharded                                                       */
    int32_t  force_wide;     /* 1: use the 128-bit path even for K <= 31 (tests pin it against the 64-bit)   */
    int32_t  payload_mode;   /* 0: De Bruijn graph nodes (link lanes); 1: contig seed index (seedidx_* below use it):
                                occurrences carry their strand instead of neighbour bases, k-mer 0 is an ordinary key  */
    int32_t  reserved[4];
} dbg_params;

typedef struct {
    uint64_t array_size;     /* P   (KmerSet.size)                                                           */
    uint64_t max_cutoff;     /* KmerSet.max                                                                  */
    uint64_t count;          /* distinct nodes incl. the k-mer-0 node once finalized (KmerSet.count)         */
    uint64_t conflict;       /* probe steps on the device table (NOT comparable to the reference's counter)  */
    uint64_t reads;          /* Total_reads_num                                                              */
    uint64_t kmers_logged;   /* Kmer_total_num: sum of untrimmed len-K+1 (DBGgraph.cpp:101)                  */
    uint64_t occurrences;    /* (k-mer,left,right) updates actually applied = the benchmark's unit of work   */
    uint64_t polyA_l, polyA_r; /* the side node's link words (DBGgraph.cpp:153-164)                         */
    uint64_t shard_lo, shard_hi; /* home-slot range owned by this context                                    */
    float    load_factor;    /* after clamping                                                               */
    int32_t  wide;           /* 1 if the 128-bit path is in use                                              */
} dbg_stats;

typedef struct dbg_ctx dbg_ctx;

/* ---- scalar helpers kept bit-identical to the reference (also used by front ends for logging) ---- */
uint64_t dbg_find_next_prime(uint64_t n);               /* kmerSet.cpp:72-95 incl. its float-sqrt quirk  */
uint64_t dbg_hash_code(uint64_t kmer);                  /* kmerSet.h:105-116                             */
uint64_t dbg_hash_code_wide(uint64_t lo, uint64_t hi);  /* == dbg_hash_code(lo) when hi == 0             */
const char *dbg_strerror(int code);
const char *dbg_last_error(void);                       /* thread-local text of the last failure         */
int  dbg_device_count(void);                            /* 0 when no CUDA device is visible              */

/* pinned host memory for read buffers (front ends read/decode straight into it) */
int  dbg_host_alloc(void **p, uint64_t bytes);
int  dbg_host_free(void *p);
int  dbg_host_register(void *p, uint64_t bytes);        /* page-lock caller-allocated memory (shared table image) */
int  dbg_host_unregister(void *p);

/* ---- life cycle ------------------------------------------------------------------------------- */
/* init_kmerset_parallel + the globals of build_debruijn_graph (DBGgraph.cpp:371-402) */
int  dbg_create(dbg_ctx **ctx, const dbg_params *params);
void dbg_destroy(dbg_ctx *ctx);

/* One reader block (parse_one_reads_file's hand-off, DBGgraph.cpp:244-321): reads are
 * bases[offs[i] .. offs[i+1]) (ASCII, domain [ACGTNacgtn]), i in [0, n_reads).  Blocks are applied in
 * call order.  Returns after the host->device copy of this block; the kernels run asynchronously and
 * overlap the next block's copy. */
int  dbg_submit_reads(dbg_ctx *ctx, const char *bases, const uint64_t *offs, uint64_t n_reads);

/* Same, inputs already resident on this context's device (offs has n_reads+1 entries, offs[0] may be
 * non-zero; total_bases = offs[n_reads]-offs[0]).  first_read_index = global index of read 0 (orders
 * occurrences across ranks); pass UINT64_MAX to continue this context's own running count.
 * stream: a cudaStream_t, or NULL for the context's stream. */
int  dbg_submit_reads_device(dbg_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                             uint64_t first_base, uint64_t total_bases, uint64_t first_read_index, void *stream);

/* ---- multi-GPU building blocks (owner-computes split of thread_updatekmers, DBGgraph.cpp:148) ---- */
/* 16-B occurrence tuple (K <= 31): kmer, meta = ordinal<<8 | right<<4 | left ; wide tuples add kmer_hi
 * (24 B, padded to 32).  Extract all occurrences of a device-resident block into d_tuples, PACKED by owner
 * shard (owner = (hash % P) / ceil(P/n_parts)): the tuples of owner 0 first, then owner 1, ...;
 * d_counts[n_parts] receives the sizes (offsets = their prefix sums).  `capacity` (in tuples) must be at
 * least total_bases (an occurrence starts at a distinct base), else DBG_ERR_BUFFER. */
int  dbg_extract_tuples_device(dbg_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                               uint64_t first_base, uint64_t total_bases, uint64_t first_read_index,
                               int32_t n_parts, void *d_tuples, uint64_t capacity,
                               uint64_t *d_counts, void *stream);
/* FUSED exchange over NVLink peer memory (one process per GPU): instead of packing tuples locally and handing them to
 * a collective, the scatter pass stores every tuple straight into its OWNER's receive buffer.
 *   dbg_peer_alloc / dbg_peer_open : a cudaMalloc'ed receive buffer + its 64-byte cudaIpcMemHandle; peers open it
 *   dbg_exchange_count_device      : pass 1 (count + scan); d_counts[b] = tuples this block has for exchange bucket b.
 *                                    by_slice = 0: one bucket per owner (n_parts buckets);
 *                                    by_slice = 1: one bucket per (owner, 16-MB table slice of the owner):
 *                                    n_parts * n_slices buckets (dbg_partition_info), owner-major
 *   (caller all-gathers the counts and derives dst_base[b] = start of bucket b inside its owner's buffer + the
 *    counts of lower ranks for b)
 *   dbg_exchange_scatter_device    : pass 2 on the SAME block; tuple of bucket b -> d_dst_ptrs[b][dst_base[b] + k]
 * The caller synchronises (stream sync + barrier) before the owners insert: dbg_insert_tuples_device for by_slice=0
 * (re-partitions by slice first), dbg_insert_sliced_device for by_slice=1 (the buffer already is in slice order). */
int  dbg_peer_alloc(dbg_ctx *ctx, uint64_t bytes, void **d_ptr, uint8_t handle[64]);
int  dbg_peer_open(dbg_ctx *ctx, const uint8_t handle[64], void **d_ptr);
int  dbg_peer_close(dbg_ctx *ctx, void *d_ptr);
int  dbg_peer_free(dbg_ctx *ctx, void *d_ptr);
int  dbg_exchange_count_device(dbg_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                               uint64_t first_base, uint64_t total_bases, int32_t n_parts, int32_t by_slice,
                               uint64_t *d_counts, void *stream);
int  dbg_exchange_scatter_device(dbg_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                                 uint64_t first_base, uint64_t total_bases, uint64_t first_read_index, int32_t n_parts,
                                 int32_t by_slice, void *const *d_dst_ptrs, const uint64_t *d_dst_base, void *stream);
int  dbg_insert_sliced_device(dbg_ctx *ctx, const void *d_tuples, uint64_t n, const uint64_t *d_slice_offs, void *stream);
/* OPTIMISTIC fused exchange (the default of the multi-GPU driver): ONE extraction pass.  Every owner keeps a fixed region
 * of cap_pair tuples for every source rank in its receive buffer; the source's scatter pass sorts every 2048-tuple batch
 * by owner in shared memory and stores the runs into its region of each owner's buffer over NVLink (d_dst_ptrs[q] = base
 * of owner q's buffer, region_off = first tuple of this source's region inside it), reserving space with atomics on its
 * LOCAL counters d_fill[0..n_parts).  d_fill[n_parts] != 0 afterwards: a region would have overflowed (skewed input) --
 * nothing was inserted, redo the block with dbg_exchange_count/scatter_device.  The owners then take the regions with
 * dbg_insert_tuple_regions_device (region r = counts[r] tuples at d_base + r * stride_tuples; counts are HOST values:
 * the all-gathered d_fill). */
int  dbg_exchange_scatter_opt_device(dbg_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads, uint64_t first_base,
                                     uint64_t total_bases, uint64_t first_read_index, int32_t n_parts, void *const *d_dst_ptrs,
                                     uint64_t region_off, uint32_t cap_pair, uint32_t *d_fill, void *stream);
/* PULL exchange (owner = slot range of hash_code(kmer) % P: the reference's `kmer % threadNum` owner-computes split of
 * thread_updatekmers, DBGgraph.cpp:148, lifted to GPUs): the source partitions its occurrences by (owner, table slice of
 * the owner) into its OWN send buffer
 * (bucket = owner * n_slices + slice, n_slices from dbg_partition_info; region of `capb` tuples per bucket, capb a multiple of
 * 512; d_fill has n_parts * n_slices + 1 counters, the last one the overflow flag) -- one extraction pass, local stores only.
 * After the fill counters have been all-gathered, every owner inserts its buckets slice by slice and READS the sources'
 * regions over NVLink peer mappings (d_src_ptrs[q]; d_fills + q * fill_stride = source q's counters): long contiguous reads,
 * no receive buffer and no owner-side partition pass.  n_tuples_upper: an upper bound of what this owner receives
 * (long-probe budget only). */
int  dbg_exchange_scatter_pull_device(dbg_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads, uint64_t first_base,
                                      uint64_t total_bases, uint64_t first_read_index, int32_t n_parts, void *d_send, uint32_t capb,
                                      uint32_t *d_fill, void *stream);
int  dbg_insert_pull_device(dbg_ctx *ctx, void *const *d_src_ptrs, int32_t n_src, uint32_t capb, const uint32_t *d_fills,
                            uint32_t fill_stride, uint64_t n_tuples_upper, void *stream);
int  dbg_exchange_scatter_undo(dbg_ctx *ctx, void *stream);     /* after an overflow: restore the side counters of the scatter */
int  dbg_insert_tuple_regions_device(dbg_ctx *ctx, const void *d_base, uint32_t n_regions, uint64_t stride_tuples,
                                     const uint64_t *counts, void *stream);
int  dbg_partition_info(const dbg_ctx *ctx, uint32_t *n_slices, int32_t *slice_shift);
/* Insert n tuples (all owned by this context's shard) produced by dbg_extract_tuples_device. */
int  dbg_insert_tuples_device(dbg_ctx *ctx, const void *d_tuples, uint64_t n, void *stream);
int  dbg_tuple_bytes(const dbg_ctx *ctx);               /* 16 or 32 */
/* add a peer's poly-A counters / read counters (tiny all-reduce done by the caller) */
int  dbg_get_polyA_counts(dbg_ctx *ctx, uint64_t counts[8]);
int  dbg_set_polyA_counts(dbg_ctx *ctx, const uint64_t counts[8]);

/* ---- cross-shard reference layout: the multi-GPU build as a drop-in ------------------------------------------
 * Owner = slot range (DBGgraph.cpp:148 lifted to GPUs), so the shards' tables concatenate into the reference's table --
 * except for probe clusters that cross a shard boundary, which must be replayed as one unit.
 * Protocol, on every rank r of the ring, after the last insert:
 *   1. dbg_shard_tail_export : the TAIL UNIT of r -- the run of occupied slots ending at r's last home slot plus the
 *                              nodes r's inserts pushed past it (they belong into r+1's first free slots) -- as a blob
 *                              of host bytes (blob == NULL: size query).  r excludes that run from its own layout.
 *   2. (caller moves the blob to rank (r+1) % n: torch.distributed, MPI, memcpy ...)
 *   3. dbg_shard_tail_import : r+1 keeps the run in the porch in front of its table and adopts the overflow nodes.
 *   4. dbg_finalize          : lays out [imported run | own range minus own tail run] in the reference's slot order.
 *   5. dbg_export_shard_slice: copies that slice into the caller's FULL table image array[P] (+ the nul_flag bytes it
 *                              covers completely; the bytes shared with a neighbour come back in edge_slots).
 *   6. once, by whoever holds the merged table: dbg_host_fix_nul_bytes(edge slots of all ranks), then
 *      dbg_host_polyA_insert (the k-mer-0 node goes in last, DBGgraph.cpp:418; link words = stats.polyA_l/_r after the
 *      side counters were summed over the ranks).
 * DBG_ERR_STATE from 1/3/4 means the boundary cluster does not fit the hand-off (tables of a few thousand slots, or
 * pathologically dense ones): merge the dbg_dump_shard outputs with dbg_replay_growth instead (same layout). */
int  dbg_shard_tail_export(dbg_ctx *ctx, void *blob, uint64_t cap_bytes, uint64_t *n_bytes);
int  dbg_shard_tail_import(dbg_ctx *ctx, const void *blob, uint64_t n_bytes);
/* the laid-out slice: first global slot, length (it may wrap past slot P-1), device pointer (reference node size) */
int  dbg_shard_slice_info(dbg_ctx *ctx, uint64_t *g_first, uint64_t *n_slots, void **d_slice);
int  dbg_export_shard_slice(dbg_ctx *ctx, void *array, uint8_t *nul_flag, uint64_t edge_slots[4]);
int  dbg_host_fix_nul_bytes(const void *array, uint8_t *nul_flag, uint64_t P, int32_t wide, const uint64_t *slots, uint64_t n);
int  dbg_host_polyA_insert(void *array, uint8_t *nul_flag, uint64_t P, int32_t wide, uint32_t l_link, uint32_t r_link, uint64_t *slot_out);

/* ---- ONE process, several GPUs: the same five calls a front end makes for one GPU ------------------------------------
 * dbg_mg_* drives n sharded contexts (devices[r], NULL = 0..n-1; peer access is enabled between them) from one host
 * process: reads of a block are dealt to the GPUs, exchanged with the optimistic fused scatter over NVLink, inserted by
 * their owners; dbg_mg_finalize hands the boundary clusters around the ring and lays out every slice;
 * dbg_mg_export_kmerset assembles the reference's KmerSet (kmerSet.h:88-99: array[P] + nul_flag, k-mer-0 node last) in
 * the caller's memory -- what build_debruijn_graph hands to build_contig_sequence (main.cpp:204-207).  params->device,
 * shard_rank and shard_count are ignored.  dbg_mg_dump_nodes lists every node with its first-occurrence ordinal (input
 * of dbg_replay_growth, like dbg_dump_shard).  info = {exchange rounds, region regrows, dump-merge fallback used, region
 * size in tuples}. */
typedef struct dbg_mg dbg_mg;
int  dbg_mg_create(dbg_mg **mg, const dbg_params *params, int32_t n_gpus, const int32_t *devices);
void dbg_mg_destroy(dbg_mg *mg);
int  dbg_mg_submit_reads(dbg_mg *mg, const char *bases, const uint64_t *offs, uint64_t n_reads);
int  dbg_mg_finalize(dbg_mg *mg, dbg_stats *stats);
int  dbg_mg_get_stats(dbg_mg *mg, dbg_stats *stats);
int  dbg_mg_export_kmerset(dbg_mg *mg, void *array, uint8_t *nul_flag);
int  dbg_mg_dump_nodes(dbg_mg *mg, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *l_link, uint32_t *r_link,
                       uint64_t *first_ordinal, uint64_t *n);
int  dbg_mg_info(const dbg_mg *mg, uint64_t info[4]);
const char *dbg_mg_last_error(void);

/* ---- results ---------------------------------------------------------------------------------- */
/* Tail of build_debruijn_graph: waits for all blocks, builds the reference-layout image on the device
 * and appends the k-mer-0 node last (DBGgraph.cpp:418).  Fills *stats (may be NULL). */
int  dbg_finalize(dbg_ctx *ctx, dbg_stats *stats);
int  dbg_get_stats(dbg_ctx *ctx, dbg_stats *stats);

/* The KmerSet the reference's traversal consumes (kmerSet.h:88-99): array[P] in reference slot
 * layout (16-B nodes, or 32-B nodes on the wide path) and nul_flag[P/8+1], MSB first. */
int  dbg_export_kmerset(dbg_ctx *ctx, void *array, uint8_t *nul_flag);
/* build_debruijn_graph's tail (DBGgraph.cpp:383-430: the last parse_one_reads_file block, add_node_to_kmerset(PolyA), and the
 * hand-over of `kset` to build_contig_sequence, main.cpp:204-207) in ONE call, for a front end that holds the LAST block of
 * reads in host memory:
 *   dbg_submit_reads(bases, offs, n_reads) + dbg_finalize(stats) + dbg_export_kmerset(array, nul_flag)
 * -- same results, but pipelined where the block is large enough for the partitioned build: the copy of the reads, the
 * extraction, and then insert / reference layout / copy of the image run slice group by slice group, so the image is on
 * the PCIe link (the bound of the hand-over) while later groups are still being inserted.  n_reads may be 0 (finalize +
 * export).  `array` should be pinned (dbg_host_alloc / dbg_host_register) for the copies to overlap.
 * Environment: DBG_B200_FINISH_GROUPS (default 8; < 2 = the plain sequence). */
int  dbg_finish_export(dbg_ctx *ctx, const char *bases, const uint64_t *offs, uint64_t n_reads, dbg_stats *stats,
                       void *array, uint8_t *nul_flag);

/* How the last dbg_export_kmerset moved the table: info = {chunks sent compact, chunks sent plain, bytes moved over the
 * link, occupied nodes}.  Default: ONE plain copy of the image ({0, 1, P*node + P/8+1, count}).  With the environment
 * variable DBG_B200_EXPORT=pipe, tables of more than a few million slots travel as occupied nodes only (compacted on the
 * device in slot order, chunk by chunk) and host threads of the library expand them into `array` while later chunks are on
 * the link; when the host threads fall behind, a chunk travels as plain image bytes instead (only if `array` is pinned:
 * dbg_host_alloc / dbg_host_register).  Same bytes in `array` / `nul_flag` either way.  It pays on hosts whose memory
 * bandwidth is several times the PCIe link's; on the measured 16-vCPU hosts the plain copy is faster, hence opt-in.
 * Knobs: DBG_B200_EXPORT_THREADS, DBG_B200_EXPORT_CHUNK (slots), DBG_B200_EXPORT_SLOTS (pinned ring slots),
 * DBG_B200_EXPORT_NO_DIRECT, DBG_B200_EXPORT_PLAIN_PCT. */
int  dbg_export_info(const dbg_ctx *ctx, uint64_t info[4]);
/* host helper of that path: array[s] = (bit s of `bits`, MSB first as in nul_flag) ? next node of `nodes` : 0 for
 * n_slots slots; returns the nodes consumed.  `nodes` must be readable 16 bytes past its last node. */
uint64_t dbg_host_expand_nodes(const uint8_t *bits, uint64_t n_slots, const void *nodes, void *array, int32_t wide);

/* calculate_kmer_links (contig.cpp:107-205) on the device: klink[P*2] (KmerLink bit layout,
 * contig.h:31-42), del_flag[P/8+1], depth_hist[256], index-ordered tip and branch lists (capacity
 * *n_tips / *n_branches on input, lengths on output), stats3 = {total, deleted, linear}. */
int  dbg_export_links(dbg_ctx *ctx, int32_t freq_cutoff, uint8_t *klink, uint8_t *del_flag,
                      int64_t depth_hist[256], uint64_t *tips, uint64_t *n_tips,
                      uint64_t *branches, uint64_t *n_branches, int64_t stats3[3]);

/* Stream-compacted dump, slot order: nodes that survive the low-frequency link filter
 * (freq_cutoff < 0: every filled slot).  Any output pointer may be NULL; *n = capacity in, count out. */
int  dbg_dump_compact(dbg_ctx *ctx, int32_t freq_cutoff, uint64_t *slots, uint64_t *kmers_lo,
                      uint64_t *kmers_hi, uint32_t *l_link, uint32_t *r_link, uint64_t *n);

/* Unordered dump of THIS context's build table (any context, sharded or not; excludes the k-mer-0 side
 * node): nodes with link words already clamped to the reference's 8-bit lanes, plus each node's
 * first-occurrence ordinal (read_index << 16 | position) when track_order is on.  *n = capacity in,
 * count out; all-NULL output pointers = size query. */
int  dbg_dump_shard(dbg_ctx *ctx, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *l_link, uint32_t *r_link,
                    uint64_t *first_ordinal, uint64_t *n);

/* device pointers of the finalized image, for callers that stay on the GPU (bench, multi-GPU gather) */
int  dbg_device_image(dbg_ctx *ctx, void **d_array, void **d_nul_flag);
/* the BUILD table behind it (n_local nodes of 32 B {kmer, ~first ordinal, 8 half-float counts}, 64 B on the wide path) and the
 * 8 side counters of the k-mer-0 node; any pointer may be NULL */
int  dbg_device_build_table(dbg_ctx *ctx, void **d_nodes, uint64_t *n_local, uint64_t **d_polyA);

/* per-phase device time of the most recent calls, milliseconds (CUDA events on the ctx stream):
 * [0] table clear, [1] build kernels (sum over blocks), [2] layout+polyA (finalize), [3] links pass,
 * [4] H2D copies, [5] D2H export, [6] the bucketed insert kernel alone (partitioned path; part of [1]),
 * [7] the scatter kernel of the optimistic peer exchange alone (multi-GPU) */
int  dbg_get_timings(dbg_ctx *ctx, float ms[8]);
/* number of kernel launches issued by this context so far */
uint64_t dbg_launch_count(const dbg_ctx *ctx);
/* how many read blocks went through which build path so far: [0] fused direct insert, [1] exact two-pass
 * partition (count, scan, scatter), [2] optimistic single-pass partition (fixed bucket regions), [3] optimistic
 * attempts that overflowed a region and were redone exactly (also counted in [1]) */
int  dbg_path_counts(const dbg_ctx *ctx, uint64_t counts[4]);
/* re-zero the table and counters so the context can build again (bench steps) */
int  dbg_reset(dbg_ctx *ctx);
/* run all of this context's kernels, memsets and copies on a caller-owned cudaStream_t (e.g. torch's
 * current stream, so that CUDA events recorded there bracket the work); NULL restores the own stream.
 * NB a NULL stream argument anywhere in this API means "the context's stream": to name the legacy default
 * stream pass cudaStreamLegacy. */
int  dbg_set_stream(dbg_ctx *ctx, void *stream);

/* ---- synthetic reads: counter-based, identical on host and device ----------------- */
typedef struct {
    uint64_t seed;
    uint64_t genome_len;
    uint32_t read_len;       /* every read has this length                                    */
    uint32_t insert;         /* paired-end fragment length (mate 1 = reverse strand of its end) */
    uint32_t err_per_2p24;   /* substitution probability * 2^24                                */
    uint32_t n_per_2p24;     /* 'N' probability * 2^24                                         */
} dbg_synth_params;
int  dbg_synth_reads_host(const dbg_synth_params *p, uint64_t first_read, uint64_t n_reads, char *out);
int  dbg_synth_reads_device(const dbg_synth_params *p, uint64_t first_read, uint64_t n_reads, char *d_out,
                            int32_t device, void *stream);

/* ---- host-side replay of the reference's table growth (-e / enlarge) ------------
 * The GPU table is sized once from -i.  When a run outgrows it (count > max at a block boundary of -b reads), the
 * reference doubles its table in place (DBGgraph.cpp:337-351, kmerSet.cpp:132-189): same nodes, different slot layout.
 * dbg_replay_growth rebuilds that layout on the HOST from the finished graph: the nodes with their first-occurrence
 * ordinals (dbg_dump_shard on an unsharded context; the k-mer-0 node excluded, its link words passed separately) and
 * the number of reads in every input file.  Call it with array == NULL first: `res` tells whether the reference would
 * have grown (doublings > 0), the final table size to allocate, or `truncated` (-e exhausted: the reference drops the
 * rest of a file, which no replay of a full build can reproduce -- nothing is written).  The plan call is one pass over
 * the ordinals; the layout call is sequential like the code it mirrors (a few hundred ns per node). */
typedef struct {
    uint64_t init_slots;        /* (uint64)(-i * 1e9), as in dbg_params                     */
    float    load_factor;       /* -l                                                       */
    int32_t  wide;              /* 1: 128-bit keys (kmer_hi given, 32-byte export nodes)    */
    uint64_t max_double_times;  /* -e                                                       */
    uint64_t buffer_reads;      /* -b  reads per block (grow check after every full block)  */
} dbg_growth_params;
typedef struct {
    uint64_t final_size, final_max, doublings, count;   /* KmerSet.size / .max / doubleHashTimes / .count (incl. k-mer 0) */
    int32_t  truncated;                                   /* 1: "Memory reach the maximum allowed" would have hit          */
    uint32_t truncated_file;                              /*    in this file (index into reads_per_file) ...               */
    uint64_t truncated_first_read;                        /*    ... from this global read index on                         */
} dbg_growth_result;
int  dbg_replay_growth(const dbg_growth_params *g, const uint64_t *reads_per_file, uint32_t n_files,
                       const uint64_t *kmer_lo, const uint64_t *kmer_hi, const uint32_t *l_link, const uint32_t *r_link,
                       const uint64_t *first_ordinal, uint64_t n_nodes, uint32_t polyA_l, uint32_t polyA_r,
                       dbg_growth_result *res, void *array, uint8_t *nul_flag);

/* ---- on-disk checkpoint of the finished graph (host code, plain file I/O) --------------------------
 * The KmerSet build_debruijn_graph hands to the traversal (kmerSet.h:88-99), stored as its filled slots in slot order
 * {slot, kmer[, kmer_hi], l_link, r_link} + a checksum: a front end can re-run the traversal with other cut-offs without
 * reading the reads again (integration/DBGgraph_b200.cpp: DBG_B200_CHECKPOINT=<file>).  magic / version / records are
 * filled in by the writer. */
typedef struct {
    uint64_t magic; uint32_t version; uint32_t K; uint32_t wide; float load_factor;
    uint64_t size, max_cutoff, count, count_conflict, reads, kmers_logged, records;
} dbg_checkpoint_header;
int  dbg_checkpoint_write(const char *path, const dbg_checkpoint_header *hdr, const void *array, const uint8_t *nul_flag);
int  dbg_checkpoint_read_header(const char *path, dbg_checkpoint_header *hdr);
int  dbg_checkpoint_read(const char *path, void *array, uint8_t *nul_flag);

/* ---- K-mer frequency table for correct_error ---------------------------------
 * What the external `kmerfreq` program writes and correct_error loads (correct_error/main_parallel_senior.cpp:
 * 273-408 1-bit form, correct_error/main.cpp:161-220 8-bit form): canonical k-mer counts in a direct-index table
 * (K <= 17), written as <prefix>.kmer.freq.cz (zlib blocks of 8 Mi k-mers), .cz.len and .kmer.freq.stat.
 * PARITY UNPINNED: kmerfreq itself is not in the reference tree; the format contract is the loaders'. */
typedef struct kfreq_ctx kfreq_ctx;
/* block_rank/block_count: this context owns a contiguous run of whole .cz blocks (multi-GPU); 0/1 = everything */
int  kfreq_create(kfreq_ctx **ctx, int32_t K, int32_t device, int32_t block_rank, int32_t block_count);
void kfreq_destroy(kfreq_ctx *ctx);
int  kfreq_submit_reads(kfreq_ctx *ctx, const char *bases, const uint64_t *offs, uint64_t n_reads);
int  kfreq_submit_reads_device(kfreq_ctx *ctx, const char *d_bases, const uint64_t *d_offs, uint64_t n_reads,
                               uint64_t first_base, uint64_t total_bases);
int  kfreq_finalize(kfreq_ctx *ctx, uint64_t *n_occurrences, uint64_t *n_reads);
int  kfreq_reset(kfreq_ctx *ctx);     /* zero the table and the counters: a fresh count on the same context */
int  kfreq_index_range(kfreq_ctx *ctx, uint64_t *lo, uint64_t *hi);
/* hist[f] = species seen f times (f = 65535: that often or more), over the owned index range */
int  kfreq_histogram(kfreq_ctx *ctx, uint64_t hist[65536]);
/* raw image of the owned range into host memory: bits=1 -> (hi-lo)/8 bytes, bit 7-idx%8 of byte idx/8 set iff
 * count > cutoff; bits=8 -> one byte min(255,count) per k-mer */
int  kfreq_export(kfreq_ctx *ctx, int32_t bits, int32_t cutoff, uint8_t *out);
/* <prefix>.kmer.freq.cz, <prefix>.kmer.freq.cz.len, and (when the context owns all of 4^K) <prefix>.kmer.freq.stat */
int  kfreq_write_cz(kfreq_ctx *ctx, const char *prefix, int32_t bits, int32_t cutoff);
const char *kfreq_last_error(void);

/* ---- contig seed index of link_scaffold (csrc/seedidx.cu) -----------------------------------------
 * The k-mer -> {contig id, position, unique?, strand} hash that map_pair / map_reads build from the contigs
 * (link_scaffold/map_pair.cpp:122-125: init_kmerset(3 x contig length, 0.5) + chop_contig_to_kmerset, map_func.cpp:119-172,
 * kmerSet.cpp:82-107,168-210) and probe with read k-mers (get_align_seed, map_func.cpp:181-237).  The exported table is the
 * reference's KmerSet byte for byte (same size, same slots): nodes {u64 kmer; u64 id:32, pos:30, freq:1, direct:1}
 * (kmerSet.h:53-60, GCC bit-field layout: value = id | pos << 32 | freq << 62 | direct << 63) + nul_flag[size/8+1], MSB first.
 * K <= 31.  Contigs are ASCII [ACGTNacgtn]; runs of upper-case 'N' separate blocks (scaffold_to_contig), blocks shorter than
 * K are skipped (the reference's loop is undefined there).  A table that the reference would enlarge (count >= size *
 * load_factor) is refused with DBG_ERR_TABLE_FULL. */
typedef struct seedidx_ctx seedidx_ctx;
typedef struct { uint64_t kmer; uint64_t value; } seed_node16;
int  seedidx_create(seedidx_ctx **ctx, int32_t K, uint64_t init_slots, float load_factor, int32_t device);
void seedidx_destroy(seedidx_ctx *ctx);
/* contigs seqs[offs[i] .. offs[i+1]); contig ids continue across calls; an empty sequence keeps its id (map_pair.cpp:100-110) */
int  seedidx_add_contigs(seedidx_ctx *ctx, const char *seqs, const uint64_t *offs, uint64_t n_contigs);
int  seedidx_finalize(seedidx_ctx *ctx, uint64_t *size, uint64_t *count, uint64_t *max_cutoff);
int  seedidx_export(seedidx_ctx *ctx, void *array, uint8_t *nul_flag);
/* get_align_seed(read, search_start[i] (NULL: 1), read length) for a batch of reads (each <= 65535 bases):
 * out[6 i ..] = {contig_id_index, seed_contig_start, seed_contig_end, seed_read_start, seed_read_end, 'F' | 'R' | 'N'},
 * the first five -1 when no seed is found */
int  seedidx_align_reads(seedidx_ctx *ctx, const char *bases, const uint64_t *offs, uint64_t n_reads, const int32_t *search_start,
                         int32_t seed_kmer_num, int32_t *out);
uint64_t seedidx_launch_count(const seedidx_ctx *ctx);
const char *seedidx_last_error(void);

/* ---- roofline denominators measured on the spot (bench.py) --------------------------------------- */
/* uniformly random 32-B sector read-modify-writes over `bytes` of device memory, `n_ops` operations;
 * returns milliseconds (CUDA events).  mode 0: 32-B load + store, 1: 32-B load + 64-bit atomicCAS,
 * 2: u32 RED only, 3: f16x8 vector RED only, 4: 32-B load only, 5: 32-B load + f16x8 RED (the insert kernel's
 * hit path).  With `bytes` below the L2 size the same call measures the on-chip (L2) transaction rates. */
int  dbg_measure_random_rmw(int32_t device, uint64_t bytes, uint64_t n_ops, int32_t mode, float *ms);

#ifdef __cplusplus
}
#endif
#endif
