"""Multi-GPU build: one process per GPU, k-mers partitioned by owner = (hash_code(kmer) % P) / ceil(P/n)
-- the reference's `kmer % threadNum` owner-computes split (DBGgraph.cpp:148) lifted to GPUs.  Every rank
extracts the (k-mer, left, right, ordinal) tuples of its own reads (dbg_extract_tuples_device), buckets them
by owner, exchanges them with ONE all-to-all over NCCL/NVLink (sizes first, then payload) and inserts what
it received into its private shard (dbg_insert_tuples_device).  No further communication until export; the
k-mer-0 side counters are summed with one tiny all-reduce.

torch.distributed is plumbing here: the kernels are the C-ABI library's.  The exchange itself is
device-agnostic (`Exchange`), so world_size-2 gloo tests on CPU cover the host logic with stand-in
extract/insert callables.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def shard_size(P: int, n: int) -> int:
    return (P + n - 1) // n


def owner_of(home_slot, P: int, n: int):
    """owner rank of a home slot (same formula as BucketSink in dbg_kernels.cuh)"""
    return home_slot // shard_size(P, n)


# ---------------------------------------------------------------------------------------------------
# cross-shard reference layout: transport-agnostic pieces (include/dbg_b200.h, dbg_shard_tail_export ...)
# ---------------------------------------------------------------------------------------------------
def blob_margin_nodes(blob: bytes):
    """the overflow (margin) nodes inside a tail blob as a dump dict (kmer, kmer_hi, l, r, ord): what a rank handed to
    its right neighbour.  Only the dump-merging fallback needs it."""
    a, mt, nbb, _ = np.frombuffer(blob[:32], dtype=np.uint64).tolist()
    raw = np.frombuffer(blob, dtype=np.uint8, offset=32 + a * nbb, count=mt * nbb).reshape(mt, nbb)
    w = raw.view(np.uint64).reshape(mt, nbb // 8)
    if nbb == 32:
        klo, khi, nord, c = w[:, 0], np.zeros(mt, np.uint64), w[:, 1], raw[:, 16:32]
    else:
        klo, khi, nord, c = w[:, 0], w[:, 1], w[:, 2], raw[:, 32:48]
    cnt = np.minimum(np.ascontiguousarray(c).view(np.float16).reshape(mt, 8).astype(np.uint32), 255)
    link = lambda q: (q[:, 0] << 24) | (q[:, 1] << 16) | (q[:, 2] << 8) | q[:, 3]      # lane of base b = byte 3-b
    return dict(kmer=klo.copy(), kmer_hi=khi.copy(), l=link(cnt[:, :4]).astype(np.uint32), r=link(cnt[:, 4:]).astype(np.uint32),
                ord=~nord)


def merge_dumps_host(dumps, P_request, load_factor, wide, polyA_l, polyA_r):
    """Fallback of the windowed cross-shard layout (boundary cluster too long for the hand-off: tiny or pathologically
    dense tables): rebuild the reference's table from the union of the shard dumps by replaying the keys in
    first-occurrence order on the host (dbg_replay_growth with the growth checks switched off)."""
    from .graph import replay_growth
    cat = {k: np.concatenate([d[k] for d in dumps]) for k in ("kmer", "kmer_hi", "l", "r", "ord")}
    # a node can show up twice (adopted by the right neighbour AND listed from the blob): keep one
    key = np.stack([cat["kmer_hi"], cat["kmer"]], axis=1)
    _, first = np.unique(key, axis=0, return_index=True)
    cat = {k: v[np.sort(first)] for k, v in cat.items()}
    cat["wide"] = bool(wide)
    n_reads = int(cat["ord"].max() >> np.uint64(16)) + 1 if len(cat["ord"]) else 1
    plan, arr, nul = replay_growth(cat, [n_reads], P_request, load_factor, max_double_times=0, buffer_reads=1 << 62,
                                   polyA_l=polyA_l, polyA_r=polyA_r)
    return arr, nul, plan


def ring_layout(builders, blobs_out=None):
    """single process, all ranks at hand (tests, one process driving several GPUs): tail hand-off around the ring, then
    every rank lays out its window.  Returns (list of stats, blobs).  blobs_out (a list) receives every blob as soon as
    it exists: a rank that refuses (DBG_ERR_STATE) leaves the caller with what the others already handed over -- their
    overflow nodes now live ONLY in those blobs -- for the dump-merging fallback."""
    n = len(builders)
    blobs = blobs_out if blobs_out is not None else []
    for b in builders:
        blobs.append(b.shard_tail_export())
    for r, b in enumerate(builders):
        b.shard_tail_import(blobs[(r - 1) % n])
    return [b.finalize() for b in builders], blobs


def export_merged(builders, stats, array=None, nul_flag=None):
    """slices of all ranks -> ONE host table in the reference's layout (array[P], nul_flag[P/8+1]) + the k-mer-0 node"""
    from .graph import NODE16, NODE32
    from . import capi
    P, wide = stats[0]["array_size"], bool(stats[0]["wide"])
    if array is None:
        array = np.zeros(P, dtype=NODE32 if wide else NODE16)
    if nul_flag is None:
        nul_flag = np.zeros(P // 8 + 1, dtype=np.uint8)
    edges = []
    for b in builders:
        edges += b.export_shard_slice(array.ctypes.data, nul_flag.ctypes.data)
    capi.host_fix_nul_bytes(array, nul_flag, P, wide, edges)
    capi.host_polyA_insert(array, nul_flag, P, wide, stats[0]["polyA_l"], stats[0]["polyA_r"])
    return array, nul_flag


class LocalShards:
    """N sharded contexts driven by ONE process (all on `devices[r]`; the same device may repeat): the complete
    multi-GPU flow -- fused peer exchange (count, exact offsets, scatter into the owners' receive buffers), owner-side
    insert, k-mer-0 counters summed, cross-shard layout, merged export -- without torch.distributed.  What the GPU
    tests run on a single device, and what a single-process front end would run on several."""

    def __init__(self, n, K, max_read_len, init_slots, load_factor=0.7, devices=None, track_order=True, force_wide=False,
                 by_slice=False, optimistic=False, cap_pair=None, pull=False):
        import torch
        from .graph import DBGBuilder
        self.n = n
        self.devices = devices or [0] * n
        self.b = [DBGBuilder(K=K, max_read_len=max_read_len, init_slots=init_slots, load_factor=load_factor, device=self.devices[r],
                             track_order=track_order, shard_rank=r, shard_count=n, force_wide=force_wide) for r in range(n)]
        self.tb = self.b[0].tuple_bytes
        self.by_slice = by_slice
        self.torch = torch
        self.P_request, self.load_factor = init_slots, load_factor
        self.blobs = None
        self.optimistic, self.cap_pair = optimistic, cap_pair
        self.pull = pull           # PULL exchange: sources partition by (owner, slice) locally, owners read the regions (cap_pair = capb)
        self.overflows = 0

    def close(self):
        for b in self.b:
            b.close()

    def _add_blocks_opt(self, blocks):
        """one round of the OPTIMISTIC exchange: every source scatters straight into its region of every owner's buffer
        (dbg_exchange_scatter_opt_device), the owners take the regions (dbg_insert_tuple_regions_device); a region that
        would overflow makes the whole round fall back to the exact exchange (nothing was inserted yet)"""
        torch, n = self.torch, self.n
        biggest = max([blk[4] for blk in blocks if blk is not None] + [1])
        cap_pair = self.cap_pair or int(biggest / n * 1.25) + 4096
        recv = [torch.zeros(n * cap_pair * (self.tb // 8), dtype=torch.int64, device=torch.device("cuda", self.devices[q])) for q in range(n)]
        fills = []
        for r, blk in enumerate(blocks):
            dev = torch.device("cuda", self.devices[r])
            fill = torch.zeros(n + 1, dtype=torch.int32, device=dev)
            if blk is not None:
                db, do, nr, fb, tbases, fri = blk
                d_ptrs = torch.tensor([t.data_ptr() for t in recv], dtype=torch.int64, device=dev)
                self.b[r].exchange_scatter_opt_device(db.data_ptr(), do.data_ptr(), nr, fb, tbases, fri, n, d_ptrs.data_ptr(), r * cap_pair,
                                                      cap_pair, fill.data_ptr())
                torch.cuda.synchronize(self.devices[r])
            fills.append(fill.cpu())
        allf = torch.stack(fills)                                       # [source][owner | flag]
        if int(allf[:, n].sum()) != 0:
            self.overflows += 1
            for r, blk in enumerate(blocks):
                if blk is not None:
                    self.b[r].exchange_scatter_undo()
                    torch.cuda.synchronize(self.devices[r])
            return None
        for q in range(n):
            counts = allf[:, q].numpy().astype(np.uint64)
            if counts.sum():
                self.b[q].insert_tuple_regions_device(recv[q].data_ptr(), cap_pair, counts)
                torch.cuda.synchronize(self.devices[q])
        return allf[:, :n].sum(dim=0).tolist()

    def _add_blocks_pull(self, blocks):
        """one round of the PULL exchange (dbg_exchange_scatter_pull_device / dbg_insert_pull_device); None = a region would
        have overflowed, nothing was inserted and the side counters were restored"""
        torch, n = self.torch, self.n
        nbl = self.b[0].partition_info()[0]
        nbt = n * nbl
        biggest = max([blk[4] for blk in blocks if blk is not None] + [1])
        capb = self.cap_pair or int(biggest / nbt * 1.25 + 6 * (biggest / nbt) ** 0.5) + 512
        capb = (capb + 511) // 512 * 512
        send = [torch.zeros(nbt * capb * (self.tb // 8), dtype=torch.int64, device=torch.device("cuda", self.devices[r])) for r in range(n)]
        fills = []
        for r, blk in enumerate(blocks):
            fill = torch.zeros(nbt + 1, dtype=torch.int32, device=torch.device("cuda", self.devices[r]))
            if blk is not None:
                db, do, nr, fb, tbases, fri = blk
                self.b[r].exchange_scatter_pull_device(db.data_ptr(), do.data_ptr(), nr, fb, tbases, fri, n, send[r].data_ptr(), capb, fill.data_ptr())
                torch.cuda.synchronize(self.devices[r])
            fills.append(fill.cpu())
        allf = torch.stack(fills)                                       # [source][bucket | flag]
        if int(allf[:, nbt].sum()) != 0:
            self.overflows += 1
            for r, blk in enumerate(blocks):
                if blk is not None:
                    self.b[r].exchange_scatter_undo()
                    torch.cuda.synchronize(self.devices[r])
            return None
        got = []
        for q in range(n):
            dev = torch.device("cuda", self.devices[q])
            d_src = torch.tensor([t.data_ptr() for t in send], dtype=torch.int64, device=dev)
            d_fills = allf.to(dev).contiguous()
            total = int(allf[:, q * nbl:(q + 1) * nbl].sum())
            if total:
                self.b[q].insert_pull_device(d_src.data_ptr(), n, capb, d_fills.data_ptr(), nbt + 1, total)
                torch.cuda.synchronize(self.devices[q])
            got.append(total)
        return got

    def add_blocks(self, blocks):
        """blocks[r] = (d_bases tensor, d_offs tensor, n_reads, first_base, total_bases, first_read_index): the reads rank r
        contributes to this round (None = nothing).  One exchange round: count on every rank, offsets, scatter, insert."""
        torch, n = self.torch, self.n
        if self.pull:
            got = self._add_blocks_pull(blocks)
            if got is not None:
                return got
        elif self.optimistic:
            got = self._add_blocks_opt(blocks)
            if got is not None:
                return got
            # a region would have overflowed (skewed input): the side counters were restored, redo the round exactly
        nbl = self.b[0].partition_info()[0] if self.by_slice else 1
        nbt = n * nbl
        counts = []
        for r, blk in enumerate(blocks):
            dev = torch.device("cuda", self.devices[r])
            c = torch.zeros(nbt, dtype=torch.int64, device=dev)
            if blk is not None:
                db, do, nr, fb, tbases, fri = blk
                self.b[r].exchange_count_device(db.data_ptr(), do.data_ptr(), nr, fb, tbases, n, c.data_ptr(), by_slice=self.by_slice)
            counts.append(c)
        for r in range(n):
            torch.cuda.synchronize(self.devices[r])
        allc = torch.stack([c.cpu() for c in counts])                      # [source][bucket]
        col = allc.sum(dim=0).view(n, nbl)
        start = torch.cumsum(col, dim=1) - col
        recv_total = col.sum(dim=1).tolist()
        recv = [torch.zeros(max(int(recv_total[q]), 1) * (self.tb // 8), dtype=torch.int64, device=torch.device("cuda", self.devices[q]))
                for q in range(n)]
        ptrs = [t.data_ptr() for t in recv]
        for r, blk in enumerate(blocks):
            if blk is None:
                continue
            dev = torch.device("cuda", self.devices[r])
            db, do, nr, fb, tbases, fri = blk
            d_base = (start.reshape(-1) + allc[:r].sum(dim=0)).contiguous().to(dev)
            d_ptrs = torch.tensor([ptrs[q] for q in range(n) for _ in range(nbl)], dtype=torch.int64, device=dev)
            self.b[r].exchange_scatter_device(db.data_ptr(), do.data_ptr(), nr, fb, tbases, fri, n, d_ptrs.data_ptr(), d_base.data_ptr(),
                                              by_slice=self.by_slice)
            torch.cuda.synchronize(self.devices[r])
        for q in range(n):
            if recv_total[q] == 0:
                continue
            if self.by_slice:
                so = torch.cat([start[q], col[q].sum().reshape(1)]).contiguous().to(recv[q].device)
                self.b[q].insert_sliced_device(recv[q].data_ptr(), int(recv_total[q]), so.data_ptr())
            else:
                self.b[q].insert_tuples_device(recv[q].data_ptr(), int(recv_total[q]))
            torch.cuda.synchronize(self.devices[q])
        return recv_total

    def finalize(self, layout=True):
        polyA = sum(b.get_polyA_counts() for b in self.b)
        for b in self.b:
            b.set_polyA_counts(polyA)
        if not layout:
            return [b.finalize() for b in self.b]
        self.blobs = []
        stats, _ = ring_layout(self.b, self.blobs)
        return stats

    def export_kmerset(self, stats):
        return export_merged(self.b, stats)

    def export_kmerset_fallback(self, stats=None):
        """merge the shard dumps on the host (used when the windowed layout refuses: DbgError with DBG_ERR_STATE)"""
        dumps = [b.dump_shard() for b in self.b]
        if self.blobs:
            dumps += [blob_margin_nodes(bl) for bl in self.blobs]
        polyA = self.b[0].get_polyA_counts()
        cl = lambda q: sum(int(min(int(q[i]), 255)) << (24 - 8 * i) for i in range(4))
        return merge_dumps_host(dumps, self.P_request, self.load_factor, self.b[0].wide, cl(polyA[:4]), cl(polyA[4:]))


class LayoutNeedsMerge(RuntimeError):
    """the windowed cross-shard layout refused on some rank (boundary cluster longer than the hand-off porch, a hand-off
    that cascades through a whole shard, a shard too dense for the cluster scratch: tables of a few thousand slots or
    pathologically dense ones).  Raised on EVERY rank together; merge the shard dumps instead (same table)."""


class SharedImage:
    """ONE host table image (array[P] + nul_flag[P/8+1], the reference's KmerSet, kmerSet.h:88-99) that the ranks of a
    one-process-per-GPU build export their slices into: a file mapping in /dev/shm (rank 0 creates it, everybody maps it).
    A single consumer process -- the reference's traversal -- would map the same file."""

    def __init__(self, path, P, wide, create):
        import mmap
        from .graph import NODE16, NODE32
        self.path, self.P, self.wide = path, int(P), bool(wide)
        nb = 32 if wide else 16
        self.img_bytes, self.nul_bytes = self.P * nb, self.P // 8 + 1
        self.nul_off = (self.img_bytes + 4095) // 4096 * 4096
        self.total = self.nul_off + (self.nul_bytes + 4095) // 4096 * 4096
        if create:
            with open(path, "wb") as f:
                f.truncate(self.total)
        self.f = open(path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), self.total)
        self.base = np.frombuffer(self.mm, dtype=np.uint8)
        self.base_ptr = self.base.ctypes.data
        self.arr = np.frombuffer(self.mm, dtype=NODE32 if wide else NODE16, count=self.P)
        self.nul = np.frombuffer(self.mm, dtype=np.uint8, count=self.nul_bytes, offset=self.nul_off)
        self.arr_ptr, self.nul_ptr = self.base_ptr, self.base_ptr + self.nul_off

    @staticmethod
    def room_for(P, wide, where="/dev/shm"):
        try:
            sv = os.statvfs(where)
            return sv.f_bavail * sv.f_frsize > P * (32 if wide else 16) + P // 8 + (1 << 30)
        except OSError:
            return False

    def close(self, unlink=False):
        self.arr = self.nul = self.base = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.f.close()
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class Exchange:
    """all-to-all(v) of fixed-width tuples: sizes first, then payload."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def exchange_counts(self, send_counts: torch.Tensor) -> torch.Tensor:
        """send_counts[q] = tuples this rank has for rank q -> recv_counts[q] = tuples rank q has for us"""
        recv = torch.empty_like(send_counts)
        dist.all_to_all_single(recv, send_counts, group=self.group)
        return recv

    def exchange_payload(self, buckets, recv_counts_host, width: int, like: torch.Tensor):
        """buckets[q]: [c_q, width] tensor for rank q.  Returns (recv [sum, width], offsets list)"""
        total = int(sum(recv_counts_host))
        recv = torch.empty((max(total, 1), width), dtype=like.dtype, device=like.device)
        outs, off = [], 0
        for c in recv_counts_host:
            outs.append(recv[off:off + int(c)])
            off += int(c)
        # one send + one recv per peer inside ONE group (ncclGroupStart/End under NCCL == all-to-all(v);
        # also what gloo offers, so the CPU tests run the same code)
        outs[self.rank].copy_(buckets[self.rank])
        ops = []
        for q in range(self.world):
            if q == self.rank:
                continue
            if buckets[q].shape[0]:
                ops.append(dist.P2POp(dist.isend, buckets[q].contiguous(), q, group=self.group))
            if outs[q].shape[0]:
                ops.append(dist.P2POp(dist.irecv, outs[q], q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv[:total], total

    def allreduce_sum_u64(self, a: np.ndarray, device) -> np.ndarray:
        t = torch.from_numpy(a.astype(np.int64)).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy().astype(np.uint64)


class ShardedBuilder:
    """The per-rank driver.  `reads` are this rank's own contiguous block of the global read sequence."""

    def __init__(self, K, max_read_len, init_slots, load_factor=0.7, device=0, track_order=True, group=None,
                 slack=None, exchange="pull", sub_blocks=4):
        from .graph import DBGBuilder
        self.ex = Exchange(group)
        self.n, self.rank = self.ex.world, self.ex.rank
        self.b = DBGBuilder(K=K, max_read_len=max_read_len, init_slots=init_slots, load_factor=load_factor, device=device,
                            track_order=track_order, shard_rank=self.rank, shard_count=self.n)
        self.device = torch.device("cuda", device)
        self.P_request, self.load_factor = init_slots, load_factor
        self.K, self.max_read_len = int(K), int(max_read_len)
        self.width = self.b.tuple_bytes // 8        # int64 words per tuple
        self._send = self._counts = None
        self.exchange_bytes = 0
        # "pull" (default): see _add_reads_pull;
        # "peer": OPTIMISTIC fused exchange -- one extraction pass stores tuples straight into fixed regions of
        #     the owners' receive buffers over NVLink peer mappings, in sub-blocks, the scatter of sub-block k+1 overlapping
        #     the owners' partition + insert of sub-block k; a region overflow (skew) redoes that sub-block exactly;
        # "peer_exact": count pass + all-gathered exact offsets + scatter pass (no overlap);
        # "peer_sliced": exact, (owner x slice) buckets;  "nccl": pack locally, then grouped send/recv
        self.exchange = exchange if self.n > 1 else "nccl"
        self.sub_blocks = max(1, int(sub_blocks))
        self.opt_fallbacks = 0
        self.sub_blocks_used = None
        self._sA = self._sB = None
        self._recv_ptr = None
        self._recv_cap = 0
        self._peer_ptrs = None
        # "pull": the source partitions by (owner, table slice) into its OWN send buffer (local stores only) and the owners read
        #     their regions over NVLink from inside the bucketed insert: no receive buffer, no owner-side partition pass
        self._send_ptr = None
        self._send_cap = 0
        self._src_ptrs = None

    def close(self):
        self._release_peers()
        self._release_send()
        self.b.close()

    # ---- pull exchange: send buffers readable by every peer ----
    def _release_send(self):
        if self._src_ptrs is not None:
            for q, p in enumerate(self._src_ptrs):
                if q != self.rank and p:
                    self.b.peer_close(p)
            self._src_ptrs = None
        if self._send_ptr:
            self.b.peer_free(self._send_ptr)
            self._send_ptr = None
            self._send_cap = 0

    def _ensure_send(self, cap_tuples):
        """(re)allocate the send buffers (every rank the same size) and map every peer's; collective"""
        if cap_tuples <= self._send_cap:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.ex.group)
        self._release_send()
        self._send_ptr, handle = self.b.peer_alloc(cap_tuples * self.width * 8)
        self._send_cap = cap_tuples
        handles = [None] * self.n
        dist.all_gather_object(handles, handle, group=self.ex.group)
        self._src_ptrs = [self._send_ptr if q == self.rank else self.b.peer_open(handles[q]) for q in range(self.n)]
        self._d_src = torch.tensor(self._src_ptrs, dtype=torch.int64, device=self.device)
        dist.barrier(group=self.ex.group)

    def _add_reads_pull(self, d_bases, d_offs, n_reads, first_base, total_bases, first_read_index, n_occ=None):
        from .graph import torch_stream_handle
        stream = torch_stream_handle(self.device)
        n, dev = self.n, self.device
        nbl = self.b.partition_info()[0]
        nbt = n * nbl
        if nbt > 4096:
            return self._add_reads_peer_opt(d_bases, d_offs, n_reads, first_base, total_bases, first_read_index, n_occ)
        if n_occ is None:
            lens = (d_offs[1:n_reads + 1] - d_offs[:n_reads]).clamp(max=int(self.max_read_len))
            n_occ = int((lens - (self.K - 1)).clamp(min=0).sum().item())
        # region size: the expected tuples of an (owner, slice) bucket + 25 % + 6 sigma, in whole insert tiles; the ranks agree
        # on the largest.  A region that is too small is detected and the block redone exactly.
        mean = n_occ / nbt
        capb = int(mean * 1.25 + 6.0 * mean ** 0.5 + 512)
        want = torch.tensor([capb], dtype=torch.int64, device=dev)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=self.ex.group)
        capb = (int(want.item()) + 511) // 512 * 512
        self._ensure_send(nbt * capb)
        fill = torch.zeros(nbt + 1, dtype=torch.int32, device=dev)
        allfill = torch.empty(n * (nbt + 1), dtype=torch.int32, device=dev)
        self.b.exchange_scatter_pull_device(d_bases.data_ptr(), d_offs.data_ptr(), n_reads, first_base, total_bases, first_read_index,
                                            n, self._send_ptr, capb, fill.data_ptr(), stream=stream)
        # the all-gather of the fill counters is also the barrier: every rank's regions are complete when it returns
        dist.all_gather_into_tensor(allfill, fill, group=self.ex.group)
        counts = allfill.view(n, nbt + 1).cpu()
        self.sub_blocks_used = 1
        if bool(counts[:, nbt].any()):
            self.opt_fallbacks += 1
            self.b.exchange_scatter_undo(stream=stream)
            torch.cuda.synchronize(dev)
            return self._add_reads_peer(d_bases, d_offs, n_reads, first_base, total_bases, first_read_index)
        mine = counts[:, self.rank * nbl:(self.rank + 1) * nbl]
        recv_total = int(mine.sum())
        self.exchange_bytes += int(counts[self.rank, :nbt].sum() - counts[self.rank, self.rank * nbl:(self.rank + 1) * nbl].sum()) * self.width * 8
        self.b.insert_pull_device(self._d_src.data_ptr(), n, capb, allfill.data_ptr(), nbt + 1, recv_total, stream=stream)
        self._keep = (fill, allfill)
        return recv_total

    # ---- peer receive buffers ----
    def _release_peers(self):
        if self._peer_ptrs is not None:
            for q, p in enumerate(self._peer_ptrs):
                if q != self.rank and p:
                    self.b.peer_close(p)
            self._peer_ptrs = None
        if self._recv_ptr:
            self.b.peer_free(self._recv_ptr)
            self._recv_ptr = None
            self._recv_cap = 0

    def _ensure_peers(self, cap_tuples):
        """(re)allocate the receive buffers so that every rank can take cap_tuples; collective"""
        want = torch.tensor([int(cap_tuples)], dtype=torch.int64, device=self.device)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=self.ex.group)
        cap = int(want.item())
        if cap <= self._recv_cap:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.ex.group)
        self._release_peers()
        cap = int(cap * 1.25) + 4096
        self._recv_ptr, handle = self.b.peer_alloc(cap * self.width * 8)
        self._recv_cap = cap
        handles = [None] * self.n
        dist.all_gather_object(handles, handle, group=self.ex.group)
        self._peer_ptrs = [self._recv_ptr if q == self.rank else self.b.peer_open(handles[q]) for q in range(self.n)]
        self._d_ptrs = torch.tensor(self._peer_ptrs, dtype=torch.int64, device=self.device)
        dist.barrier(group=self.ex.group)

    def _add_reads_peer_opt(self, d_bases, d_offs, n_reads, first_base, total_bases, first_read_index, n_occ=None):
        """optimistic fused exchange, pipelined over sub-blocks (see __init__).  Streams: A = scatter + the small
        all-gather of the fill counters (which doubles as the cross-rank barrier), B = owner-side partition + insert."""
        n, dev, rank, S = self.n, self.device, self.rank, self.sub_blocks
        tb = self.width * 8
        if self._sA is None:
            self._sA, self._sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        sA, sB = self._sA, self._sB
        cur = torch.cuda.current_stream(dev)
        # a sub-block must still carry enough tuples per table slice for the owners' partitioned insert to pay (the library's
        # path selection: tuples x 80 B (150 B with 128-bit keys) against the bytes of the shard's table); fewer, larger
        # sub-blocks when the table is large relative to the block (K = 63: 64-B nodes)
        from . import capi
        P = capi.find_next_prime(max(3, int(self.P_request)))
        table_bytes = (P + n - 1) // n * (64 if self.b.wide else 32)
        est_tuples = int(n_occ) if n_occ else int(total_bases) * 0.8
        S = max(1, min(S, n_reads, int(est_tuples * (150 if self.b.wide else 80) / (1.25 * table_bytes))))
        self.sub_blocks_used = S
        cuts = [n_reads * k // S for k in range(S + 1)]
        offs_h = d_offs[torch.tensor(cuts, device=dev)].cpu().tolist()
        # region size: the sub-block's OCCURRENCES (sum of min(len, -r) - K + 1 over its reads, counted on the device), not its
        # bases (100-bp reads at K = 63 carry 38 occurrences per 100 bases: sizing by bases took 96 GB of receive buffers on
        # C3); a region that is too small is detected and the sub-block redone exactly, so the estimate only has to be good
        lens = (d_offs[1:n_reads + 1] - d_offs[:n_reads]).clamp(max=int(self.max_read_len))
        occ_cum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), (lens - (self.K - 1)).clamp(min=0).cumsum(0)])
        occ_h = occ_cum[torch.tensor(cuts, device=dev)].cpu().tolist()
        sub_occ = max(occ_h[k + 1] - occ_h[k] for k in range(S))
        want = torch.tensor([int(sub_occ / n * 1.25) + 4096], dtype=torch.int64, device=dev)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=self.ex.group)
        cap_pair = int(want.item())
        n_sets = 2 if S > 1 else 1                                  # double buffering only when there is something to overlap
        self._ensure_peers(n_sets * n * cap_pair)
        d_ptrs = self._d_ptrs
        fills = [torch.zeros(n + 1, dtype=torch.int32, device=dev) for _ in range(S)]
        allfill = [torch.empty(n * (n + 1), dtype=torch.int32, device=dev) for _ in range(S)]
        sA.wait_stream(cur); sB.wait_stream(cur)

        def scatter(k):
            r0, r1 = cuts[k], cuts[k + 1]
            with torch.cuda.stream(sA):
                self.b.exchange_scatter_opt_device(d_bases.data_ptr(), d_offs.data_ptr() + 8 * r0, r1 - r0, offs_h[k], offs_h[k + 1] - offs_h[k],
                                                   first_read_index + r0, n, d_ptrs.data_ptr(), ((k & 1) * n + rank) * cap_pair, cap_pair,
                                                   fills[k].data_ptr(), stream=sA.cuda_stream)

        def gather(k):
            with torch.cuda.stream(sA):
                dist.all_gather_into_tensor(allfill[k], fills[k], group=self.ex.group)
                return allfill[k].view(n, n + 1).cpu()                 # host waits for A: scatter(k) everywhere, insert(k-1) here

        recv_total = 0
        scatter(0)
        counts = gather(0)
        for k in range(S):
            overflow = bool(counts[:, n].any())
            if k + 1 < S and not overflow:
                scatter(k + 1)                                          # runs while the owners work on sub-block k
            if overflow:
                # some region was too small (skewed input): nothing of sub-block k was inserted; everybody redoes it exactly
                self.opt_fallbacks += 1
                self.b.exchange_scatter_undo(stream=sA.cuda_stream)
                torch.cuda.synchronize(dev)
                r0, r1 = cuts[k], cuts[k + 1]
                with torch.cuda.stream(sB):
                    recv_total += self._add_reads_peer(d_bases, d_offs[r0:], r1 - r0, offs_h[k], offs_h[k + 1] - offs_h[k], first_read_index + r0)
                torch.cuda.synchronize(dev)
                self._ensure_peers(n_sets * n * cap_pair)
                d_ptrs = self._d_ptrs
                if k + 1 < S:
                    scatter(k + 1)
            else:
                mine = counts[:, rank].numpy().astype(np.uint64)
                recv_total += int(mine.sum())
                self.exchange_bytes += int(counts[rank, :n].sum() - counts[rank, rank]) * tb
                # (the host has already waited for the all-gather of sub-block k: every rank's stores into this rank's
                #  regions have landed; B must NOT wait for A here, scatter(k+1) is running there)
                with torch.cuda.stream(sB):
                    self.b.insert_tuple_regions_device(self._recv_ptr + (k & 1) * n * cap_pair * tb, cap_pair, mine, stream=sB.cuda_stream)
            if k + 1 < S:
                sA.wait_stream(sB)                                      # set (k+1)&1 ... is free again only after everybody's insert(k): gate the next barrier
                counts = gather(k + 1)
        cur.wait_stream(sA); cur.wait_stream(sB)
        return recv_total

    def _add_reads_peer(self, d_bases, d_offs, n_reads, first_base, total_bases, first_read_index):
        """fused exchange.  "peer": exchange buckets = owners; stores of one CTA go to n streams, so NVLink sees long
        contiguous runs, and the owner re-partitions what it received by table slice.  "peer_sliced": buckets =
        (owner, slice): the receive buffer arrives in slice order and goes straight to the bucketed insert, but the
        16-B stores scatter over n * n_slices streams (measured slower over NVLink: short packets)."""
        from .graph import torch_stream_handle
        stream = torch_stream_handle(self.device)
        n, dev = self.n, self.device
        sliced = self.exchange == "peer_sliced"
        nbl = self.b.partition_info()[0] if sliced else 1      # table slices per shard (same on every rank)
        nbt = n * nbl                                          # exchange buckets, owner-major
        if self._counts is None or self._counts.numel() != nbt:
            self._counts = torch.zeros(nbt, dtype=torch.int64, device=dev)
        # pass 1: how many tuples this block has for every bucket
        self.b.exchange_count_device(d_bases.data_ptr(), d_offs.data_ptr(), n_reads, first_base, total_bases, n,
                                     self._counts.data_ptr(), by_slice=sliced, stream=stream)
        allc = torch.empty(n * nbt, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, self._counts, group=self.ex.group)
        allc = allc.view(n, nbt)                                   # [source rank][bucket]
        col = allc.sum(dim=0).view(n, nbl)                         # tuples per (owner, slice) over all sources
        start = (torch.cumsum(col, dim=1) - col)                   # start of every slice inside its owner's buffer
        d_base = (start.reshape(-1) + allc[: self.rank].sum(dim=0)).contiguous()   # + tuples of lower ranks for that bucket
        recv_total = int(col[self.rank].sum().item())
        self._ensure_peers(recv_total)
        d_ptrs = self._d_ptrs.repeat_interleave(nbl).contiguous() if sliced else self._d_ptrs
        # everybody is done with the previous content of the receive buffers before anybody overwrites them
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.ex.group)
        # pass 2: regenerate the occurrences and store every tuple into its owner's buffer over NVLink
        self.b.exchange_scatter_device(d_bases.data_ptr(), d_offs.data_ptr(), n_reads, first_base, total_bases, first_read_index,
                                       n, d_ptrs.data_ptr(), d_base.data_ptr(), by_slice=sliced, stream=stream)
        torch.cuda.synchronize(dev)                                # my stores have landed ...
        dist.barrier(group=self.ex.group)                          # ... and so have everybody else's
        mine = self._counts.view(n, nbl).sum(dim=1).cpu().tolist()
        self.exchange_bytes += (sum(mine) - mine[self.rank]) * self.width * 8
        if sliced:
            slice_offs = torch.cat([start[self.rank], col[self.rank].sum().reshape(1)]).contiguous()
            self.b.insert_sliced_device(self._recv_ptr, recv_total, slice_offs.data_ptr(), stream=stream)
            self._keep = (slice_offs, d_ptrs, d_base)
        else:
            self.b.insert_tuples_device(self._recv_ptr, recv_total, stream=stream)
            self._keep = (d_ptrs, d_base)
        return recv_total

    def _buffers(self, capacity):
        need = capacity * self.width
        if self._send is None or self._send.numel() < need:
            self._send = torch.empty(need, dtype=torch.int64, device=self.device)
            self._counts = torch.zeros(self.n, dtype=torch.int64, device=self.device)

    def add_reads_device(self, d_bases: torch.Tensor, d_offs: torch.Tensor, n_reads, first_base, total_bases,
                         first_read_index, n_occ_upper=None):
        """one block of this rank's reads, device resident (an occurrence starts at a distinct base, so
        total_bases bounds the tuple count)"""
        if self.exchange == "pull":
            return self._add_reads_pull(d_bases, d_offs, n_reads, first_base, total_bases, first_read_index, n_occ_upper)
        if self.exchange == "peer":
            return self._add_reads_peer_opt(d_bases, d_offs, n_reads, first_base, total_bases, first_read_index, n_occ_upper)
        if self.exchange in ("peer_exact", "peer_sliced"):
            return self._add_reads_peer(d_bases, d_offs, n_reads, first_base, total_bases, first_read_index)
        from .graph import torch_stream_handle
        stream = torch_stream_handle(self.device)
        cap = int(total_bases)
        self._buffers(cap)
        # tuples come back packed by owner rank: sizes in _counts, offsets = their prefix sums
        self.b.extract_tuples_device(d_bases.data_ptr(), d_offs.data_ptr(), n_reads, first_base, total_bases, first_read_index,
                                     self.n, self._send.data_ptr(), cap, self._counts.data_ptr(), stream=stream)
        recv_counts = self.ex.exchange_counts(self._counts)
        sc = self._counts.cpu().tolist()
        rc = recv_counts.cpu().tolist()
        view = self._send[: sum(sc) * self.width].view(-1, self.width)
        buckets, off = [], 0
        for q in range(self.n):
            buckets.append(view[off:off + sc[q]])
            off += sc[q]
        recv, total = self.ex.exchange_payload(buckets, rc, self.width, self._send)
        self.exchange_bytes += (sum(sc) - sc[self.rank]) * self.width * 8
        self.b.insert_tuples_device(recv.data_ptr(), total, stream=stream)
        self._keep = recv    # keep alive until the insert kernel ran
        return total

    def merged_table_from_dumps(self):
        """fallback after LayoutNeedsMerge (collective): every rank's nodes (+ the overflow nodes it had already handed
        over, which live only in its tail blob) are gathered on rank 0, which replays them in first-occurrence order on
        the host -- the reference's layout by construction.  Returns (array, nul_flag) on rank 0, (None, None) elsewhere."""
        d = self.b.dump_shard()
        blob = getattr(self, "_tail_blob", None)
        extra = blob_margin_nodes(blob) if blob else None
        polyA = self.b.get_polyA_counts()
        gathered = [None] * self.n
        dist.all_gather_object(gathered, (d, extra), group=self.ex.group)
        if self.rank != 0:
            return None, None
        dumps = [g[0] for g in gathered] + [g[1] for g in gathered if g[1] is not None]
        cl = lambda q: sum(int(min(int(q[i]), 255)) << (24 - 8 * i) for i in range(4))
        arr, nul, _ = merge_dumps_host(dumps, self.P_request, self.load_factor, self.b.wide, cl(polyA[:4]), cl(polyA[4:]))
        return arr, nul

    def _ring_blobs(self, blob: bytes) -> bytes:
        """every rank's tail blob goes to the rank on its right (sizes first, then one padded all-gather: blobs are
        a few hundred bytes)"""
        n, dev = self.n, self.device
        size = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
        sizes = torch.empty(n, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, size, group=self.ex.group)
        sizes = sizes.cpu().tolist()
        mx = (max(sizes) + 7) // 8 * 8
        mine = torch.zeros(mx, dtype=torch.uint8)
        mine[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        allb = torch.empty(n * mx, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, mine.to(dev), group=self.ex.group)
        left = (self.rank - 1) % n
        return bytes(allb[left * mx: left * mx + sizes[left]].cpu().numpy().tobytes())

    def export_into(self, img: "SharedImage", st):
        """after finalize(layout=True): this rank's slice goes into the shared table image; once every slice is in, rank 0
        fixes the nul_flag bytes shared between neighbours and appends the k-mer-0 node (DBGgraph.cpp:418).  Collective."""
        from . import capi
        edges = self.b.export_shard_slice(img.arr_ptr, img.nul_ptr)
        torch.cuda.synchronize(self.device)
        e = torch.full((4,), -1, dtype=torch.int64, device=self.device)
        for i, x in enumerate(edges):
            e[i] = x
        alle = torch.empty(4 * self.n, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(alle, e, group=self.ex.group)          # also the barrier: every slice is in the image
        if self.rank == 0:
            ed = [int(x) for x in alle.cpu().tolist() if x >= 0]
            capi.host_fix_nul_bytes(img.arr, img.nul, img.P, img.wide, ed)
            capi.host_polyA_insert(img.arr, img.nul, img.P, img.wide, st["polyA_l"], st["polyA_r"])
        dist.barrier(group=self.ex.group)

    def finalize(self, layout=False):
        """layout=True: cross-shard hand-off of the boundary clusters, then every rank lays out its slice of the
        reference's table (dbg_export_shard_slice / export_slice copies it out)"""
        torch.cuda.synchronize(self.device)
        polyA = self.ex.allreduce_sum_u64(self.b.get_polyA_counts(), self.device)
        self.b.set_polyA_counts(polyA)
        if layout and self.n > 1:
            from . import capi

            def together(fn):
                """run a step that may refuse with DBG_ERR_STATE; all ranks learn whether anybody refused"""
                out, ok = None, 1
                try:
                    out = fn()
                except capi.DbgError as e:
                    if e.code != capi.DBG_ERR_STATE:
                        raise
                    ok = 0
                t = torch.tensor([ok], dtype=torch.int64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.ex.group)
                return out, bool(t.item())
            self._tail_blob, ok = together(self.b.shard_tail_export)
            if not ok:
                raise LayoutNeedsMerge("dbg_shard_tail_export refused on some rank")
            incoming = self._ring_blobs(self._tail_blob)
            _, ok = together(lambda: self.b.shard_tail_import(incoming))
            if not ok:
                raise LayoutNeedsMerge("dbg_shard_tail_import refused on some rank")
            st, ok = together(self.b.finalize)
            if not ok:
                raise LayoutNeedsMerge("the windowed layout refused on some rank")
        else:
            st = self.b.finalize()
        tot = self.ex.allreduce_sum_u64(np.array([st["count"], st["occurrences"]], dtype=np.uint64), self.device)
        st["global_count"] = int(tot[0]) + 1          # + the k-mer-0 node
        st["global_occurrences"] = int(tot[1])
        return st
