"""Host-side mirror of the reference's build interface (DBG_contig/DBGgraph.h:25-66, kmerSet.h:88-99) on
top of the C ABI.  Names, argument meaning and error behaviour follow the reference:

    KmerSet            <- struct KmerSet (kmerSet.h:88-99): size, count, max, load_factor, array, nul_flag, del_flag
    build_debruijn_graph(reads_files, KmerSize=31, maxReadLen=250, Input_file_format=1, initHashSize=1.0, ...)
                       <- DBGgraph.cpp:364 + the globals main.cpp:166-193 sets from the command line
    calculate_kmer_links(kset, KmerFreqCutoff=2)   <- contig.cpp:107-205

All computation happens in libdbgb200.so on the GPU; this file only moves buffers.
"""
from __future__ import annotations

import ctypes as C
import gzip
import sys
from dataclasses import dataclass, field

import numpy as np

from . import capi

NODE16 = np.dtype([("kmer", "<u8"), ("l_link", "<u4"), ("r_link", "<u4")])
NODE32 = np.dtype([("kmer", "<u8"), ("kmer_hi", "<u8"), ("l_link", "<u4"), ("r_link", "<u4"), ("pad", "<u8")])
UINT64_MAX = (1 << 64) - 1


CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the C ABI reads a NULL stream as "the context's own stream"


def torch_stream_handle(device=None) -> int:
    """cudaStream_t of torch's current stream as an int the C ABI accepts (the default stream's handle is 0,
    which the ABI would read as NULL, so it is mapped to cudaStreamLegacy)"""
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream) or CUDA_STREAM_LEGACY


def init_slots_from_g(init_hash_size_g: float) -> int:
    """(uint64)(initHashSize * 1000000000), DBGgraph.cpp:381"""
    return int(float(init_hash_size_g) * 1000000000)


class DBGBuilder:
    """One dbg_ctx: create -> submit* -> finalize -> export*.  Context manager; frees the device table."""

    def __init__(self, K=31, max_read_len=250, init_slots=None, init_g=None, load_factor=0.7, device=0,
                 track_order=True, shard_rank=0, shard_count=1, force_wide=False):
        self.L = capi.load()
        if init_slots is None:
            init_slots = init_slots_from_g(1.0 if init_g is None else init_g)
        p = capi.dbg_params()
        p.K, p.max_read_len, p.init_slots = int(K), int(max_read_len), int(init_slots)
        p.load_factor, p.device, p.track_order = float(load_factor), int(device), int(bool(track_order))
        p.shard_rank, p.shard_count, p.force_wide = int(shard_rank), int(shard_count), int(bool(force_wide))
        self.K, self.device = int(K), int(device)
        self.h = C.c_void_p()
        rc = self.L.dbg_create(C.byref(self.h), C.byref(p))
        if rc != capi.DBG_OK:
            msg = self.L.dbg_last_error().decode()
            if self.h:
                self.L.dbg_destroy(self.h)
                self.h = C.c_void_p()
            raise capi.DbgError(rc, "dbg_create", msg)
        self.wide = int(K) > 31 or bool(force_wide)
        self.stats = None

    # ---- life cycle ----
    def close(self):
        if getattr(self, "h", None):
            self.L.dbg_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(self.L.dbg_reset(self.h), "dbg_reset")
        self.stats = None

    def set_stream(self, stream):
        """run this context on a caller-owned cudaStream_t (int handle), e.g. torch.cuda.current_stream().cuda_stream"""
        capi.check(self.L.dbg_set_stream(self.h, stream), "dbg_set_stream")

    # ---- input ----
    def submit(self, bases, offs):
        """one reader block from host memory: reads are bases[offs[i]:offs[i+1]] (ASCII uint8)"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        if n <= 0:
            return
        bp = bases.ctypes.data if bases.size else 0
        capi.check(self.L.dbg_submit_reads(self.h, bp or None, offs.ctypes.data, n), "dbg_submit_reads")

    def submit_ptr(self, bases_ptr, offs_ptr, n_reads):
        capi.check(self.L.dbg_submit_reads(self.h, bases_ptr, offs_ptr, int(n_reads)), "dbg_submit_reads")

    def submit_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases, first_read_index=None, stream=None):
        idx = UINT64_MAX if first_read_index is None else int(first_read_index)
        capi.check(self.L.dbg_submit_reads_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base),
                                                  int(total_bases), idx, stream), "dbg_submit_reads_device")

    def extract_tuples_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases, first_read_index,
                              n_parts, d_tuples_ptr, capacity, d_counts_ptr, stream=None):
        capi.check(self.L.dbg_extract_tuples_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base),
                                                    int(total_bases), int(first_read_index), int(n_parts), d_tuples_ptr,
                                                    int(capacity), d_counts_ptr, stream), "dbg_extract_tuples_device")

    # ---- pull exchange ----
    def exchange_scatter_pull_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases, first_read_index, n_parts,
                                     d_send_ptr, capb, d_fill_ptr, stream=None):
        capi.check(self.L.dbg_exchange_scatter_pull_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base), int(total_bases),
                                                           int(first_read_index), int(n_parts), d_send_ptr, int(capb), d_fill_ptr, stream),
                   "dbg_exchange_scatter_pull_device")

    def insert_pull_device(self, d_src_ptrs_ptr, n_src, capb, d_fills_ptr, fill_stride, n_tuples_upper, stream=None):
        capi.check(self.L.dbg_insert_pull_device(self.h, d_src_ptrs_ptr, int(n_src), int(capb), d_fills_ptr, int(fill_stride),
                                                 int(n_tuples_upper), stream), "dbg_insert_pull_device")

    # ---- fused exchange over peer memory ----
    def peer_alloc(self, nbytes):
        """-> (device pointer, 64-byte IPC handle) of a cudaMalloc'ed receive buffer on this context's device"""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        capi.check(self.L.dbg_peer_alloc(self.h, int(nbytes), C.byref(ptr), handle), "dbg_peer_alloc")
        return ptr.value, bytes(handle)

    def peer_open(self, handle: bytes):
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        capi.check(self.L.dbg_peer_open(self.h, buf, C.byref(ptr)), "dbg_peer_open")
        return ptr.value

    def peer_close(self, ptr):
        capi.check(self.L.dbg_peer_close(self.h, ptr), "dbg_peer_close")

    def peer_free(self, ptr):
        capi.check(self.L.dbg_peer_free(self.h, ptr), "dbg_peer_free")

    def exchange_count_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases, n_parts, d_counts_ptr, by_slice=False,
                              stream=None):
        capi.check(self.L.dbg_exchange_count_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base), int(total_bases),
                                                    int(n_parts), int(bool(by_slice)), d_counts_ptr, stream), "dbg_exchange_count_device")

    def exchange_scatter_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases, first_read_index, n_parts,
                                d_dst_ptrs_ptr, d_dst_base_ptr, by_slice=False, stream=None):
        capi.check(self.L.dbg_exchange_scatter_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base), int(total_bases),
                                                      int(first_read_index), int(n_parts), int(bool(by_slice)), d_dst_ptrs_ptr,
                                                      d_dst_base_ptr, stream), "dbg_exchange_scatter_device")

    def exchange_scatter_opt_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases, first_read_index, n_parts,
                                    d_dst_ptrs_ptr, region_off, cap_pair, d_fill_ptr, stream=None):
        capi.check(self.L.dbg_exchange_scatter_opt_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base), int(total_bases),
                                                          int(first_read_index), int(n_parts), d_dst_ptrs_ptr, int(region_off),
                                                          int(cap_pair), d_fill_ptr, stream), "dbg_exchange_scatter_opt_device")

    def exchange_scatter_undo(self, stream=None):
        capi.check(self.L.dbg_exchange_scatter_undo(self.h, stream), "dbg_exchange_scatter_undo")

    def insert_tuple_regions_device(self, d_base_ptr, stride_tuples, counts, stream=None):
        a = np.ascontiguousarray(counts, dtype=np.uint64)
        capi.check(self.L.dbg_insert_tuple_regions_device(self.h, d_base_ptr, len(a), int(stride_tuples), a.ctypes.data, stream),
                   "dbg_insert_tuple_regions_device")

    def insert_sliced_device(self, d_tuples_ptr, n, d_slice_offs_ptr, stream=None):
        capi.check(self.L.dbg_insert_sliced_device(self.h, d_tuples_ptr, int(n), d_slice_offs_ptr, stream), "dbg_insert_sliced_device")

    def partition_info(self):
        n, sh = C.c_uint32(0), C.c_int32(0)
        capi.check(self.L.dbg_partition_info(self.h, C.byref(n), C.byref(sh)), "dbg_partition_info")
        return n.value, sh.value

    def insert_tuples_device(self, d_tuples_ptr, n, stream=None):
        capi.check(self.L.dbg_insert_tuples_device(self.h, d_tuples_ptr, int(n), stream), "dbg_insert_tuples_device")

    @property
    def tuple_bytes(self):
        return int(self.L.dbg_tuple_bytes(self.h))

    def get_polyA_counts(self):
        a = np.zeros(8, dtype=np.uint64)
        capi.check(self.L.dbg_get_polyA_counts(self.h, a.ctypes.data), "dbg_get_polyA_counts")
        return a

    def set_polyA_counts(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        capi.check(self.L.dbg_set_polyA_counts(self.h, a.ctypes.data), "dbg_set_polyA_counts")

    # ---- cross-shard reference layout (include/dbg_b200.h: dbg_shard_tail_export ...) ----
    def shard_tail_export(self) -> bytes:
        """this rank's tail unit (boundary cluster + overflow nodes) for the rank on its right"""
        n = C.c_uint64(0)
        capi.check(self.L.dbg_shard_tail_export(self.h, None, 0, C.byref(n)), "dbg_shard_tail_export")
        buf = (C.c_uint8 * n.value)()
        capi.check(self.L.dbg_shard_tail_export(self.h, buf, n.value, C.byref(n)), "dbg_shard_tail_export")
        return bytes(buf)

    def shard_tail_import(self, blob: bytes):
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        capi.check(self.L.dbg_shard_tail_import(self.h, buf, len(blob)), "dbg_shard_tail_import")

    def shard_slice_info(self):
        """-> (first global slot, number of slots, device pointer) of the laid-out slice"""
        g, n, d = C.c_uint64(0), C.c_uint64(0), C.c_void_p()
        capi.check(self.L.dbg_shard_slice_info(self.h, C.byref(g), C.byref(n), C.byref(d)), "dbg_shard_slice_info")
        return g.value, n.value, d.value

    def export_shard_slice(self, array_ptr, nul_ptr):
        """copy the slice into the FULL table image at array_ptr / nul_ptr (host); -> edge slots whose nul_flag byte is
        shared with a neighbour (fix them with capi host_fix_nul_bytes once every slice is in)"""
        edges = (C.c_uint64 * 4)()
        capi.check(self.L.dbg_export_shard_slice(self.h, array_ptr, nul_ptr, edges), "dbg_export_shard_slice")
        return [int(e) for e in edges if e != UINT64_MAX]

    # ---- results ----
    def finalize(self):
        st = capi.dbg_stats()
        capi.check(self.L.dbg_finalize(self.h, C.byref(st)), "dbg_finalize")
        self.stats = st.as_dict()
        return self.stats

    def get_stats(self):
        st = capi.dbg_stats()
        capi.check(self.L.dbg_get_stats(self.h, C.byref(st)), "dbg_get_stats")
        return st.as_dict()

    def export_kmerset(self, array=None, nul_flag=None):
        """-> (array[P] of NODE16/NODE32 in reference slot layout, nul_flag[P/8+1])"""
        P = self.stats["array_size"] if self.stats else self.get_stats()["array_size"]
        dt = NODE32 if self.wide else NODE16
        if array is None:
            array = np.empty(P, dtype=dt)
        if nul_flag is None:
            nul_flag = np.empty(P // 8 + 1, dtype=np.uint8)
        capi.check(self.L.dbg_export_kmerset(self.h, array.ctypes.data, nul_flag.ctypes.data), "dbg_export_kmerset")
        return array, nul_flag

    def finish_export(self, bases, offs, array=None, nul_flag=None):
        """last block + finalize + export in one pipelined call (dbg_finish_export) -> (stats, array, nul_flag)"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = max(len(offs) - 1, 0)
        P = self.get_stats()["array_size"]
        if array is None:
            array = np.empty(P, dtype=NODE32 if self.wide else NODE16)
        if nul_flag is None:
            nul_flag = np.empty(P // 8 + 1, dtype=np.uint8)
        st = capi.dbg_stats()
        capi.check(self.L.dbg_finish_export(self.h, bases.ctypes.data if bases.size else None, offs.ctypes.data if n else None, n,
                                            C.byref(st), array.ctypes.data, nul_flag.ctypes.data), "dbg_finish_export")
        self.stats = st.as_dict()
        return self.stats, array, nul_flag

    def finish_export_ptr(self, bases_ptr, offs_ptr, n_reads, array_ptr, nul_ptr):
        st = capi.dbg_stats()
        capi.check(self.L.dbg_finish_export(self.h, bases_ptr, offs_ptr, int(n_reads), C.byref(st), array_ptr, nul_ptr), "dbg_finish_export")
        self.stats = st.as_dict()
        return self.stats

    def export_info(self):
        """how the last export_kmerset moved the table (include/dbg_b200.h: dbg_export_info)"""
        a = np.zeros(4, dtype=np.uint64)
        capi.check(self.L.dbg_export_info(self.h, a.ctypes.data), "dbg_export_info")
        return dict(chunks_compact=int(a[0]), chunks_plain=int(a[1]), link_bytes=int(a[2]), nodes=int(a[3]))

    def export_links(self, freq_cutoff=2, lists=True):
        P = self.stats["array_size"]
        klink = np.empty(2 * P, dtype=np.uint8)
        del_flag = np.empty(P // 8 + 1, dtype=np.uint8)
        depth = np.zeros(256, dtype=np.int64)
        stats3 = np.zeros(3, dtype=np.int64)
        nt, nb = C.c_uint64(0), C.c_uint64(0)
        # size query first (lists NULL), then exact-size lists
        capi.check(self.L.dbg_export_links(self.h, int(freq_cutoff), klink.ctypes.data, del_flag.ctypes.data, depth.ctypes.data,
                                           None, C.byref(nt), None, C.byref(nb), stats3.ctypes.data), "dbg_export_links")
        tips = np.zeros(max(nt.value, 1), dtype=np.uint64)
        branches = np.zeros(max(nb.value, 1), dtype=np.uint64)
        if lists:
            ct, cb = C.c_uint64(len(tips)), C.c_uint64(len(branches))
            capi.check(self.L.dbg_export_links(self.h, int(freq_cutoff), None, None, None, tips.ctypes.data, C.byref(ct),
                                               branches.ctypes.data, C.byref(cb), None), "dbg_export_links")
        return dict(klink=klink, del_flag=del_flag, depth_stat=depth, tips=tips[:nt.value], branches=branches[:nb.value],
                    total=int(stats3[0]), deleted=int(stats3[1]), linear=int(stats3[2]))

    def dump_compact(self, freq_cutoff=-1):
        n = C.c_uint64(0)
        capi.check(self.L.dbg_dump_compact(self.h, int(freq_cutoff), None, None, None, None, None, C.byref(n)), "dbg_dump_compact")
        m = max(n.value, 1)
        out = dict(slot=np.zeros(m, np.uint64), kmer=np.zeros(m, np.uint64), kmer_hi=np.zeros(m, np.uint64),
                   l=np.zeros(m, np.uint32), r=np.zeros(m, np.uint32))
        cap = C.c_uint64(m)
        capi.check(self.L.dbg_dump_compact(self.h, int(freq_cutoff), out["slot"].ctypes.data, out["kmer"].ctypes.data,
                                           out["kmer_hi"].ctypes.data, out["l"].ctypes.data, out["r"].ctypes.data,
                                           C.byref(cap)), "dbg_dump_compact")
        return {k: v[:n.value] for k, v in out.items()}

    def dump_shard(self):
        """unordered nodes of this context's build table: dict kmer, kmer_hi, l, r, ord (no k-mer-0 node)"""
        n = C.c_uint64(0)
        capi.check(self.L.dbg_dump_shard(self.h, None, None, None, None, None, C.byref(n)), "dbg_dump_shard")
        m = max(n.value, 1)
        out = dict(kmer=np.zeros(m, np.uint64), kmer_hi=np.zeros(m, np.uint64), l=np.zeros(m, np.uint32),
                   r=np.zeros(m, np.uint32), ord=np.zeros(m, np.uint64))
        cap = C.c_uint64(m)
        capi.check(self.L.dbg_dump_shard(self.h, out["kmer"].ctypes.data, out["kmer_hi"].ctypes.data, out["l"].ctypes.data,
                                         out["r"].ctypes.data, out["ord"].ctypes.data, C.byref(cap)), "dbg_dump_shard")
        return {k: v[:n.value] for k, v in out.items()}

    def device_image(self):
        a, f = C.c_void_p(), C.c_void_p()
        capi.check(self.L.dbg_device_image(self.h, C.byref(a), C.byref(f)), "dbg_device_image")
        return a.value, f.value

    def timings(self):
        ms = np.zeros(8, dtype=np.float32)
        capi.check(self.L.dbg_get_timings(self.h, ms.ctypes.data), "dbg_get_timings")
        return dict(clear_ms=float(ms[0]), build_ms=float(ms[1]), layout_ms=float(ms[2]), links_ms=float(ms[3]),
                    h2d_ms=float(ms[4]), d2h_ms=float(ms[5]), insert_ms=float(ms[6]), scatter_ms=float(ms[7]))

    @property
    def launches(self):
        return int(self.L.dbg_launch_count(self.h))

    def path_counts(self):
        """read blocks per build path: direct, exact partition, optimistic partition, optimistic overflows (redone exactly)"""
        c = np.zeros(4, dtype=np.uint64)
        capi.check(self.L.dbg_path_counts(self.h, c.ctypes.data), "dbg_path_counts")
        return dict(direct=int(c[0]), exact=int(c[1]), optimistic=int(c[2]), overflows=int(c[3]))


class MultiGpuBuilder:
    """dbg_mg_*: ONE process driving several GPUs behind the calls of a single-GPU build (include/dbg_b200.h)."""

    def __init__(self, n_gpus, K=31, max_read_len=250, init_slots=None, init_g=None, load_factor=0.7, devices=None, track_order=True,
                 force_wide=False):
        self.L = capi.load()
        if init_slots is None:
            init_slots = init_slots_from_g(1.0 if init_g is None else init_g)
        p = capi.dbg_params()
        p.K, p.max_read_len, p.init_slots = int(K), int(max_read_len), int(init_slots)
        p.load_factor, p.track_order, p.force_wide = float(load_factor), int(bool(track_order)), int(bool(force_wide))
        self.h = C.c_void_p()
        dev = (C.c_int32 * n_gpus)(*devices) if devices is not None else None
        rc = self.L.dbg_mg_create(C.byref(self.h), C.byref(p), int(n_gpus), dev)
        if rc != capi.DBG_OK:
            msg = self.L.dbg_mg_last_error().decode()
            if self.h:
                self.L.dbg_mg_destroy(self.h)
                self.h = C.c_void_p()
            raise capi.DbgError(rc, "dbg_mg_create", msg)
        self.wide = int(K) > 31 or bool(force_wide)
        self.stats = None

    def _check(self, rc, where):
        if rc != capi.DBG_OK:
            raise capi.DbgError(rc, where, self.L.dbg_mg_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.dbg_mg_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def submit(self, bases, offs):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        if n > 0:
            self._check(self.L.dbg_mg_submit_reads(self.h, bases.ctypes.data if bases.size else None, offs.ctypes.data, n), "dbg_mg_submit_reads")

    def submit_ptr(self, bases_ptr, offs_ptr, n_reads):
        self._check(self.L.dbg_mg_submit_reads(self.h, bases_ptr, offs_ptr, int(n_reads)), "dbg_mg_submit_reads")

    def finalize(self):
        st = capi.dbg_stats()
        self._check(self.L.dbg_mg_finalize(self.h, C.byref(st)), "dbg_mg_finalize")
        self.stats = st.as_dict()
        return self.stats

    def export_kmerset(self, array=None, nul_flag=None):
        P = self.stats["array_size"]
        if array is None:
            array = np.zeros(P, dtype=NODE32 if self.wide else NODE16)
        if nul_flag is None:
            nul_flag = np.zeros(P // 8 + 1, dtype=np.uint8)
        self._check(self.L.dbg_mg_export_kmerset(self.h, array.ctypes.data, nul_flag.ctypes.data), "dbg_mg_export_kmerset")
        return array, nul_flag

    def info(self):
        a = np.zeros(4, dtype=np.uint64)
        self._check(self.L.dbg_mg_info(self.h, a.ctypes.data), "dbg_mg_info")
        return dict(rounds=int(a[0]), regrows=int(a[1]), fallback=bool(a[2]), cap_pair=int(a[3]))


def replay_growth(nodes, reads_per_file, init_slots, load_factor=0.7, max_double_times=10, buffer_reads=10000,
                  polyA_l=0, polyA_r=0, layout=True):
    """Host-side replay of the reference's table growth (-e / enlarge) from a finished graph: `nodes` is the dict that
    DBGBuilder.dump_shard() returns (kmer, kmer_hi, l, r, ord; no k-mer-0 node).  Returns (plan dict, array, nul_flag);
    array/nul_flag are None when layout=False, or when the reference would have dropped reads (plan["truncated"]).
    The table comes back in the reference's slot layout AFTER its doublings (dbg_replay_growth, include/dbg_b200.h)."""
    L = capi.load()
    n = len(nodes["kmer"])
    wide = bool(np.any(nodes.get("kmer_hi", np.zeros(0, np.uint64)) != 0)) or bool(nodes.get("wide", False))
    klo = np.ascontiguousarray(nodes["kmer"], dtype=np.uint64)
    khi = np.ascontiguousarray(nodes["kmer_hi"], dtype=np.uint64) if "kmer_hi" in nodes else np.zeros(n, np.uint64)
    ll = np.ascontiguousarray(nodes["l"], dtype=np.uint32); rr = np.ascontiguousarray(nodes["r"], dtype=np.uint32)
    oo = np.ascontiguousarray(nodes["ord"], dtype=np.uint64)
    rpf = np.ascontiguousarray(reads_per_file, dtype=np.uint64)
    g = capi.dbg_growth_params(int(init_slots), float(load_factor), int(wide), int(max_double_times), int(buffer_reads))
    res = capi.dbg_growth_result()

    def call(arr, nul):
        capi.check(L.dbg_replay_growth(C.byref(g), rpf.ctypes.data, len(rpf), klo.ctypes.data, khi.ctypes.data, ll.ctypes.data,
                                       rr.ctypes.data, oo.ctypes.data, n, int(polyA_l), int(polyA_r), C.byref(res),
                                       arr.ctypes.data if arr is not None else None, nul.ctypes.data if nul is not None else None),
                   "dbg_replay_growth")

    call(None, None)
    plan = res.as_dict()
    if not layout or plan["truncated"]:
        return plan, None, None
    arr = np.zeros(plan["final_size"], dtype=NODE32 if wide else NODE16)
    nul = np.zeros(plan["final_size"] // 8 + 1, dtype=np.uint8)
    call(arr, nul)
    return res.as_dict(), arr, nul


# ---------------------------------------------------------------------------------------------------
# reference-shaped interface
# ---------------------------------------------------------------------------------------------------
@dataclass
class KmerSet:
    """struct KmerSet, kmerSet.h:88-99"""
    e_size: int
    size: int
    count: int
    count_conflict: int
    max: int
    load_factor: float
    iter_ptr: int
    array: np.ndarray
    nul_flag: np.ndarray
    del_flag: np.ndarray
    # extras (not in the reference struct)
    Total_reads_num: int = 0
    Kmer_total_num: int = 0
    occurrences: int = 0
    timings: dict = field(default_factory=dict)

    def is_entity_null(self, idx):   # kmerSet.h:144-147
        return 1 - ((int(self.nul_flag[idx // 8]) >> (7 - idx % 8)) & 1)

    def filled_slots(self):
        bits = np.unpackbits(self.nul_flag)[: self.size]
        return np.nonzero(bits)[0].astype(np.uint64)


def read_reads_file(path, Input_file_format=1):
    """The reference's reader (DBGgraph.cpp:244-272): -f 1: a line starting with '@' -> next line is the
    read, two more lines skipped; -f 2: a line starting with '>' -> next single line is the read.  Plain or
    gzip files (gzstream).  Returns (bases uint8, offs uint64)."""
    opener = gzip.open if _is_gzip(path) else open
    seqs = []
    with opener(path, "rb") as f:
        if Input_file_format == 1:
            for line in f:
                if line[:1] == b"@":
                    seqs.append(f.readline().rstrip(b"\n"))
                    f.readline(); f.readline()
        else:
            for line in f:
                if line[:1] == b">":
                    seqs.append(f.readline().rstrip(b"\n"))
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    return bases, offs


def _is_gzip(path):
    with open(path, "rb") as f:
        return f.read(2) == b"\x1f\x8b"


def build_debruijn_graph(reads_files, KmerSize=31, maxReadLen=250, Input_file_format=1, initHashSize=1.0,
                         hashLoadFactor=0.7, BufferNum=10000, maxDoubleHashTimes=10, device=0, track_order=True, log=None):
    """build_debruijn_graph (DBGgraph.cpp:364-430) with the build phase on the GPU.  `reads_files` is the
    list reading_file_list() returns (seqKmer.cpp:101-114), or a list of (bases, offs) arrays.
    Table growth (-e): the device table never grows; when -i cannot hold the nodes the build is redone with a larger
    device table, and when the reference would have enlarged its hash the returned KmerSet is laid out by the
    host-side replay of its doublings (replay_growth).  When -e is exhausted the reference stops reading the current
    file and takes only the first block of every later one (DBGgraph.cpp:346-350): the replay finds that point and the
    build is redone on exactly those reads -- the same flow as integration/DBGgraph_b200.cpp."""
    log = log or (lambda s: None)
    files = [f if isinstance(f, tuple) else read_reads_file(f, Input_file_format) for f in reads_files]
    ref_slots = init_slots_from_g(initHashSize)
    ref_P = 3 if ref_slots < 3 else capi.find_next_prime(ref_slots)
    lf = np.float32(hashLoadFactor)
    lf = np.float32(0.25) if lf <= 0 else (np.float32(0.75) if lf >= 1 else lf)
    ref_max = int(np.float32(ref_P) * lf)                       # uint64 * float -> float, kmerSet.cpp:115
    dev_slots = ref_slots
    limited = False
    for attempt in range(26):
        try:
            with DBGBuilder(K=KmerSize, max_read_len=maxReadLen, init_slots=dev_slots, load_factor=hashLoadFactor,
                            device=device, track_order=track_order) as b:
                for bases, offs in files:
                    b.submit(bases, offs)
                st = b.finalize()
                plan = None
                if st["count"] - 1 > ref_max and track_order:
                    nodes = b.dump_shard()
                    nodes["wide"] = b.wide
                    rpf = [len(offs) - 1 for _, offs in files]
                    plan, garr, gnul = replay_growth(nodes, rpf, ref_slots, hashLoadFactor, maxDoubleHashTimes, BufferNum,
                                                     st["polyA_l"], st["polyA_r"])
                    if plan["truncated"]:
                        if limited:
                            raise AssertionError("the truncated read set is truncated again")
                        tf, base, cut = plan["truncated_file"], 0, []
                        for f, (bases, offs) in enumerate(files):
                            n = len(offs) - 1
                            lim = plan["truncated_first_read"] - base if f == tf else (min(n, BufferNum) if f > tf else n)
                            cut.append((bases[: int(offs[lim])], offs[: lim + 1]))
                            base += n
                        log(f"-e {maxDoubleHashTimes} is exhausted inside file {tf}: rebuilding on exactly the reads the CPU program used")
                        files, limited, dev_slots = cut, True, ref_slots
                        continue
                grown = plan is not None and plan["doublings"] > 0
                if grown:
                    array, nul, size, mx = garr, gnul, plan["final_size"], plan["final_max"]
                else:
                    if st["array_size"] != ref_P:
                        raise capi.DbgError(capi.DBG_ERR_TABLE_FULL, "build_debruijn_graph",
                                            f"the input needs more than -i {initHashSize} and -e {maxDoubleHashTimes} allow")
                    array, nul = b.export_kmerset()
                    size, mx = st["array_size"], st["max_cutoff"]
                return KmerSet(e_size=32 if b.wide else 16, size=size, count=st["count"], count_conflict=st["conflict"],
                               max=mx, load_factor=st["load_factor"], iter_ptr=0, array=array, nul_flag=nul,
                               del_flag=np.zeros(size // 8 + 1, dtype=np.uint8), Total_reads_num=st["reads"],
                               Kmer_total_num=st["kmers_logged"], occurrences=st["occurrences"], timings=b.timings())
        except capi.DbgError as e:
            if e.code != capi.DBG_ERR_TABLE_FULL or attempt == 25:
                raise
            dev_slots = 2048 if dev_slots < 1024 else dev_slots * 2
            log(f"-i {initHashSize} cannot hold this input; rebuilding with a device table of {dev_slots} slots")
    raise AssertionError("unreachable")
