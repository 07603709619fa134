"""Contig seed index of link_scaffold (map_pair / map_reads) on the GPU: thin wrapper over the seedidx_* C ABI
(include/dbg_b200.h, csrc/seedidx.cu).  Mirrors the reference's interface for this path:

    init_kmerset(3 x contig length, 0.5) + chop_contig_to_kmerset   (link_scaffold/map_pair.cpp:97-125, map_func.cpp:119-172)
    get_align_seed(kset, read, 1, read.size(), ...)                  (map_func.cpp:181-237)

Everything computes on the GPU; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

SEED_NODE = np.dtype([("kmer", "<u8"), ("value", "<u8")])     # value = id | pos << 32 | freq << 62 | direct << 63 (kmerSet.h:53-60)


def _arrays(seqs):
    if isinstance(seqs, tuple):
        bases, offs = seqs
        return np.ascontiguousarray(bases, dtype=np.uint8), np.ascontiguousarray(offs, dtype=np.uint64)
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        np.cumsum([len(s) for s in seqs], out=offs[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(bases), offs


def unpack_value(value):
    """-> dict of arrays id, pos, freq, direct (the bit-fields of KmerNode, kmerSet.h:53-60)"""
    v = np.asarray(value, dtype=np.uint64)
    return dict(id=(v & np.uint64(0xFFFFFFFF)).astype(np.uint32), pos=((v >> np.uint64(32)) & np.uint64(0x3FFFFFFF)).astype(np.uint32),
                freq=((v >> np.uint64(62)) & np.uint64(1)).astype(np.uint8), direct=(v >> np.uint64(63)).astype(np.uint8))


class SeedIndex:
    def __init__(self, K=31, init_slots=3, load_factor=0.5, device=0):
        self.L = capi.load()
        self.K = int(K)
        self.h = C.c_void_p()
        rc = self.L.seedidx_create(C.byref(self.h), int(K), int(init_slots), float(load_factor), int(device))
        if rc != capi.DBG_OK:
            msg = self.L.seedidx_last_error().decode()
            if self.h:
                self.L.seedidx_destroy(self.h)
                self.h = C.c_void_p()
            raise capi.DbgError(rc, "seedidx_create", msg)
        self.size = self.count = self.max = None

    @classmethod
    def from_contigs(cls, contig_seqs, K=31, min_ctg_len=125, device=0):
        """what map_pair does with the contig file (map_pair.cpp:97-125): contigs shorter than -l are blanked but keep their
        index, the table gets 3 x the remaining length at load factor 0.5"""
        kept = [s if len(s) >= min_ctg_len else b"" for s in contig_seqs]
        idx = cls(K=K, init_slots=3 * sum(len(s) for s in kept), load_factor=0.5, device=device)
        idx.add_contigs(kept)
        idx.finalize()
        return idx

    def _check(self, rc, where):
        if rc != capi.DBG_OK:
            raise capi.DbgError(rc, where, self.L.seedidx_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.seedidx_destroy(self.h)
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

    def add_contigs(self, seqs):
        """chop_contig_to_kmerset: a list of bytes, or (bases uint8, offsets uint64)"""
        bases, offs = _arrays(seqs)
        if len(offs) > 1:
            self._check(self.L.seedidx_add_contigs(self.h, bases.ctypes.data if bases.size else None, offs.ctypes.data, len(offs) - 1),
                        "seedidx_add_contigs")

    def finalize(self):
        size, count, mx = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.seedidx_finalize(self.h, C.byref(size), C.byref(count), C.byref(mx)), "seedidx_finalize")
        self.size, self.count, self.max = size.value, count.value, mx.value
        return dict(size=self.size, count=self.count, max=self.max)

    def export(self):
        """-> (array[size] of SEED_NODE, nul_flag[size/8+1]): the KmerSet map_pair probes"""
        if self.size is None:
            self.finalize()
        arr = np.zeros(self.size, dtype=SEED_NODE)
        nul = np.zeros(self.size // 8 + 1, dtype=np.uint8)
        self._check(self.L.seedidx_export(self.h, arr.ctypes.data, nul.ctypes.data), "seedidx_export")
        return arr, nul

    def align(self, reads, search_start=None, seed_kmer_num=5):
        """get_align_seed for a batch of reads -> int32 [n, 6]: contig_id_index, seed_contig_start, seed_contig_end,
        seed_read_start, seed_read_end, ord('F'|'R'|'N')"""
        bases, offs = _arrays(reads)
        n = len(offs) - 1
        out = np.zeros((n, 6), dtype=np.int32)
        ss = None if search_start is None else np.ascontiguousarray(search_start, dtype=np.int32)
        if n:
            self._check(self.L.seedidx_align_reads(self.h, bases.ctypes.data if bases.size else None, offs.ctypes.data, n,
                                                   ss.ctypes.data if ss is not None else None, int(seed_kmer_num), out.ctypes.data),
                        "seedidx_align_reads")
        return out

    @property
    def launches(self):
        return int(self.L.seedidx_launch_count(self.h))
