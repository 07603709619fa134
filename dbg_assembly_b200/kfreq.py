"""K-mer frequency table for correct_error (what the external `kmerfreq` writes): thin wrapper over the kfreq_* C ABI.
Counting, spectrum and the 1-bit / 8-bit images are computed on the GPU; zlib compression runs on host threads."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class KmerFreq:
    def __init__(self, K=17, device=0, block_rank=0, block_count=1):
        self.L = capi.load()
        self.K = int(K)
        self.h = C.c_void_p()
        rc = self.L.kfreq_create(C.byref(self.h), int(K), int(device), int(block_rank), int(block_count))
        if rc != capi.DBG_OK:
            msg = self.L.kfreq_last_error().decode()
            if self.h:
                self.L.kfreq_destroy(self.h)
                self.h = C.c_void_p()
            raise capi.DbgError(rc, "kfreq_create", msg)

    def _check(self, rc, where):
        if rc != capi.DBG_OK:
            raise capi.DbgError(rc, where, self.L.kfreq_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.kfreq_destroy(self.h)
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

    def submit(self, bases, offs):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        if len(offs) > 1:
            self._check(self.L.kfreq_submit_reads(self.h, bases.ctypes.data if bases.size else None, offs.ctypes.data, len(offs) - 1),
                        "kfreq_submit_reads")

    def submit_device(self, d_bases_ptr, d_offs_ptr, n_reads, first_base, total_bases):
        self._check(self.L.kfreq_submit_reads_device(self.h, d_bases_ptr, d_offs_ptr, int(n_reads), int(first_base), int(total_bases)),
                    "kfreq_submit_reads_device")

    def reset(self):
        self._check(self.L.kfreq_reset(self.h), "kfreq_reset")

    def finalize(self):
        occ, reads = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.kfreq_finalize(self.h, C.byref(occ), C.byref(reads)), "kfreq_finalize")
        return dict(occurrences=occ.value, reads=reads.value)

    def index_range(self):
        lo, hi = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.kfreq_index_range(self.h, C.byref(lo), C.byref(hi)), "kfreq_index_range")
        return lo.value, hi.value

    def histogram(self):
        h = np.zeros(65536, dtype=np.uint64)
        self._check(self.L.kfreq_histogram(self.h, h.ctypes.data), "kfreq_histogram")
        return h

    def export(self, bits=1, cutoff=0):
        lo, hi = self.index_range()
        n = hi - lo
        out = np.zeros((n + 7) // 8 if bits == 1 else n, dtype=np.uint8)
        self._check(self.L.kfreq_export(self.h, int(bits), int(cutoff), out.ctypes.data), "kfreq_export")
        return out

    def write_cz(self, prefix, bits=1, cutoff=10):
        self._check(self.L.kfreq_write_cz(self.h, str(prefix).encode(), int(bits), int(cutoff)), "kfreq_write_cz")
