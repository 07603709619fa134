"""ctypes binding of libdbgb200.so (include/dbg_b200.h).  Plumbing only: every compute call goes to the
CUDA library; there is no Python/NumPy fallback.  Importing this module without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DBG_B200_LIB") or os.path.join(PKG, "libdbgb200.so")   # env override: tuning experiments only

DBG_OK = 0
DBG_ERR_INVALID, DBG_ERR_CUDA, DBG_ERR_NOMEM, DBG_ERR_TABLE_FULL, DBG_ERR_STATE, DBG_ERR_BUFFER = -1, -2, -3, -4, -5, -6


class DbgError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {_strerror(code)} ({code}) {detail}".strip())


class dbg_params(C.Structure):
    _fields_ = [("K", C.c_int32), ("max_read_len", C.c_int32), ("init_slots", C.c_uint64),
                ("load_factor", C.c_float), ("device", C.c_int32), ("track_order", C.c_int32),
                ("shard_rank", C.c_int32), ("shard_count", C.c_int32), ("force_wide", C.c_int32),
                ("payload_mode", C.c_int32), ("reserved", C.c_int32 * 4)]


class dbg_stats(C.Structure):
    _fields_ = [("array_size", C.c_uint64), ("max_cutoff", C.c_uint64), ("count", C.c_uint64),
                ("conflict", C.c_uint64), ("reads", C.c_uint64), ("kmers_logged", C.c_uint64),
                ("occurrences", C.c_uint64), ("polyA_l", C.c_uint64), ("polyA_r", C.c_uint64),
                ("shard_lo", C.c_uint64), ("shard_hi", C.c_uint64), ("load_factor", C.c_float), ("wide", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class dbg_synth_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genome_len", C.c_uint64), ("read_len", C.c_uint32), ("insert", C.c_uint32),
                ("err_per_2p24", C.c_uint32), ("n_per_2p24", C.c_uint32)]


class dbg_growth_params(C.Structure):
    _fields_ = [("init_slots", C.c_uint64), ("load_factor", C.c_float), ("wide", C.c_int32),
                ("max_double_times", C.c_uint64), ("buffer_reads", C.c_uint64)]


class dbg_growth_result(C.Structure):
    _fields_ = [("final_size", C.c_uint64), ("final_max", C.c_uint64), ("doublings", C.c_uint64), ("count", C.c_uint64),
                ("truncated", C.c_int32), ("truncated_file", C.c_uint32), ("truncated_first_read", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class dbg_checkpoint_header(C.Structure):
    _fields_ = [("magic", C.c_uint64), ("version", C.c_uint32), ("K", C.c_uint32), ("wide", C.c_uint32), ("load_factor", C.c_float),
                ("size", C.c_uint64), ("max_cutoff", C.c_uint64), ("count", C.c_uint64), ("count_conflict", C.c_uint64),
                ("reads", C.c_uint64), ("kmers_logged", C.c_uint64), ("records", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class kfreq_params(C.Structure):
    _fields_ = [("K", C.c_int32), ("device", C.c_int32), ("block_rank", C.c_int32), ("block_count", C.c_int32),
                ("reserved", C.c_int32 * 4)]


# every symbol include/dbg_b200.h declares; tests check that the library exports all of them
SYMBOLS = [
    "dbg_find_next_prime", "dbg_hash_code", "dbg_hash_code_wide", "dbg_strerror", "dbg_last_error",
    "dbg_device_count", "dbg_host_alloc", "dbg_host_free", "dbg_host_register", "dbg_host_unregister", "dbg_create", "dbg_destroy", "dbg_submit_reads",
    "dbg_submit_reads_device", "dbg_extract_tuples_device", "dbg_insert_tuples_device", "dbg_tuple_bytes",
    "dbg_peer_alloc", "dbg_peer_open", "dbg_peer_close", "dbg_peer_free", "dbg_exchange_count_device", "dbg_exchange_scatter_device", "dbg_insert_sliced_device", "dbg_partition_info",
    "dbg_get_polyA_counts", "dbg_set_polyA_counts", "dbg_finalize", "dbg_get_stats", "dbg_export_kmerset",
    "dbg_export_links", "dbg_dump_compact", "dbg_dump_shard", "dbg_device_image", "dbg_get_timings", "dbg_launch_count", "dbg_path_counts", "dbg_replay_growth",
    "dbg_exchange_scatter_opt_device", "dbg_exchange_scatter_undo", "dbg_insert_tuple_regions_device",
    "dbg_shard_tail_export", "dbg_shard_tail_import", "dbg_shard_slice_info", "dbg_export_shard_slice", "dbg_host_fix_nul_bytes", "dbg_host_polyA_insert",
    "dbg_mg_create", "dbg_mg_destroy", "dbg_mg_submit_reads", "dbg_mg_finalize", "dbg_mg_get_stats", "dbg_mg_export_kmerset", "dbg_mg_dump_nodes",
    "dbg_mg_info", "dbg_mg_last_error",
    "dbg_checkpoint_write", "dbg_checkpoint_read_header", "dbg_checkpoint_read",
    "dbg_reset", "dbg_set_stream", "dbg_synth_reads_host", "dbg_synth_reads_device", "dbg_measure_random_rmw",
    "kfreq_create", "kfreq_destroy", "kfreq_submit_reads", "kfreq_submit_reads_device", "kfreq_finalize",
    "kfreq_index_range", "kfreq_histogram", "kfreq_export", "kfreq_write_cz", "kfreq_last_error",
    "dbg_device_build_table", "seedidx_create", "seedidx_destroy", "seedidx_add_contigs", "seedidx_finalize", "seedidx_export",
    "seedidx_align_reads", "seedidx_launch_count", "seedidx_last_error",
    "dbg_export_info", "dbg_host_expand_nodes", "dbg_finish_export", "kfreq_reset", "dbg_exchange_scatter_pull_device", "dbg_insert_pull_device",
]

_lib = None


def load(build_if_missing: bool = True):
    """Load libdbgb200.so (building it in-tree with nvcc if it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise ImportError(f"{LIB_PATH} is missing: run `python -m dbg_assembly_b200.csrc.build`")
        from .csrc import build as _b
        _b.build()
    L = C.CDLL(LIB_PATH)
    u64, i32, vp = C.c_uint64, C.c_int32, C.c_void_p
    sig = {
        "dbg_find_next_prime": (u64, [u64]),
        "dbg_hash_code": (u64, [u64]),
        "dbg_hash_code_wide": (u64, [u64, u64]),
        "dbg_strerror": (C.c_char_p, [C.c_int]),
        "dbg_last_error": (C.c_char_p, []),
        "dbg_device_count": (C.c_int, []),
        "dbg_host_alloc": (C.c_int, [C.POINTER(vp), u64]),
        "dbg_host_free": (C.c_int, [vp]),
        "dbg_host_register": (C.c_int, [vp, u64]),
        "dbg_host_unregister": (C.c_int, [vp]),
        "dbg_create": (C.c_int, [C.POINTER(vp), C.POINTER(dbg_params)]),
        "dbg_destroy": (None, [vp]),
        "dbg_submit_reads": (C.c_int, [vp, vp, vp, u64]),
        "dbg_submit_reads_device": (C.c_int, [vp, vp, vp, u64, u64, u64, u64, vp]),
        "dbg_extract_tuples_device": (C.c_int, [vp, vp, vp, u64, u64, u64, u64, i32, vp, u64, vp, vp]),
        "dbg_insert_tuples_device": (C.c_int, [vp, vp, u64, vp]),
        "dbg_peer_alloc": (C.c_int, [vp, u64, C.POINTER(vp), vp]),
        "dbg_peer_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "dbg_peer_close": (C.c_int, [vp, vp]),
        "dbg_peer_free": (C.c_int, [vp, vp]),
        "dbg_exchange_count_device": (C.c_int, [vp, vp, vp, u64, u64, u64, i32, i32, vp, vp]),
        "dbg_exchange_scatter_device": (C.c_int, [vp, vp, vp, u64, u64, u64, u64, i32, i32, vp, vp, vp]),
        "dbg_insert_sliced_device": (C.c_int, [vp, vp, u64, vp, vp]),
        "dbg_partition_info": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(i32)]),
        "dbg_tuple_bytes": (C.c_int, [vp]),
        "dbg_get_polyA_counts": (C.c_int, [vp, vp]),
        "dbg_set_polyA_counts": (C.c_int, [vp, vp]),
        "dbg_finalize": (C.c_int, [vp, C.POINTER(dbg_stats)]),
        "dbg_get_stats": (C.c_int, [vp, C.POINTER(dbg_stats)]),
        "dbg_export_kmerset": (C.c_int, [vp, vp, vp]),
        "dbg_export_links": (C.c_int, [vp, i32, vp, vp, vp, vp, C.POINTER(u64), vp, C.POINTER(u64), vp]),
        "dbg_dump_compact": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, C.POINTER(u64)]),
        "dbg_dump_shard": (C.c_int, [vp, vp, vp, vp, vp, vp, C.POINTER(u64)]),
        "dbg_device_image": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "dbg_get_timings": (C.c_int, [vp, vp]),
        "dbg_launch_count": (u64, [vp]),
        "dbg_path_counts": (C.c_int, [vp, vp]),
        "dbg_replay_growth": (C.c_int, [C.POINTER(dbg_growth_params), vp, C.c_uint32, vp, vp, vp, vp, vp, u64, C.c_uint32, C.c_uint32,
                              C.POINTER(dbg_growth_result), vp, vp]),
        "dbg_exchange_scatter_opt_device": (C.c_int, [vp, vp, vp, u64, u64, u64, u64, i32, vp, u64, C.c_uint32, vp, vp]),
        "dbg_insert_tuple_regions_device": (C.c_int, [vp, vp, C.c_uint32, u64, vp, vp]),
        "dbg_exchange_scatter_undo": (C.c_int, [vp, vp]),
        "dbg_shard_tail_export": (C.c_int, [vp, vp, u64, C.POINTER(u64)]),
        "dbg_shard_tail_import": (C.c_int, [vp, vp, u64]),
        "dbg_shard_slice_info": (C.c_int, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp)]),
        "dbg_export_shard_slice": (C.c_int, [vp, vp, vp, vp]),
        "dbg_host_fix_nul_bytes": (C.c_int, [vp, vp, u64, i32, vp, u64]),
        "dbg_host_polyA_insert": (C.c_int, [vp, vp, u64, i32, C.c_uint32, C.c_uint32, C.POINTER(u64)]),
        "dbg_mg_create": (C.c_int, [C.POINTER(vp), C.POINTER(dbg_params), i32, vp]),
        "dbg_mg_destroy": (None, [vp]),
        "dbg_mg_submit_reads": (C.c_int, [vp, vp, vp, u64]),
        "dbg_mg_finalize": (C.c_int, [vp, C.POINTER(dbg_stats)]),
        "dbg_mg_get_stats": (C.c_int, [vp, C.POINTER(dbg_stats)]),
        "dbg_mg_export_kmerset": (C.c_int, [vp, vp, vp]),
        "dbg_mg_dump_nodes": (C.c_int, [vp, vp, vp, vp, vp, vp, C.POINTER(u64)]),
        "dbg_mg_info": (C.c_int, [vp, vp]),
        "dbg_mg_last_error": (C.c_char_p, []),
        "dbg_checkpoint_write": (C.c_int, [C.c_char_p, C.POINTER(dbg_checkpoint_header), vp, vp]),
        "dbg_checkpoint_read_header": (C.c_int, [C.c_char_p, C.POINTER(dbg_checkpoint_header)]),
        "dbg_checkpoint_read": (C.c_int, [C.c_char_p, vp, vp]),
        "dbg_reset": (C.c_int, [vp]),
        "dbg_set_stream": (C.c_int, [vp, vp]),
        "dbg_synth_reads_host": (C.c_int, [C.POINTER(dbg_synth_params), u64, u64, vp]),
        "dbg_synth_reads_device": (C.c_int, [C.POINTER(dbg_synth_params), u64, u64, vp, i32, vp]),
        "dbg_measure_random_rmw": (C.c_int, [i32, u64, u64, i32, C.POINTER(C.c_float)]),
        "kfreq_create": (C.c_int, [C.POINTER(vp), i32, i32, i32, i32]),
        "kfreq_destroy": (None, [vp]),
        "kfreq_submit_reads": (C.c_int, [vp, vp, vp, u64]),
        "kfreq_submit_reads_device": (C.c_int, [vp, vp, vp, u64, u64, u64]),
        "kfreq_finalize": (C.c_int, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "kfreq_reset": (C.c_int, [vp]),
        "kfreq_index_range": (C.c_int, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "kfreq_histogram": (C.c_int, [vp, vp]),
        "kfreq_export": (C.c_int, [vp, i32, i32, vp]),
        "kfreq_write_cz": (C.c_int, [vp, C.c_char_p, i32, i32]),
        "kfreq_last_error": (C.c_char_p, []),
        "dbg_device_build_table": (C.c_int, [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp)]),
        "seedidx_create": (C.c_int, [C.POINTER(vp), i32, u64, C.c_float, i32]),
        "seedidx_destroy": (None, [vp]),
        "seedidx_add_contigs": (C.c_int, [vp, vp, vp, u64]),
        "seedidx_finalize": (C.c_int, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "seedidx_export": (C.c_int, [vp, vp, vp]),
        "seedidx_align_reads": (C.c_int, [vp, vp, vp, u64, vp, i32, vp]),
        "seedidx_launch_count": (u64, [vp]),
        "seedidx_last_error": (C.c_char_p, []),
        "dbg_export_info": (C.c_int, [vp, vp]),
        "dbg_exchange_scatter_pull_device": (C.c_int, [vp, vp, vp, u64, u64, u64, u64, i32, vp, C.c_uint32, vp, vp]),
        "dbg_insert_pull_device": (C.c_int, [vp, vp, i32, C.c_uint32, vp, C.c_uint32, u64, vp]),
        "dbg_finish_export": (C.c_int, [vp, vp, vp, u64, C.POINTER(dbg_stats), vp, vp]),
        "dbg_host_expand_nodes": (u64, [vp, u64, vp, vp, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _strerror(code):
    try:
        return load().dbg_strerror(code).decode()
    except Exception:
        return "error"


def check(code, where):
    if code != DBG_OK:
        raise DbgError(code, where, load().dbg_last_error().decode())


def device_count() -> int:
    return int(load().dbg_device_count())


def find_next_prime(n: int) -> int:
    return int(load().dbg_find_next_prime(int(n)))


def hash_code(k: int) -> int:
    return int(load().dbg_hash_code(int(k)))


def host_fix_nul_bytes(array: np.ndarray, nul_flag: np.ndarray, P: int, wide: bool, slots):
    a = np.ascontiguousarray(slots, dtype=np.uint64)
    check(load().dbg_host_fix_nul_bytes(array.ctypes.data, nul_flag.ctypes.data, int(P), int(bool(wide)), a.ctypes.data, len(a)),
          "dbg_host_fix_nul_bytes")


def host_polyA_insert(array: np.ndarray, nul_flag: np.ndarray, P: int, wide: bool, l_link: int, r_link: int) -> int:
    s = C.c_uint64(0)
    check(load().dbg_host_polyA_insert(array.ctypes.data, nul_flag.ctypes.data, int(P), int(bool(wide)), int(l_link), int(r_link),
                                       C.byref(s)), "dbg_host_polyA_insert")
    return int(s.value)


def host_expand_nodes(bits: np.ndarray, n_slots: int, nodes: np.ndarray, array: np.ndarray, wide: bool) -> int:
    """array[s] = next node of `nodes` where bit s of `bits` (MSB first) is set, else 0 (the host half of the pipelined export)"""
    return int(load().dbg_host_expand_nodes(bits.ctypes.data, int(n_slots), nodes.ctypes.data, array.ctypes.data, int(bool(wide))))


def hash_code_wide(lo: int, hi: int) -> int:
    return int(load().dbg_hash_code_wide(int(lo), int(hi)))


class PinnedBuffer:
    """cudaHostAlloc'ed bytes exposed as a numpy array (front ends decode reads straight into it)."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        check(load().dbg_host_alloc(C.byref(self.ptr), int(nbytes)), "dbg_host_alloc")
        self.nbytes = int(nbytes)
        self.array = np.frombuffer((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr.value), dtype=np.uint8,
                                   count=self.nbytes)

    def close(self):
        if self.ptr:
            self.array = None
            load().dbg_host_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
