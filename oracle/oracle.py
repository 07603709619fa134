"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the CPU oracle (oracle/liboracle.so) and runner for
the compiled reference (oracle/_ref/*).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package dbg_assembly_b200 never does (tests/test_layout.py greps for that).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_build_driver")
REF_CONTIG = os.path.join(REF_DIR, "debruijn_contig_ref")
REF_ELF = os.path.join(REF_DIR, "debruijn_contig_elf")
B200_CONTIG = os.path.join(REF_DIR, "debruijn_contig_b200")
KMER_TABLE_DRIVER = os.path.join(REF_DIR, "ref_kmer_table_driver")   # reference's construct_ref_kmer_table
CORRECT_ELF = os.path.join(REF_DIR, "correct_error_reads_elf")   # shipped consumer of the .cz table

_lib = None


def build():
    """(Re)build liboracle.so and, when /root/reference exists, oracle/_ref/*."""
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    u64, u8p, u64p, u32p, i64p = C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
    L.orc_base_code.restype = C.c_int; L.orc_base_code.argtypes = [C.c_ubyte]
    L.orc_seq2bit.restype = u64; L.orc_seq2bit.argtypes = [C.c_char_p, C.c_int]
    L.orc_rev_com_kbit.restype = u64; L.orc_rev_com_kbit.argtypes = [u64, C.c_int]
    L.orc_hash_code.restype = u64; L.orc_hash_code.argtypes = [u64]
    L.orc_hash_code_wide.restype = u64; L.orc_hash_code_wide.argtypes = [u64, u64]
    L.orc_rev_com_wide.restype = None; L.orc_rev_com_wide.argtypes = [u64, u64, C.c_int, u64p, u64p]
    L.orc_is_prime.restype = C.c_int; L.orc_is_prime.argtypes = [u64]
    L.orc_find_next_prime.restype = u64; L.orc_find_next_prime.argtypes = [u64]
    L.orc64_parse_read.restype = C.c_int
    L.orc64_parse_read.argtypes = [C.c_char_p, u64, C.c_int, C.c_int, u64p, u8p, u8p]
    L.orc128_parse_read.restype = C.c_int
    L.orc128_parse_read.argtypes = [C.c_char_p, u64, C.c_int, C.c_int, u64p, u64p, u8p, u8p]
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, u64, C.c_float, u64, u64, C.c_int]
    L.orc_destroy.restype = None; L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_add_file.restype = u64; L.orc_add_file.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64]
    L.orc_finish.restype = None; L.orc_finish.argtypes = [C.c_void_p]
    for name in ("size", "count", "conflict", "max", "doublings", "total_reads", "kmers_logged", "occurrences"):
        f = getattr(L, "orc_" + name); f.restype = u64; f.argtypes = [C.c_void_p]
    L.orc_node_bytes.restype = C.c_int; L.orc_node_bytes.argtypes = [C.c_void_p]
    L.orc_array.restype = C.c_void_p; L.orc_array.argtypes = [C.c_void_p]
    L.orc_nul_flag.restype = C.c_void_p; L.orc_nul_flag.argtypes = [C.c_void_p]
    L.orc_dump.restype = u64; L.orc_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.orc_calculate_kmer_links.restype = None
    L.orc_calculate_kmer_links.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
    L.orc_kfreq_count.restype = None
    L.orc_kfreq_count.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_int, C.c_void_p]
    _lib = L
    return L


def kfreq_count(bases, offs, K):
    """canonical K-mer counts of every read position, direct-index u32 table of 4^K entries"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    if bases.size == 0:
        bases = np.zeros(1, dtype=np.uint8)
    counts = np.zeros(1 << (2 * K), dtype=np.uint32)
    lib().orc_kfreq_count(bases.ctypes.data, offs.ctypes.data, len(offs) - 1, K, counts.ctypes.data)
    return counts


def ref_kmer_table(genome_fasta, K):
    """the reference's own both-strand 1-bit table of a genome FASTA (correct_error/simulate_lowfreq_kmer.cpp:189-260
    through oracle/_ref/ref_kmer_table_driver) -> bit array (uint8 0/1) of 4^K entries, MSB-first bytes unpacked"""
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "t.bits")
        subprocess.run([KMER_TABLE_DRIVER, str(K), genome_fasta, out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        raw = np.fromfile(out, dtype=np.uint8)
    return np.unpackbits(raw)[: 1 << (2 * K)]


def load_cz_1bit(prefix, K):
    """What correct_error's loader does with a 1-bit .cz (main_parallel_senior.cpp:334-408): inflate block i to
    KmerFreq + i * 1 MiB, then set the reverse-complement bit of every set canonical entry.  -> bit array (uint8 0/1)"""
    import zlib
    total = 1 << (2 * K)
    table = np.zeros(total // 8 if total >= 8 else 1, dtype=np.uint8)
    lens = [int(x) for x in open(prefix + ".kmer.freq.cz.len").read().split()]
    raw = open(prefix + ".kmer.freq.cz", "rb").read()
    pos = 0
    for i, ln in enumerate(lens):
        blk = zlib.decompress(raw[pos:pos + ln]); pos += ln
        a = np.frombuffer(blk, dtype=np.uint8)
        table[i * (1 << 20): i * (1 << 20) + len(a)] = a
    assert pos == len(raw)
    bits = np.unpackbits(table)[:total]              # MSB first == get_freq (seqKmer.cpp:102-106)
    idx = np.nonzero(bits)[0].astype(np.uint64)
    L = lib()
    rc = np.array([L.orc_rev_com_kbit(int(i), K) for i in idx.tolist()], dtype=np.uint64)
    out = bits.copy()
    out[rc[idx <= rc].astype(np.int64)] = 1          # thread_setrevcompkmer: only for i <= rc(i)
    return out, bits


def load_cz_8bit(prefix, K, low_freq_cutoff):
    """correct_error/main.cpp:161-220: bytes > cutoff set the bit of the k-mer and of its reverse complement"""
    import zlib
    total = 1 << (2 * K)
    lens = [int(x) for x in open(prefix + ".kmer.freq.cz.len").read().split()]
    raw = open(prefix + ".kmer.freq.cz", "rb").read()
    vals = np.zeros(total, dtype=np.uint8)
    pos = 0
    for i, ln in enumerate(lens):
        a = np.frombuffer(zlib.decompress(raw[pos:pos + ln]), dtype=np.uint8); pos += ln
        vals[i * (8 << 20): i * (8 << 20) + len(a)] = a
    bits = np.zeros(total, dtype=np.uint8)
    hi = np.nonzero(vals > low_freq_cutoff)[0]
    bits[hi] = 1
    L = lib()
    for i in hi.tolist():
        bits[L.orc_rev_com_kbit(int(i), K)] = 1
    return bits, vals


NODE16 = np.dtype([("kmer", "<u8"), ("l", "<u4"), ("r", "<u4")])
NODE32 = np.dtype([("kmer", "<u8"), ("kmer_hi", "<u8"), ("l", "<u4"), ("r", "<u4"), ("pad", "<u8")])


class OracleGraph:
    """Sequential CPU build == reference with -t 1 (see dbg_oracle.h for the pinning status)."""

    def __init__(self, K, max_read_len, init_slots, load_factor=0.7, max_double_times=10,
                 buffer_reads=10000, wide=False):
        self.L = lib()
        self.K, self.wide = K, bool(wide)
        self.h = self.L.orc_create(K, max_read_len, int(init_slots), float(load_factor),
                                   int(max_double_times), int(buffer_reads), int(bool(wide)))
        if not self.h:
            raise ValueError("orc_create rejected the parameters")
        self.finished = False

    def add_file(self, bases: np.ndarray, offs: np.ndarray) -> int:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        if bases.size == 0:
            bases = np.zeros(1, dtype=np.uint8)
        return int(self.L.orc_add_file(self.h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1))

    def finish(self):
        if not self.finished:
            self.L.orc_finish(self.h)
            self.finished = True
        return self

    def __getattr__(self, name):
        if name in ("size", "count", "conflict", "max", "doublings", "total_reads", "kmers_logged", "occurrences"):
            return int(getattr(self.L, "orc_" + name)(self.h))
        raise AttributeError(name)

    def array(self) -> np.ndarray:
        n = self.size
        dt = NODE32 if self.wide else NODE16
        p = self.L.orc_array(self.h)
        buf = (C.c_uint8 * (n * dt.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=n).copy()

    def nul_flag(self) -> np.ndarray:
        n = self.size // 8 + 1
        buf = (C.c_uint8 * n).from_address(self.L.orc_nul_flag(self.h))
        return np.frombuffer(buf, dtype=np.uint8, count=n).copy()

    def dump(self):
        """filled slots in slot order -> dict of arrays slot, kmer, kmer_hi, l, r"""
        n = self.count
        out = {k: np.zeros(n, dtype=dt) for k, dt in
               (("slot", np.uint64), ("kmer", np.uint64), ("kmer_hi", np.uint64), ("l", np.uint32), ("r", np.uint32))}
        m = self.L.orc_dump(self.h, *(out[k].ctypes.data for k in ("slot", "kmer", "kmer_hi", "l", "r")))
        assert m == n, (m, n)
        return out

    def kmer_links(self, freq_cutoff=2):
        P, n = self.size, self.count
        klink = np.zeros(2 * P, dtype=np.uint8)
        del_flag = np.zeros(P // 8 + 1, dtype=np.uint8)
        depth = np.zeros(256, dtype=np.int64)
        tips = np.zeros(max(n, 1), dtype=np.uint64)
        branches = np.zeros(max(n, 1), dtype=np.uint64)
        nt, nb = C.c_uint64(0), C.c_uint64(0)
        stats = np.zeros(3, dtype=np.int64)
        self.L.orc_calculate_kmer_links(self.h, int(freq_cutoff), klink.ctypes.data, del_flag.ctypes.data,
                                        depth.ctypes.data, tips.ctypes.data, C.addressof(nt),
                                        branches.ctypes.data, C.addressof(nb), stats.ctypes.data)
        return dict(klink=klink, del_flag=del_flag, depth_stat=depth, tips=tips[:nt.value].copy(),
                    branches=branches[:nb.value].copy(), total=int(stats[0]), deleted=int(stats[1]),
                    linear=int(stats[2]))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parse_read(read: bytes, K: int, max_read_len: int, wide=False):
    L = lib()
    n = max(0, min(len(read), max_read_len) - K + 1)
    cap = max(n, 1)
    lo = np.zeros(cap, np.uint64); hi = np.zeros(cap, np.uint64)
    lb = np.zeros(cap, np.uint8); rb = np.zeros(cap, np.uint8)
    u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    if wide:
        m = L.orc128_parse_read(read, len(read), K, max_read_len, lo.ctypes.data_as(u64p), hi.ctypes.data_as(u64p),
                                lb.ctypes.data_as(u8p), rb.ctypes.data_as(u8p))
    else:
        m = L.orc64_parse_read(read, len(read), K, max_read_len, lo.ctypes.data_as(u64p),
                               lb.ctypes.data_as(u8p), rb.ctypes.data_as(u8p))
    assert m == n
    return lo[:m], hi[:m], lb[:m], rb[:m]


# ---------------------------------------------------------------------------------------------------
# the compiled reference (only where oracle/_ref exists: this container, or a box the snapshot reached)
# ---------------------------------------------------------------------------------------------------
def have_reference() -> bool:
    return os.access(REF_DRIVER, os.X_OK)


def write_fasta(path: str, bases: np.ndarray, offs: np.ndarray):
    """one-line FASTA (-f 2), the format the reference's reader expects (DBGgraph.cpp:261-271)"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offs = np.asarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    # each record: ">\n" + seq + "\n"
    rec_len = lens + 3
    out_offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rec_len, out=out_offs[1:])
    out = np.empty(int(out_offs[-1]), dtype=np.uint8)
    out[out_offs[:-1]] = ord(">")
    out[out_offs[:-1] + 1] = ord("\n")
    out[out_offs[1:] - 1] = ord("\n")
    # scatter sequences: index arithmetic instead of a python loop
    if bases.size:
        dst = np.arange(int(offs[-1]), dtype=np.int64)
        read_id = np.repeat(np.arange(n, dtype=np.int64), lens)
        dst += (out_offs[:-1] + 2 - offs[:-1].astype(np.int64))[read_id]
        out[dst] = bases[: int(offs[-1])]
    with open(path, "wb") as f:
        f.write(out.tobytes())


def run_ref_build(files, K, max_read_len, init_g, threads=1, load=0.7, max_double=10, buffer_reads=10000,
                  fmt=2, dump=True, timeout=600):
    """Run the reference build phase (oracle/_ref/ref_build_driver).  Returns (stats, dump dict|None)."""
    with tempfile.TemporaryDirectory() as td:
        dump_path = os.path.join(td, "dump.bin")
        cmd = [REF_DRIVER, "-k", str(K), "-r", str(max_read_len), "-f", str(fmt), "-t", str(threads),
               "-i", repr(float(init_g)), "-l", repr(float(load)), "-e", str(max_double), "-b", str(buffer_reads)]
        if dump:
            cmd += ["-d", dump_path]
        cmd += list(files)
        p = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout)
        stats = json.loads(p.stdout.decode().strip().splitlines()[-1])
        d = None
        if dump:
            raw = open(dump_path, "rb").read()
            magic, size, count = struct.unpack("<QQQ", raw[:24])
            assert magic == 0x4442474B53455431
            rec = np.frombuffer(raw, dtype=np.dtype([("slot", "<u8"), ("kmer", "<u8"), ("l", "<u4"), ("r", "<u4")]),
                                offset=24, count=count)
            d = dict(size=size, count=count, slot=rec["slot"].copy(), kmer=rec["kmer"].copy(),
                     l=rec["l"].copy(), r=rec["r"].copy())
        return stats, d


# ---------------------------------------------------------------------------------------------------
# f-4: contig seed index of link_scaffold (oracle/seed_oracle.c; reference through oracle/_ref/ref_seed_driver)
# ---------------------------------------------------------------------------------------------------
SEED_DRIVER = os.path.join(REF_DIR, "ref_seed_driver")
SEED_NODE = np.dtype([("kmer", "<u8"), ("value", "<u8")])     # value = id | pos << 32 | freq << 62 | direct << 63


def _seed_lib():
    L = lib()
    if not hasattr(L, "_seed_ready"):
        u64 = C.c_uint64
        L.orc_seed_create.restype = C.c_void_p; L.orc_seed_create.argtypes = [C.c_int, u64, C.c_float]
        L.orc_seed_destroy.restype = None; L.orc_seed_destroy.argtypes = [C.c_void_p]
        L.orc_seed_add_contigs.restype = C.c_int; L.orc_seed_add_contigs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64, u64]
        for name in ("size", "count", "max", "conflict"):
            f = getattr(L, "orc_seed_" + name); f.restype = u64; f.argtypes = [C.c_void_p]
        L.orc_seed_array.restype = C.c_void_p; L.orc_seed_array.argtypes = [C.c_void_p]
        L.orc_seed_nul_flag.restype = C.c_void_p; L.orc_seed_nul_flag.argtypes = [C.c_void_p]
        L.orc_seed_align.restype = None
        L.orc_seed_align.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L._seed_ready = True
    return L


def seqs_to_arrays(seqs):
    """list of bytes -> (uint8 bases, uint64 offsets)"""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=offs[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, dtype=np.uint8)
    return bases, offs


class SeedOracle:
    """init_kmerset + chop_contig_to_kmerset + get_align_seed, sequential (link_scaffold/map_func.cpp, kmerSet.cpp)"""

    def __init__(self, K, init_size, load_factor=0.5):
        self.L = _seed_lib()
        self.K = K
        self.h = self.L.orc_seed_create(K, int(init_size), float(load_factor))

    def add_contigs(self, seqs, id0=0):
        bases, offs = seqs_to_arrays(seqs)
        if bases.size == 0:
            bases = np.zeros(1, dtype=np.uint8)
        return int(self.L.orc_seed_add_contigs(self.h, bases.ctypes.data, offs.ctypes.data, len(seqs), int(id0)))

    def __getattr__(self, name):
        if name in ("size", "count", "max", "conflict"):
            return int(getattr(self.L, "orc_seed_" + name)(self.h))
        raise AttributeError(name)

    def array(self):
        n = self.size
        buf = (C.c_uint8 * (n * 16)).from_address(self.L.orc_seed_array(self.h))
        return np.frombuffer(buf, dtype=SEED_NODE, count=n).copy()

    def nul_flag(self):
        n = self.size // 8 + 1
        buf = (C.c_uint8 * n).from_address(self.L.orc_seed_nul_flag(self.h))
        return np.frombuffer(buf, dtype=np.uint8, count=n).copy()

    def align(self, reads, seed_kmer_num=5):
        """get_align_seed(read, 1, len) for every read long enough (map_pair.cpp:284) -> int32 [n, 6]"""
        out = np.full((len(reads), 6), -1, dtype=np.int32)
        out[:, 5] = ord("N")
        rec = np.zeros(6, dtype=np.int32)
        for i, r in enumerate(reads):
            if len(r) >= self.K + seed_kmer_num:
                self.L.orc_seed_align(self.h, r, len(r), 1, len(r), int(seed_kmer_num), rec.ctypes.data)
                out[i] = rec
        return out

    def close(self):
        if self.h:
            self.L.orc_seed_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def have_seed_reference() -> bool:
    return os.access(SEED_DRIVER, os.X_OK)


def run_ref_seed(contig_names, contig_seqs, reads, K, min_ctg_len, seed_kmer_num=5, hash_size=None, timeout=600):
    """the reference's own seed index (oracle/_ref/ref_seed_driver) -> dict(size, count, max, conflict, array, nul, seeds)"""
    with tempfile.TemporaryDirectory() as td:
        fa, rd, out = os.path.join(td, "c.fa"), os.path.join(td, "r.txt"), os.path.join(td, "o")
        with open(fa, "wb") as f:
            for n, s in zip(contig_names, contig_seqs):
                f.write(b">" + n + b"\n" + s + b"\n")
        with open(rd, "wb") as f:
            for r in reads:
                f.write(r + b"\n")
        cmd = [SEED_DRIVER, str(K), str(min_ctg_len), str(seed_kmer_num), fa, rd, out]
        if hash_size is not None:
            cmd.append(str(int(hash_size)))
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
        raw = open(out + ".table", "rb").read()
        size, count, mx, conflict = struct.unpack("<QQQQ", raw[:32])
        array = np.frombuffer(raw, dtype=SEED_NODE, offset=32, count=size).copy()
        nul = np.frombuffer(raw, dtype=np.uint8, offset=32 + 16 * size, count=size // 8 + 1).copy()
        seeds = np.fromfile(out + ".seeds", dtype=np.int32).reshape(-1, 6)
    return dict(size=size, count=count, max=mx, conflict=conflict, array=array, nul=nul, seeds=seeds)
