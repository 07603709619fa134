"""Deterministic synthetic reads -- thin wrapper over dbg_synth_reads_host/_device.
The generator is counter-based: host and device produce identical bytes for the same parameters."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

# named workloads of BASELINE.json / SURVEY.md 8d (fixed-length paired reads)
CONFIGS = {
    # C1: bundled-test shape: E. coli, 2 x 20x PE250 after correction (~0.1 % errors)
    "C1": dict(seed=1, genome_len=4_640_000, read_len=250, insert=400, err=0.001, n_rate=0.0, n_reads=742_648, K=31, max_read_len=250, init_g=0.1),
    # C2: E. coli-scale 4.6 Mb, 100x PE150, 1 % errors, 0.1 % N -- the configuration the metric is quoted on
    "C2": dict(seed=2, genome_len=4_600_000, read_len=150, insert=500, err=0.01, n_rate=0.001, n_reads=3_066_666, K=31, max_read_len=150, init_g=0.2),
    # C3: yeast-scale 12 Mb, 100 bp reads, K=63 (128-bit path)
    "C3": dict(seed=3, genome_len=12_000_000, read_len=100, insert=500, err=0.01, n_rate=0.0, n_reads=12_000_000, K=63, max_read_len=100, init_g=0.6),
    # C4: the correct_error K=17 frequency table on 50x of a 100 Mb genome over 8 GPUs -- PER GPU: 12.5 Mb of genome and
    # 4.17 M reads (x8 = 100 Mb, 33.3 M reads, 4.47e9 occurrences); the 4^17-entry table is sharded by .cz block
    "C4": dict(seed=4, genome_len=12_500_000, read_len=150, insert=500, err=0.01, n_rate=0.0, n_reads=4_166_666, K=17, max_read_len=150, init_g=0.0),
    # C5s: the human-scale configuration (3.1 Gb, 30x PE150, K=31, 8 GPUs) at 1/8 of its size PER GPU -- 48.4 Mb of genome and
    # 9.69 M reads per GPU (x8 GPUs = 387.5 Mb, 77.5 M reads, 9.3e9 occurrences); DESIGN.md has the memory plan of the full C5
    "C5s": dict(seed=5, genome_len=48_437_500, read_len=150, insert=500, err=0.01, n_rate=0.0, n_reads=9_687_500, K=31, max_read_len=150, init_g=0.7),
}


def make_params(seed, genome_len, read_len, insert=500, err=0.01, n_rate=0.0):
    p = capi.dbg_synth_params()
    p.seed, p.genome_len, p.read_len, p.insert = int(seed), int(genome_len), int(read_len), int(insert)
    p.err_per_2p24 = int(round(err * (1 << 24)))
    p.n_per_2p24 = int(round(n_rate * (1 << 24)))
    return p


def reads_host(params, first_read, n_reads, out=None):
    """-> (bases uint8 [n_reads*read_len], offs uint64 [n_reads+1])"""
    L = capi.load()
    n = int(n_reads) * int(params.read_len)
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    capi.check(L.dbg_synth_reads_host(C.byref(params), int(first_read), int(n_reads), out.ctypes.data), "dbg_synth_reads_host")
    offs = np.arange(int(n_reads) + 1, dtype=np.uint64) * np.uint64(params.read_len)
    return out, offs


def reads_device(params, first_read, n_reads, d_out_ptr, device=0, stream=None):
    L = capi.load()
    capi.check(L.dbg_synth_reads_device(C.byref(params), int(first_read), int(n_reads), d_out_ptr, int(device), stream),
               "dbg_synth_reads_device")
