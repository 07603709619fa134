"""dbg_assembly_b200 -- B200-native (sm_100a) De Bruijn graph build for fanagislab/DBG_assembly.

Only the data-parallel hot path lives here: 2-bit encode + canonical k-mers
(seqKmer), hash-set counting with the per-neighbour-base link counters (kmerSet + the build phase of
DBGgraph), the low-frequency link pass and the compacted dump -- as hand-written CUDA behind the C ABI of
include/dbg_b200.h.  Everything computes on the GPU; there is no CPU fallback.
"""
from .graph import DBGBuilder, MultiGpuBuilder, KmerSet, build_debruijn_graph, read_reads_file, NODE16, NODE32  # noqa: F401
from . import capi, synth  # noqa: F401
from .seedidx import SeedIndex  # noqa: F401

__all__ = ["DBGBuilder", "MultiGpuBuilder", "KmerSet", "build_debruijn_graph", "read_reads_file", "capi", "synth", "NODE16", "NODE32", "SeedIndex"]
