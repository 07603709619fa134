"""GPU parity tests of the multi-GPU flow on ONE device (what the round-end driver can run): n sharded contexts in one
process (dbg_assembly_b200.sharded.LocalShards) go through the default multi-GPU kernel path -- dbg_exchange_count_device /
dbg_exchange_scatter_device (PeerStagedSink, owner buckets and owner x slice buckets) with every "peer" receive buffer
local, dbg_insert_tuples_device / dbg_insert_sliced_device per owner -- then the cross-shard hand-off of the boundary
clusters and the windowed layout, and the merged table must be the reference's table bit for bit: nodes, link words, SLOT
LAYOUT (the oracle == reference -t 1), k-mer-0 node last.  Reference seam: DBG_contig/main.cpp:204-207 (kset handed to
the traversal), ownership DBGgraph.cpp:148."""
import numpy as np
import pytest

from conftest import load_golden, random_reads, reads_to_arrays
from test_gpu_build import image_to_dump, oracle_build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dbg():
    import dbg_assembly_b200 as m
    if m.capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (no CPU fallback exists)")
    return m


def deal_blocks(torch, bases, offs, n, rounds=1):
    """contiguous blocks of the read sequence dealt to n ranks, `rounds` exchange rounds"""
    db = torch.from_numpy(bases).cuda() if len(bases) else torch.zeros(16, dtype=torch.uint8, device="cuda")
    do = torch.from_numpy(offs.astype(np.int64)).cuda()
    n_reads = len(offs) - 1
    cuts = np.linspace(0, n_reads, n * rounds + 1).astype(np.int64)
    out = []
    for rd in range(rounds):
        blocks = []
        for r in range(n):
            r0, r1 = int(cuts[rd * n + r]), int(cuts[rd * n + r + 1])
            if r1 == r0:
                blocks.append(None)
                continue
            # the offsets tensor is indexed from read r0; offsets stay global (first_base = offs[r0])
            blocks.append((db, do[r0:], r1 - r0, int(offs[r0]), int(offs[r1] - offs[r0]), r0))
        out.append(blocks)
    return out, (db, do)


def run_sharded(dbg, files, K, R, P_req, n, by_slice=False, force_wide=False, rounds=1, load=0.7):
    import torch
    from dbg_assembly_b200.sharded import LocalShards
    bases = np.concatenate([f[0] for f in files]) if files else np.zeros(0, np.uint8)
    offs = [np.zeros(1, np.uint64)]
    for b, o in files:
        offs.append(o[1:] + offs[-1][-1])
    offs = np.concatenate(offs)
    ls = LocalShards(n, K, R, P_req, load_factor=load, force_wide=force_wide, by_slice=by_slice)
    try:
        rds, keep = deal_blocks(torch, bases, offs, n, rounds)
        for blocks in rds:
            ls.add_blocks(blocks)
        used_fallback = False
        try:
            stats = ls.finalize(layout=True)
            arr, nul = ls.export_kmerset(stats)
        except dbg.capi.DbgError as e:
            if e.code != dbg.capi.DBG_ERR_STATE:
                raise
            arr, nul, _ = ls.export_kmerset_fallback()
            used_fallback = True
        counts = sum(b.get_stats()["count"] for b in ls.b)
        return arr, nul, counts, used_fallback
    finally:
        ls.close()


def check_against_oracle(o, arr, nul, wide):
    e = o.dump()
    d = image_to_dump(arr, nul, o.size)
    for k in ("slot", "kmer", "l", "r") + (("kmer_hi",) if wide else ()):
        assert np.array_equal(d[k], e[k]), k


@pytest.mark.parametrize("n", [2, 3, 8])
@pytest.mark.parametrize("by_slice", [False, True])
@pytest.mark.parametrize("K,wide", [(31, False), (31, True), (63, True)])
def test_sharded_build_merges_into_the_reference_table(dbg, oracle_mod, monkeypatch, n, by_slice, K, wide):
    monkeypatch.setenv("DBG_B200_PART_SHIFT", "10")        # many table slices per shard at test size
    reads = random_reads(171 + n, 5000, 40, 150, genome_len=25000) + [b"A" * 70] * 30 + [b"T" * 64] * 7
    bases, offs = reads_to_arrays(reads)
    P_req = 400_000
    o = oracle_build(oracle_mod, [(bases, offs)], K, 150, P_req, wide=wide)
    arr, nul, counts, fb = run_sharded(dbg, [(bases, offs)], K, 150, P_req, n, by_slice=by_slice, force_wide=wide and K <= 31, rounds=2)
    assert not fb, "the windowed layout must handle an ordinary table"
    assert counts + 1 == o.count
    check_against_oracle(o, arr, nul, wide)
    o.close()


@pytest.mark.parametrize("name", ["contig_k31", "ragged_k31", "even_k16_two_files", "saturate_k21"])
@pytest.mark.parametrize("n", [2, 4])
def test_sharded_build_reproduces_the_golden_reference_tables(dbg, name, n):
    """the reference's own runs (tests/golden, -t 1): a build sharded over n contexts merges into the same table"""
    g = load_golden(name)
    arr, nul, counts, fb = run_sharded(dbg, g["files"], g["K"], g["R"], g["init_slots"], n, load=g["load"])
    assert counts + 1 == g["count"]
    d = image_to_dump(arr, nul, g["size"])
    for k, gk in (("slot", "slot"), ("kmer", "kmer"), ("l", "l"), ("r", "r")):
        assert np.array_equal(d[k], g[gk]), (name, n, k, fb)


@pytest.mark.parametrize("n,P_req,n_reads", [(2, 40_000, 600), (4, 9_000, 150), (8, 5_003, 90), (3, 2_000, 40)])
def test_sharded_layout_on_dense_and_tiny_tables(dbg, oracle_mod, n, P_req, n_reads):
    """boundary clusters that are long, that cascade, shards of a few hundred slots: the windowed layout either handles
    them or refuses (DBG_ERR_STATE) and the dump-merging fallback produces the same table"""
    reads = random_reads(900 + n, n_reads, 60, 100, genome_len=4000, err=0.03)
    bases, offs = reads_to_arrays(reads)
    o = oracle_build(oracle_mod, [(bases, offs)], 21, 100, P_req)
    assert o.count < o.size
    arr, nul, counts, fb = run_sharded(dbg, [(bases, offs)], 21, 100, P_req, n)
    check_against_oracle(o, arr, nul, False)
    o.close()


def test_sharded_layout_of_an_empty_build(dbg, oracle_mod):
    o = oracle_build(oracle_mod, [], 31, 100, 50_000)
    arr, nul, counts, fb = run_sharded(dbg, [], 31, 100, 50_000, 2)
    check_against_oracle(o, arr, nul, False)
    o.close()
