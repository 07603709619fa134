"""CPU tests of the host-side replay of the reference's table growth (dbg_replay_growth): given the
finished graph as the GPU holds it -- nodes + first-occurrence ordinals -- it must produce the table the REFERENCE ended
with, doublings included.  Checked against the golden fixtures made by running the reference itself (slot layout of the
enlarge x2 run, of all the no-enlarge runs, and the "memory reach the maximum" run, which a replay must refuse)."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_ALL, load_golden


def full_build_nodes(orc, g):
    """what a full GPU build knows: every canonical k-mer of ALL reads with the ordinal (read << 16 | j) of its first
    occurrence and its saturating link words (from the oracle run without growth limits)"""
    L = orc.lib()
    K, R = g["K"], g["R"]
    u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    first = {}
    read_index = 0
    reads_per_file = []
    for bases, offs in g["files"]:
        reads_per_file.append(len(offs) - 1)
        for i in range(len(offs) - 1):
            rd = bases[int(offs[i]):int(offs[i + 1])].tobytes()
            m = max(len(rd), 1)
            lo = np.zeros(m, np.uint64); lb = np.zeros(m, np.uint8); rb = np.zeros(m, np.uint8)
            n = L.orc64_parse_read(rd, len(rd), K, R, lo.ctypes.data_as(u64p), lb.ctypes.data_as(u8p), rb.ctypes.data_as(u8p))
            for j in range(n):
                k = int(lo[j])
                if k != 0 and k not in first:
                    first[k] = (read_index << 16) | j
            read_index += 1
    # link words: a big table that never grows
    o = orc.OracleGraph(K, R, 4_000_000, 0.7, 10, 1 << 40)
    for bases, offs in g["files"]:
        o.add_file(bases, offs)
    o.finish()
    d = o.dump()
    o.close()
    nz = d["kmer"] != 0
    polyA = (int(d["l"][~nz][0]), int(d["r"][~nz][0]))
    kmer = d["kmer"][nz]
    nodes = dict(kmer=kmer, kmer_hi=np.zeros(len(kmer), np.uint64), l=d["l"][nz], r=d["r"][nz],
                 ord=np.array([first[int(k)] for k in kmer.tolist()], dtype=np.uint64))
    assert len(first) == len(kmer)
    return nodes, reads_per_file, polyA


@pytest.mark.parametrize("name", GOLDEN_ALL)
def test_replay_reproduces_the_reference_table(oracle_mod, name):
    from dbg_assembly_b200.graph import replay_growth
    g = load_golden(name)
    nodes, rpf, polyA = full_build_nodes(oracle_mod, g)
    plan, arr, nul = replay_growth(nodes, rpf, g["init_slots"], g["load"], g["max_double"], g["B"], polyA[0], polyA[1])
    if name == "maxmem_k25":
        # -e 1 exhausted: the reference dropped the rest of file 0 -- not reproducible from a full build
        assert plan["truncated"] == 1 and plan["truncated_file"] == 0 and arr is None
        assert plan["doublings"] == 1
        return
    assert plan["truncated"] == 0
    assert (plan["final_size"], plan["final_max"], plan["doublings"], plan["count"]) == (g["size"], g["max"], g["doublings"], g["count"])
    slot = np.nonzero(np.unpackbits(nul)[:plan["final_size"]])[0]
    assert np.array_equal(slot.astype(np.uint64), g["slot"])
    assert np.array_equal(arr["kmer"][slot], g["kmer"])
    assert np.array_equal(arr["l_link"][slot], g["l"]) and np.array_equal(arr["r_link"][slot], g["r"])
    mask = np.ones(plan["final_size"], bool); mask[slot] = False
    assert not arr["kmer"][mask].any() and not arr["l_link"][mask].any()


def test_replay_matches_oracle_on_random_growth(oracle_mod):
    """several doublings, two files, block sizes that do and do not divide the file lengths; forced-wide keys too"""
    from dbg_assembly_b200.graph import replay_growth
    from conftest import random_reads, reads_to_arrays
    for seed, B, init_slots, n0, n1 in ((1, 7, 600, 140, 63), (2, 10, 900, 200, 100), (3, 5, 400, 75, 50)):
        files = [reads_to_arrays(random_reads(seed, n0, 30, 70, genome_len=3000)), reads_to_arrays(random_reads(seed + 50, n1, 30, 70, genome_len=3000))]
        g = dict(K=21, R=70, files=files)
        nodes, rpf, polyA = full_build_nodes(oracle_mod, g)
        o = oracle_mod.OracleGraph(21, 70, init_slots, 0.7, 10, B)
        for b, of in files:
            o.add_file(b, of)
        o.finish()
        e = o.dump()
        for wide in (False, True):
            nn = dict(nodes, wide=wide)
            plan, arr, nul = replay_growth(nn, rpf, init_slots, 0.7, 10, B, polyA[0], polyA[1])
            assert plan["doublings"] == o.doublings >= 1 and plan["final_size"] == o.size and plan["final_max"] == o.max
            slot = np.nonzero(np.unpackbits(nul)[:o.size])[0]
            assert np.array_equal(slot.astype(np.uint64), e["slot"]) and np.array_equal(arr["kmer"][slot], e["kmer"])
            assert np.array_equal(arr["l_link"][slot], e["l"]) and np.array_equal(arr["r_link"][slot], e["r"])
        o.close()


def test_replay_reports_a_full_table_instead_of_spinning(oracle_mod):
    """one block that holds more new k-mers than the table has slots: the reference probes forever; the replay says so"""
    from dbg_assembly_b200 import capi
    from dbg_assembly_b200.graph import replay_growth
    n = 500
    nodes = dict(kmer=np.arange(1, n + 1, dtype=np.uint64), kmer_hi=np.zeros(n, np.uint64), l=np.zeros(n, np.uint32),
                 r=np.zeros(n, np.uint32), ord=np.arange(n, dtype=np.uint64))          # all first seen in read 0
    with pytest.raises(capi.DbgError):
        replay_growth(nodes, [10], 100, 0.7, 10, 5)


def test_replay_matches_a_live_reference_run_that_enlarges(oracle_mod, tmp_path):
    """the reference itself (oracle/_ref/ref_build_driver, -t 1) on two files with -i far too small and -b 100: it
    enlarges three times; the replay from the full-build nodes gives the same table, slot for slot"""
    from dbg_assembly_b200.graph import replay_growth
    from conftest import random_reads, reads_to_arrays
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref not built")
    reads = random_reads(98, 3000, 100, 100, genome_len=15000, err=0.004, n_rate=0.0, lower=0.0)
    files = [reads_to_arrays(reads[:1800]), reads_to_arrays(reads[1800:])]
    paths = []
    for i, (b, o) in enumerate(files):
        p = str(tmp_path / f"f{i}.fa"); oracle_mod.write_fasta(p, b, o); paths.append(p)
    init_g, B = 0.000008, 100
    stats, ref = oracle_mod.run_ref_build(paths, 25, 100, init_g, threads=1, max_double=10, buffer_reads=B)
    nodes, rpf, polyA = full_build_nodes(oracle_mod, dict(K=25, R=100, files=files))
    plan, arr, nul = replay_growth(nodes, rpf, int(init_g * 1e9), 0.7, 10, B, polyA[0], polyA[1])
    assert plan["doublings"] == 3 and plan["final_size"] == ref["size"] == 64151 and plan["count"] == ref["count"]
    slot = np.nonzero(np.unpackbits(nul)[:plan["final_size"]])[0]
    assert np.array_equal(slot.astype(np.uint64), ref["slot"]) and np.array_equal(arr["kmer"][slot], ref["kmer"])
    assert np.array_equal(arr["l_link"][slot], ref["l"]) and np.array_equal(arr["r_link"][slot], ref["r"])
    # -e 1: the reference drops reads -> the replay refuses
    plan1, arr1, _ = replay_growth(nodes, rpf, int(init_g * 1e9), 0.7, 1, B, polyA[0], polyA[1])
    assert plan1["truncated"] == 1 and plan1["truncated_first_read"] == 300 and arr1 is None


@pytest.mark.parametrize("name", ["contig_k31", "ragged_k31", "even_k16_two_files"])
def test_shard_dump_merge_rebuilds_the_reference_table(oracle_mod, name):
    """fallback of the cross-shard layout (sharded.merge_dumps_host / dbg_mg_export_kmerset): the nodes of a build dealt
    to 3 "shards" by owner slot range -- with some overflow nodes listed twice (adopted by the neighbour AND present in the
    tail blob), one shard's overflow nodes only inside a raw tail blob (32-byte build nodes: key, ~ordinal, eight half-float
    counters) -- merged by replaying first occurrences on the host == the reference's table (golden fixture), k-mer-0 node
    included.  CPU only: pins the host-side merge and the blob decoder."""
    from dbg_assembly_b200 import capi
    from dbg_assembly_b200.sharded import blob_margin_nodes, merge_dumps_host
    g = load_golden(name)
    nodes, _, polyA = full_build_nodes(oracle_mod, g)
    P = g["size"]
    n = 3
    home = np.array([capi.hash_code(int(k)) % P for k in nodes["kmer"].tolist()], dtype=np.int64)
    owner = home // ((P + n - 1) // n)
    dumps = []
    for r in range(n):
        sel = owner == r
        dumps.append({k: v[sel] for k, v in nodes.items()})
    # shard 0 hands its last 5 nodes over as a tail blob (raw build nodes), and 2 of them also show up adopted in shard 1
    take = 5
    d0 = dumps[0]
    moved = {k: v[-take:] for k, v in d0.items()}
    dumps[0] = {k: v[:-take] for k, v in d0.items()}
    dumps[1] = {k: np.concatenate([dumps[1][k], moved[k][:2]]) for k in dumps[1]}
    blob = bytearray(np.array([0, take, 32, 0], dtype=np.uint64).tobytes())
    for i in range(take):
        cnt = np.zeros(8, dtype=np.float16)
        for s, w in enumerate((int(moved["l"][i]), int(moved["r"][i]))):
            for b in range(4):
                c = (w >> (24 - 8 * b)) & 0xFF
                cnt[4 * s + b] = 300.0 if c == 255 else float(c)          # a saturated lane holds "255 or more"
        blob += np.array([moved["kmer"][i], ~moved["ord"][i]], dtype=np.uint64).tobytes() + cnt.tobytes()
    extra = blob_margin_nodes(bytes(blob))
    for k in ("kmer", "l", "r", "ord"):
        assert np.array_equal(extra[k], moved[k]), k
    arr, nul, plan = merge_dumps_host(dumps + [extra], g["init_slots"], g["load"], False, polyA[0], polyA[1])
    assert plan["final_size"] == P and plan["count"] == g["count"]
    slot = np.nonzero(np.unpackbits(nul)[:P])[0].astype(np.uint64)
    sel = slot.astype(np.int64)
    assert np.array_equal(slot, g["slot"]) and np.array_equal(arr["kmer"][sel], g["kmer"])
    assert np.array_equal(arr["l_link"][sel], g["l"]) and np.array_equal(arr["r_link"][sel], g["r"])
