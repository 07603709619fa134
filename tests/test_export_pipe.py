"""The pipelined KmerSet export (csrc/export_pipe.cu, export_expand.cpp): occupied nodes only over the link, host threads
expand them into the reference's P-slot table image (kmerSet.h:88-99).  Bar: the bytes handed to the consumer are the
bytes of the plain copy -- array[P] AND nul_flag -- whatever mix of compact / plain chunks the pipe chose.

CPU: the host expander against numpy.  GPU: pipe vs plain copy of the same finalized table."""
import numpy as np
import pytest

from conftest import random_reads, reads_to_arrays


@pytest.fixture(scope="module")
def capi():
    from dbg_assembly_b200 import capi as m
    m.load()
    return m


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("n_slots,density,misalign", [(0, 0.5, 0), (5, 1.0, 0), (8, 0.5, 0), (1000, 0.5, 0), (4099, 0.1, 16), (65536, 0.5, 0),
                                                     (70001, 0.9, 32), (70001, 0.0, 0), (12345, 1.0, 48)])
def test_host_expand_matches_numpy(capi, wide, n_slots, density, misalign):
    rng = np.random.default_rng(n_slots * 7 + int(density * 10) + int(wide))
    occ = rng.random(n_slots) < density
    bits = np.packbits(occ.astype(np.uint8))                 # MSB first, like nul_flag
    bits = np.concatenate([bits, np.zeros(1, np.uint8)])
    q = 4 if wide else 2
    n = int(occ.sum())
    nodes = rng.integers(1, 1 << 63, size=(n + 2, q), dtype=np.uint64)       # padded: the expander may read one node past the end
    raw = np.full(n_slots * q * 8 + 128, 0xAB, dtype=np.uint8)               # poisoned destination, chosen alignment
    base = (-raw.ctypes.data) % 64 + misalign
    dst = raw[base: base + n_slots * q * 8].view(np.uint64).reshape(n_slots, q)
    used = capi.host_expand_nodes(bits, n_slots, nodes, dst, wide)
    assert used == n
    want = np.zeros((n_slots, q), dtype=np.uint64)
    want[occ] = nodes[:n]
    assert np.array_equal(dst, want)
    assert (raw[:base] == 0xAB).all() and (raw[base + n_slots * q * 8:] == 0xAB).all()      # nothing written outside


def _build(dbg, K, slots, n_reads, seed, force_wide=False):
    reads = random_reads(seed, n_reads, 60, 120, genome_len=40000, err=0.02)
    bases, offs = reads_to_arrays(reads)
    b = dbg.DBGBuilder(K=K, max_read_len=120, init_slots=slots, track_order=True, force_wide=force_wide)
    b.submit(bases, offs)
    st = b.finalize()
    return b, st


@pytest.mark.gpu
@pytest.mark.parametrize("K,force_wide", [(31, False), (25, True), (47, False)])
def test_pipe_export_equals_plain_copy(monkeypatch, K, force_wide):
    import dbg_assembly_b200 as dbg
    if dbg.capi.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    b, st = _build(dbg, K, 400_000, 6000, seed=K, force_wide=force_wide)
    try:
        P = st["array_size"]
        monkeypatch.setenv("DBG_B200_EXPORT", "plain")
        ref_arr, ref_nul = b.export_kmerset()
        assert b.export_info()["chunks_compact"] == 0
        monkeypatch.setenv("DBG_B200_EXPORT", "pipe")
        monkeypatch.setenv("DBG_B200_EXPORT_CHUNK", "8192")        # 49 chunks: the pipe takes tables of >= 4 chunks
        nb = ref_arr.dtype.itemsize
        pinned_a = dbg.capi.PinnedBuffer(P * nb)
        pinned_n = dbg.capi.PinnedBuffer(P // 8 + 1)
        try:
            for threads, slots, pinned, no_direct in [(1, 2, False, False), (4, 8, False, False), (3, 1, True, False), (4, 8, True, False),
                                                      (2, 2, True, True), (0, 4, False, False)]:
                monkeypatch.setenv("DBG_B200_EXPORT_THREADS", str(threads))
                monkeypatch.setenv("DBG_B200_EXPORT_SLOTS", str(slots))
                if no_direct:
                    monkeypatch.setenv("DBG_B200_EXPORT_NO_DIRECT", "1")
                else:
                    monkeypatch.delenv("DBG_B200_EXPORT_NO_DIRECT", raising=False)
                if pinned:
                    arr = pinned_a.array.view(ref_arr.dtype)
                    nul = pinned_n.array
                else:
                    arr = np.empty(P, dtype=ref_arr.dtype)
                    nul = np.empty(P // 8 + 1, dtype=np.uint8)
                arr.view(np.uint8)[:] = 0xCD
                nul[:] = 0xCD
                b.export_kmerset(arr, nul)
                info = b.export_info()
                assert np.array_equal(nul, ref_nul), (threads, slots, pinned)
                assert arr.tobytes() == ref_arr.tobytes(), (threads, slots, pinned)
                n_chunks = (P + 8191) // 8192
                if threads == 0:
                    assert info["chunks_compact"] == 0 and info["chunks_plain"] == 1          # the pipe stepped aside
                else:
                    assert info["chunks_compact"] + info["chunks_plain"] == n_chunks
                    assert info["nodes"] == st["count"]
                    if not pinned or no_direct:
                        assert info["chunks_plain"] == 0       # pageable destination: every chunk goes through the ring
                        assert info["link_bytes"] == st["count"] * nb + P // 8 + 1
        finally:
            pinned_a.close(); pinned_n.close()
    finally:
        b.close()


@pytest.mark.gpu
def test_pipe_export_default_settings_on_a_table_of_several_chunks(monkeypatch):
    """DBG_B200_EXPORT=pipe with its default chunk size (2^20 slots) and threads: a 6 M-slot table, pinned destination"""
    import dbg_assembly_b200 as dbg
    from dbg_assembly_b200 import synth
    p = synth.make_params(seed=5, genome_len=400_000, read_len=150, insert=400, err=0.01, n_rate=0.001)
    bases, offs = synth.reads_host(p, 0, 60_000)
    with dbg.DBGBuilder(K=31, max_read_len=150, init_slots=6_000_000, track_order=True) as b:
        b.submit(bases, offs)
        st = b.finalize()
        P = st["array_size"]
        monkeypatch.setenv("DBG_B200_EXPORT", "plain")
        ref_arr, ref_nul = b.export_kmerset()
        monkeypatch.setenv("DBG_B200_EXPORT", "pipe")
        pa = dbg.capi.PinnedBuffer(P * 16); pn = dbg.capi.PinnedBuffer(P // 8 + 1)
        try:
            arr = pa.array.view(ref_arr.dtype); nul = pn.array
            for _ in range(2):           # second call reuses the ring and the workers
                arr.view(np.uint8)[:] = 0x5A; nul[:] = 0x5A
                b.export_kmerset(arr, nul)
                assert np.array_equal(nul, ref_nul) and arr.tobytes() == ref_arr.tobytes()
            info = b.export_info()
            assert info["chunks_compact"] >= 1 and info["chunks_compact"] + info["chunks_plain"] == (P + (1 << 20) - 1) >> 20
        finally:
            pa.close(); pn.close()
