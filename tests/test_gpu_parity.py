"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the oracle on seeded inputs.
Bit-exact for every stage: hit lists, inverted index, pair scores, cut (n, s*), selected edges, components."""
import numpy as np
import pytest

import compare
import datagen

pytestmark = pytest.mark.gpu


def _pack(seqs):
    seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    bases = b"".join(seqs)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=off[1:])
    return bases, off


def _selected_set(sx, sy, ss, n):
    """undirected edges inside the first n entries of the canonical directed list"""
    e = set()
    for a, b, s in zip(sx[:n].tolist(), sy[:n].tolist(), ss[:n].tolist()):
        e.add((min(a, b), max(a, b), s))
    return e


def _run_and_compare(oracle, bases, off, k, kmers, fraction=0.15, min_size=30, min_score=1, check_sorted_hits=True):
    import hga_b200
    kmers = np.unique(np.asarray(kmers, dtype=np.uint64))
    ref = oracle.run(bases, off, k, kmers, fraction=fraction, min_size=min_size, min_score=min_score)
    with hga_b200.Handle(kmers, k) as h:
        h.scan(bases, off)
        row_off, kid, pos = h.get_hits()
        assert np.array_equal(row_off, ref["row_off"]), "row offsets differ"
        assert np.array_equal(kid, ref["hit_kid"]), "hit k-mer ids differ"
        assert np.array_equal(pos, ref["hit_pos"]), "hit positions differ"
        if check_sorted_hits:
            r2, kid2, pos2 = h.get_hits(sorted_by_kmer_id=True)
            ro = ref["row_off"].astype(np.int64)
            rows = np.repeat(np.arange(len(ro) - 1), np.diff(ro))
            order = np.lexsort((ref["hit_pos"], ref["hit_kid"], rows))
            assert np.array_equal(kid2, ref["hit_kid"][order]) and np.array_equal(pos2, ref["hit_pos"][order])
        h.build_index()
        inv_off, inv_read = h.get_index()
        assert np.array_equal(inv_off, ref["inv_off"]) and np.array_equal(inv_read, ref["inv_read"]), "inverted index differs"
        h.pair_count(min_score=min_score)
        x, y, s, inc = h.get_pairs()
        ux, uy, us = compare.undirected(*ref["conn"])
        assert np.array_equal(x, ux) and np.array_equal(y, uy) and np.array_equal(s.astype(np.uint64), us), "pair scores differ"
        lens = np.diff(ref["inv_off"].astype(np.int64))
        assert inc == int((lens * (lens - 1) // 2).sum())
        h.select_edges(fraction=fraction)
        sel = h.get_selection()
        assert sel["n_directed"] == ref["cut_n"] and sel["cut_score"] == ref["cut_score"], "cut differs"
        got = set(zip(sel["x"].tolist(), sel["y"].tolist(), sel["score"].tolist()))
        assert got == _selected_set(*ref["conn"], ref["cut_n"]), "selected edge set differs"
        h.components(min_size=min_size)
        comp = h.get_components()
        co, cm, _, _, _ = ref["comp"]
        want = compare.components_partition(co, cm)
        label = comp["label"]
        gotc = sorted(tuple(sorted(int(v) for v in (np.nonzero(label == r)[0] + comp["read_id_first"]))) for r in comp["comp_label"])
        assert gotc == want, "components differ"
        assert comp["comp_size"].tolist() == [len(c) for c in sorted(want, key=lambda c: c[0])]
        return ref, h.metrics()


def test_kat2_multiplicity(oracle):
    rng = np.random.default_rng(7)
    X = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8)); Y = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8))
    reads = [X + Y + X, X, Y + X + X + X, Y, "ACGT"]
    bases, off = _pack(reads)
    kmers = [oracle.kmer_windows(X[i:i + 15].encode(), 15)[0][0] for i in range(5)]
    ref, _ = _run_and_compare(oracle, bases, off, 15, kmers, fraction=1.0, min_size=1)
    und = compare.undirected(*ref["conn"])
    assert list(zip(und[0].tolist(), und[1].tolist(), und[2].tolist())) == [(1, 2, 10), (1, 3, 30), (2, 3, 15)]


@pytest.mark.parametrize("k", [1, 2, 5, 15, 16, 17, 19, 21, 27, 31, 32])
def test_scan_k_sweep_with_exceptions(oracle, k):
    """every k, reads of awkward lengths (0, < k, == k, tile-crossing), non-ACGT bytes and lowercase"""
    import hga_b200
    rng = np.random.default_rng(100 + k)
    g = datagen.random_genome(60000, 5 + k)
    reads = []
    for L in [0, 1, k - 1, k, k + 1, 0, 0, 33, 4096, 4095, 4097, 9000, 150, 150, 12000, 2, 0]:
        L = max(L, 0)
        s0 = int(rng.integers(0, 60000 - L + 1))
        reads.append(bytearray(datagen.to_ascii(g[s0:s0 + L]).encode()))
    for r in reads:
        for _ in range(len(r) // 300):
            r[int(rng.integers(0, len(r)))] = b"NnacgtRY*\r"[int(rng.integers(0, 10))]
    reads.append(bytearray(datagen.to_ascii(g[100:400]).lower().encode()))
    reads += [bytearray(datagen.to_ascii(r).encode()) for r in datagen.sample_reads(g, 200, 100, k)]
    reads.append(bytearray())
    bases, off = _pack([bytes(r) for r in reads])
    allk = np.unique(datagen.canonical_kmers(g, k))
    kmers = allk[::3] if allk.shape[0] > 10 else allk
    kmers = np.unique(np.concatenate([kmers, np.array([0], dtype=np.uint64)]))   # poly-A / exception-made k-mers are members
    ref_ro, ref_kid, ref_pos = oracle.scan(bases, off, k, kmers)
    with hga_b200.Handle(kmers, k) as h:
        h.scan(bases, off)
        row_off, kid, pos = h.get_hits()
    assert np.array_equal(row_off, ref_ro)
    assert np.array_equal(kid, ref_kid)
    assert np.array_equal(pos, ref_pos)
    assert ref_kid.shape[0] > 0


def test_config1_like(oracle):
    a = datagen.random_genome(20000, 1000)
    b = datagen.mutate(a, 0.03, 1001)
    reads = datagen.sample_reads(a, 4000, 150, 1, error_rate=0.005) + datagen.sample_reads(b, 4000, 150, 2, error_rate=0.005)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], 19)
    ref, m = _run_and_compare(oracle, bases, off, 19, kmers)
    assert ref["cut_n"] > 10000


@pytest.mark.parametrize("k", [15, 17, 19, 21])
def test_long_reads(oracle, k):
    a = datagen.random_genome(120000, 2000 + k)
    b = datagen.mutate(a, 0.02, 2001 + k)
    reads = datagen.sample_reads(a, 300, 7800, 3, error_rate=0.08, length_sigma=0.6, max_len=60000) + \
        datagen.sample_reads(b, 300, 7800, 4, error_rate=0.08, length_sigma=0.6, max_len=60000)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], k)
    _run_and_compare(oracle, bases, off, k, kmers, min_size=5, check_sorted_hits=(k == 19))


def test_config5_like_tetraploid(oracle):
    """BASELINE config 5 in small: four haplotypes, dense discriminative set (k-mers absent from at least one haplotype);
    all stages plus the merge + enrichment stage"""
    import hga_b200
    import oracle_lib
    from test_gpu_golden import _gpu_enrichment
    base = datagen.random_genome(60000, 7000)
    haps = [base] + [datagen.mutate(base, 0.01, 7001 + i) for i in range(3)]
    reads = []
    for i, h in enumerate(haps):
        reads += datagen.sample_reads(h, 240, 5000, 7010 + i, error_rate=0.03, length_sigma=0.5, max_len=40000)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers(haps, 19, mode="not_all")
    ref, _ = _run_and_compare(oracle, bases, off, 19, kmers, min_size=5, check_sorted_hits=False)
    want = oracle_lib.enrich(oracle, ref, len(kmers), min_size=5, enrich_min=20)
    with hga_b200.Handle(kmers, 19) as h:
        h.scan(bases, off); h.build_index(); h.pair_count(min_score=1); h.select_edges(fraction=0.15)
        got = _gpu_enrichment(h, 5, 20)
    assert np.array_equal(got["core_id"], want["core_id"]) and np.array_equal(got["final_id"], want["final_id"])
    assert np.array_equal(got["purged_off"], want["purged_off"]) and np.array_equal(got["purged_read"], want["purged_read"])
    for g, w in zip(got["econn"], want["econn"]):
        assert np.array_equal(g, w)
    for g, w in zip(got["final_reads"], want["final_reads"]):
        assert np.array_equal(g, w)
    assert len(want["core_id"]) >= 1


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_all_stages(oracle, seed):
    """the fuzz cases of tests/test_oracle_vs_ref.py (where the oracle is checked against the real reference) through the C-ABI"""
    import hga_b200
    import oracle_lib
    from test_gpu_golden import _gpu_enrichment
    haps, reads, k, fraction, min_size, enrich_min = datagen.fuzz_case(seed)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers(haps, k)
    ref, _ = _run_and_compare(oracle, bases, off, k, kmers, fraction=fraction, min_size=min_size, check_sorted_hits=False)
    want = oracle_lib.enrich(oracle, ref, len(kmers), min_size=min_size, enrich_min=enrich_min)
    with hga_b200.Handle(kmers, k) as h:
        h.scan(bases, off); h.build_index(); h.pair_count(min_score=1); h.select_edges(fraction=fraction)
        got = _gpu_enrichment(h, min_size, enrich_min)
    assert np.array_equal(got["core_id"], want["core_id"]) and np.array_equal(got["final_id"], want["final_id"])
    assert np.array_equal(got["purged_off"], want["purged_off"]) and np.array_equal(got["purged_read"], want["purged_read"])
    for g, w in zip(got["econn"], want["econn"]):
        assert np.array_equal(g, w)
    assert len(got["final_reads"]) == len(want["final_reads"])
    for g, w in zip(got["final_reads"], want["final_reads"]):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("max_size", [40, 120])
def test_sc_max_size_with_enrichment(oracle, max_size):
    """--sc_max_size (hga_enrich_ex): the size-limited sequential union_find replayed on the host over all selected edges"""
    import hga_b200
    import oracle_lib
    from test_gpu_golden import _gpu_enrichment
    a = datagen.random_genome(30000, 71); b = datagen.mutate(a, 0.03, 72)
    reads = datagen.sample_reads(a, 6000, 150, 73, error_rate=0.005) + datagen.sample_reads(b, 6000, 150, 74, error_rate=0.005)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], 19)
    res = oracle.run(bases, off, 19, kmers, min_size=10, max_size=max_size)
    want = oracle_lib.enrich(oracle, res, len(kmers), min_size=10, enrich_min=20)
    with hga_b200.Handle(kmers, 19) as h:
        h.scan(bases, off); h.build_index(); h.pair_count(min_score=1); h.select_edges(fraction=0.15)
        h.enrich(min_size=10, enrichment_min_score=20, max_size=max_size)
        e = h.get_enrichment()
        po, pr = h.get_purged_index()
    assert len(want["core_id"]) >= 2 and max(len(r) for r in want["core_reads"]) <= max_size
    assert np.array_equal(e["core_id"], want["core_id"]) and np.array_equal(e["core_read"], np.concatenate(want["core_reads"]))
    assert np.array_equal(po, want["purged_off"]) and np.array_equal(pr, want["purged_read"])
    assert np.array_equal(e["conn_x"], want["econn"][0]) and np.array_equal(e["conn_y"], want["econn"][1]) and np.array_equal(e["conn_score"].astype(np.uint64), want["econn"][2])
    assert np.array_equal(e["final_id"], want["final_id"]) and np.array_equal(e["final_read"], np.concatenate(want["final_reads"]))


def test_sc_score_mode_with_enrichment(oracle):
    """--sc_score S: pivot subset + score > S selection, then merge + enrichment on top of it"""
    import hga_b200
    import oracle_lib
    from test_gpu_golden import _gpu_enrichment
    a = datagen.random_genome(30000, 61); b = datagen.mutate(a, 0.03, 62)
    reads = datagen.sample_reads(a, 6000, 150, 63, error_rate=0.005) + datagen.sample_reads(b, 6000, 150, 64, error_rate=0.005)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], 19)
    S = 60
    res = oracle.run(bases, off, 19, kmers, min_size=10, sc_score=S)
    want = oracle_lib.enrich(oracle, res, len(kmers), min_size=10, enrich_min=20)
    with hga_b200.Handle(kmers, 19) as h:
        h.scan(bases, off); h.build_index()
        ro, _, _ = h.get_hits()
        pivots = (np.nonzero(np.diff(ro.astype(np.int64)) >= S)[0] + 1).astype(np.uint32)
        h.pair_count(min_score=S, pivots=pivots)
        h.select_edges(score_threshold=S)
        sel = h.get_selection()
        assert sel["n_directed"] == res["cut_n"]
        got = _gpu_enrichment(h, 10, 20)
    assert len(want["core_id"]) >= 2
    assert np.array_equal(got["core_id"], want["core_id"]) and np.array_equal(got["final_id"], want["final_id"])
    assert np.array_equal(got["purged_off"], want["purged_off"]) and np.array_equal(got["purged_read"], want["purged_read"])
    for g, w in zip(got["econn"], want["econn"]):
        assert np.array_equal(g, w)
    for g, w in zip(got["final_reads"], want["final_reads"]):
        assert np.array_equal(g, w)


def test_heavy_rows_overflow_shared_accumulator(oracle):
    """one repeated segment present in > 3072 reads: partner sets overflow the shared-memory table"""
    rng = np.random.default_rng(5)
    rep = datagen.to_ascii(rng.integers(0, 4, 60, dtype=np.uint8))
    reads = []
    for i in range(3600):
        u = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8))
        reads.append(u + rep if i % 2 else rep + u)
    bases, off = _pack(reads)
    kmers = np.unique(datagen.canonical_kmers(np.array(["ACGT".index(c) for c in rep], dtype=np.uint8), 21))
    ref, m = _run_and_compare(oracle, bases, off, 21, kmers, min_size=2, check_sorted_hits=False)
    assert m["heavy_pivots"] > 0


def test_pivot_subset_and_min_score(oracle):
    import hga_b200
    a = datagen.random_genome(30000, 31)
    b = datagen.mutate(a, 0.03, 32)
    reads = datagen.sample_reads(a, 250, 1200, 5) + datagen.sample_reads(b, 250, 1200, 6)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], 19)
    row_off, kid, pos = oracle.scan(bases, off, 19, kmers)
    inv_off, inv_read = oracle.index(row_off, kid, len(kmers))
    pivots = np.array([3, 10, 11, 57, 200, 201, 499, 500], dtype=np.uint32)
    for min_score in (1, 20):
        cx, cy, cs = oracle.connections(row_off, kid, inv_off, inv_read, min_score=min_score, pivots=pivots)
        want = {}
        for x, y, s in zip(cx.tolist(), cy.tolist(), cs.tolist()):
            want[(min(x, y), max(x, y))] = s
        with hga_b200.Handle(kmers, 19) as h:
            h.scan(bases, off)
            h.build_index()
            h.pair_count(min_score=min_score, pivots=pivots)
            x, y, s, _ = h.get_pairs()
        got = {(a_, b_): c_ for a_, b_, c_ in zip(x.tolist(), y.tolist(), s.tolist())}
        assert got == want
    # the EMPTY subset (get_connections({}) in the reference: no connections) is not "all reads"
    with hga_b200.Handle(kmers, 19) as h:
        h.scan(bases, off)
        h.build_index()
        h.pair_count(min_score=1, pivots=np.zeros(0, dtype=np.uint32))
        x, y, s, _ = h.get_pairs()
        assert x.shape[0] == 0 and y.shape[0] == 0
        h.pair_count(min_score=1)
        assert h.get_pairs()[0].shape[0] > 0


def test_score_threshold_mode(oracle):
    import hga_b200
    a = datagen.random_genome(30000, 41)
    b = datagen.mutate(a, 0.03, 42)
    reads = datagen.sample_reads(a, 250, 1200, 7) + datagen.sample_reads(b, 250, 1200, 8)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], 19)
    ref = oracle.run(bases, off, 19, kmers)
    ux, uy, us = compare.undirected(*ref["conn"])
    S = int(np.median(us))
    with hga_b200.Handle(kmers, 19) as h:
        h.scan(bases, off); h.build_index(); h.pair_count(min_score=1)
        h.select_edges(score_threshold=S)
        sel = h.get_selection()
    m = us > S
    assert set(zip(sel["x"].tolist(), sel["y"].tolist(), sel["score"].tolist())) == set(zip(ux[m].tolist(), uy[m].tolist(), us[m].tolist()))
    assert sel["n_directed"] == 2 * int(m.sum())


def test_empty_inputs(oracle):
    import hga_b200
    kmers = np.array([5, 9, 77], dtype=np.uint64)
    with hga_b200.Handle(kmers, 5) as h:
        for seqs in ([], [b""], [b"", b"", b""], [b"ACG"]):
            bases, off = _pack(seqs)
            h.scan(bases, off)
            row_off, kid, pos = h.get_hits()
            assert kid.shape[0] == 0 and np.all(row_off == 0) and row_off.shape[0] == len(seqs) + 1
            h.build_index(); h.pair_count(); h.select_edges(); h.components()
            assert h.get_pairs()[0].shape[0] == 0 and h.get_components()["comp_label"].shape[0] == 0
    with hga_b200.Handle(np.zeros(0, dtype=np.uint64), 19) as h:
        bases, off = _pack([b"ACGTACGTACGTACGTACGTACGTACGT"])
        h.scan(bases, off)
        assert h.get_hits()[1].shape[0] == 0


def test_errors():
    import hga_b200
    with pytest.raises(hga_b200.HgaError):
        hga_b200.Handle(np.array([1], dtype=np.uint64), 33)      # "Kmer size is too big"
    with pytest.raises(hga_b200.HgaError):
        hga_b200.Handle(np.array([1, 1], dtype=np.uint64), 5)    # duplicates
    with hga_b200.Handle(np.array([1], dtype=np.uint64), 5) as h:
        with pytest.raises(hga_b200.HgaError):
            h.build_index()                                       # stage out of order


def test_host_scan_pipelined_chunks(oracle, monkeypatch):
    """hga_scan with host buffers copies the bases in chunks on a second stream and scans every chunk as it lands; with
    HGA_SCAN_CHUNK_MB=1 a 3 MB input takes that path (tiles at chunk seams need the previous chunk's last bases)"""
    import hga_b200
    monkeypatch.setenv("HGA_SCAN_CHUNK_MB", "1")
    a = datagen.random_genome(150000, 77)
    b = datagen.mutate(a, 0.02, 78)
    reads = datagen.sample_reads(a, 210, 7800, 5, error_rate=0.05, length_sigma=0.5, max_len=60000) + \
        datagen.sample_reads(b, 210, 7800, 6, error_rate=0.05, length_sigma=0.5, max_len=60000)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    assert len(bases) > 2 * (1 << 20) + 4096
    kmers = datagen.discriminative_kmers([a, b], 19)
    ref_ro, ref_kid, ref_pos = oracle.scan(bases, off, 19, kmers)
    with hga_b200.Handle(kmers, 19) as h:
        h.scan(bases, off)
        row_off, kid, pos = h.get_hits()
        m = h.metrics()
    assert np.array_equal(row_off, ref_ro) and np.array_equal(kid, ref_kid) and np.array_equal(pos, ref_pos)
    assert m["h2d_ms"] > 0


@pytest.mark.parametrize("case", ["short", "long", "long_low_threshold", "no_cores"])
def test_merge_and_enrichment(oracle, case):
    """SURVEY §8f-1 (hga_enrich) against the C restatement of merge_components / get_connections / restricted union_find"""
    import hga_b200
    import oracle_lib
    from test_gpu_golden import _gpu_enrichment
    if case == "short":
        a = datagen.random_genome(30000, 41); b = datagen.mutate(a, 0.03, 42); k = 19; min_size, enrich_min = 30, 20
        reads = datagen.sample_reads(a, 6000, 150, 43, error_rate=0.005) + datagen.sample_reads(b, 6000, 150, 44, error_rate=0.005)
    elif case == "no_cores":
        a = datagen.random_genome(5000, 45); b = datagen.mutate(a, 0.03, 46); k = 19; min_size, enrich_min = 30, 20
        reads = datagen.sample_reads(a, 60, 150, 47) + datagen.sample_reads(b, 60, 150, 48)
    else:
        a = datagen.random_genome(60000, 51); b = datagen.mutate(a, 0.02, 52); k = 15
        min_size, enrich_min = (5, 20) if case == "long" else (3, 2)
        reads = datagen.sample_reads(a, 350, 2000, 53, error_rate=0.05, length_sigma=0.5) + datagen.sample_reads(b, 350, 2000, 54, error_rate=0.05, length_sigma=0.5)
    bases, off = _pack([datagen.to_ascii(r) for r in reads])
    kmers = datagen.discriminative_kmers([a, b], k)
    res = oracle.run(bases, off, k, kmers, min_size=min_size)
    want = oracle_lib.enrich(oracle, res, len(kmers), min_size=min_size, enrich_min=enrich_min)
    with hga_b200.Handle(kmers, k) as h:
        h.scan(bases, off)
        h.build_index()
        h.pair_count(min_score=1)
        h.select_edges(fraction=0.15)
        got = _gpu_enrichment(h, min_size, enrich_min)
    assert np.array_equal(got["core_id"], want["core_id"])
    assert len(got["core_kmers"]) == len(want["core_kmers"])
    for g, w in zip(got["core_kmers"], want["core_kmers"]):
        assert np.array_equal(np.sort(g), w)
    for g, w in zip(got["core_reads"], want["core_reads"]):
        assert np.array_equal(g, w)
    assert np.array_equal(got["purged_off"], want["purged_off"]) and np.array_equal(got["purged_read"], want["purged_read"])
    for g, w in zip(got["econn"], want["econn"]):
        assert np.array_equal(g, w)
    assert np.array_equal(got["final_id"], want["final_id"])
    assert len(got["final_reads"]) == len(want["final_reads"])
    for g, w in zip(got["final_reads"], want["final_reads"]):
        assert np.array_equal(g, w)
    if case == "no_cores":
        assert len(want["core_id"]) == 0 and len(got["final_id"]) == 0
    else:
        assert len(want["core_id"]) >= 2 and len(want["econn"][0]) > 0
