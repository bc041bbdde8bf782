"""GPU: size-independent properties of a run far larger than the oracle can check in seconds (~100 Mbases of long reads):
CSR consistency, sortedness, the pair-score checksum computed independently from the inverted index, selection counts,
and components that are closed under the selected edges."""
import numpy as np
import pytest

import datagen

pytestmark = pytest.mark.gpu


def test_large_run_invariants():
    import hga_b200
    k = 19
    a = datagen.random_genome(2_000_000, 901)
    b = datagen.mutate(a, 0.01, 902)
    reads = datagen.sample_reads(a, 5000, 10000, 11, error_rate=0.05, length_sigma=0.5, min_len=100, max_len=60000) + \
        datagen.sample_reads(b, 5000, 10000, 12, error_rate=0.05, length_sigma=0.5, min_len=100, max_len=60000)
    seqs = [datagen.to_ascii(r).encode() for r in reads]
    bases = b"".join(seqs)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    kmers = datagen.discriminative_kmers([a, b], k)
    n_reads = len(seqs)
    with hga_b200.Handle(kmers, k) as h:
        h.scan(bases, off)
        row_off, kid, pos = h.get_hits()
        h.build_index()
        inv_off, inv_read = h.get_index()
        h.pair_count(min_score=1)
        x, y, s, inc = h.get_pairs()
        h.select_edges(fraction=0.15)
        sel = h.get_selection()
        h.components(min_size=30)
        comp = h.get_components()
        m = h.metrics()
    E = kid.shape[0]
    assert E > 1_000_000 and m["n_bases"] == len(bases)
    # hits: CSR, ids in range, positions inside the read, ascending inside a row
    ro = row_off.astype(np.int64)
    assert ro[0] == 0 and ro[-1] == E and np.all(np.diff(ro) >= 0)
    assert kid.max() < len(kmers)
    rows = np.repeat(np.arange(n_reads), np.diff(ro))
    lens = np.diff(off.astype(np.int64))
    assert np.all(pos >= k) and np.all(pos <= lens[rows])
    same_row = rows[1:] == rows[:-1]
    assert np.all(pos[1:][same_row] > pos[:-1][same_row])
    # a sample of hits re-derived from the bases: the window ending at pos is the k-mer the id names
    rng = np.random.default_rng(5)
    for i in rng.integers(0, E, 200):
        r = rows[i]
        w = seqs[r][pos[i] - k:pos[i]]
        codes = np.array([b"ACGT".index(c) for c in w], dtype=np.uint8)
        assert datagen.canonical_kmers(codes, k)[0] == kmers[kid[i]]
    # inverted index: same multiset as the hits, lists ascending
    io = inv_off.astype(np.int64)
    assert io[-1] == E and inv_read.shape[0] == E
    assert np.array_equal(np.bincount(kid, minlength=len(kmers)), np.diff(io))
    list_id = np.repeat(np.arange(len(kmers)), np.diff(io))
    same_list = list_id[1:] == list_id[:-1]
    assert np.all(inv_read[1:][same_list] >= inv_read[:-1][same_list])
    # pairs: strictly ascending (x, y), x < y, and the checksum sum(score) = sum_k (len_k^2 - sum_x mult_x(k)^2) / 2
    assert np.all(x < y)
    key = x.astype(np.uint64) << np.uint64(32) | y.astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    run_start = np.ones(E, dtype=bool)
    run_start[1:] = ~(same_list & (inv_read[1:] == inv_read[:-1]))
    run_len = np.diff(np.append(np.nonzero(run_start)[0], E)).astype(np.int64)
    run_list = list_id[run_start]
    sq = np.bincount(run_list, weights=(run_len * run_len).astype(np.float64), minlength=len(kmers))
    L = np.diff(io).astype(np.float64)
    assert int(s.astype(np.int64).sum()) == int(round(((L * L - sq) / 2).sum()))
    assert inc == int((np.diff(io) * (np.diff(io) - 1) // 2).sum())
    # selection: n = (size_t)(2P * 0.15); everything above the cut is in, nothing below
    P = x.shape[0]
    assert sel["n_directed"] == int(2 * P * 0.15)
    cut = sel["cut_score"]
    above = int((s > cut).sum())
    assert int((sel["score"] > cut).sum()) == above and np.all(sel["score"] >= cut)
    assert sel["x"].shape[0] == above + (sel["n_directed"] - 2 * above + 1) // 2
    # components: closed under the selected edges; label = smallest member; listed sizes match
    label = comp["label"].astype(np.int64)
    first = comp["read_id_first"]
    assert np.array_equal(label[sel["x"] - first], label[sel["y"] - first])
    touched = np.zeros(n_reads, dtype=bool)
    touched[sel["x"] - first] = True; touched[sel["y"] - first] = True
    assert np.all(label[touched] <= np.nonzero(touched)[0] + first)
    sizes = np.bincount(label[touched] - first, minlength=n_reads)
    for lab, sz in zip(comp["comp_label"], comp["comp_size"]):
        assert sizes[lab - first] == sz and sz >= 30
    assert int((sizes >= 30).sum()) == comp["comp_label"].shape[0]
