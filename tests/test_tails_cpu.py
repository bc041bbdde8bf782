"""CPU: the product's host-side tail connections (hga_host_tail_connections, SURVEY §8f-2 second piece) against the reference's own
get_core_component_connections (spanning-tree tails, approximate overlaps, tail amplification, tail k-mer unions) run by
oracle/_ref/ref_driver --full, and against a committed golden fixture generated from it. The inputs (hits, scaffold components with
their spanning trees, purged index) come from the C oracle, which the other tests pin against the reference stage by stage."""
import os

import numpy as np
import pytest

import datagen
import golden_util
import refdump


def _inputs(oracle, seq, seq_off, k, kmers, min_size):
    import oracle_lib
    res = oracle.run(seq, seq_off, k, kmers, min_size=min_size)
    n = len(seq_off) - 1
    length = np.diff(np.asarray(seq_off).astype(np.int64)).astype(np.uint32)
    ro = res["row_off"].astype(np.int64)
    rows = np.repeat(np.arange(n), np.diff(ro))
    o = np.lexsort((res["hit_pos"], res["hit_kid"], rows))          # rows sorted by (kmer_id, pos): hga_get_hits(h, 1, ..)
    co, cm, to, tx, ty = res["comp"]
    eng = oracle_lib.Engine(oracle, res["row_off"], res["hit_kid"], len(kmers), res["inv_off"], res["inv_read"])
    try:
        eng.merge(co, cm)                                          # state after merge_components(scaffold_components)
        po, pr = eng.index()
    finally:
        eng.close()
    avg = int(length.astype(np.int64).sum() // n)                  # meta.avg_read_length (integer division, SequenceRecordIterator.cpp:64)
    return dict(row_off=res["row_off"], kmer_id=res["hit_kid"][o], pos=res["hit_pos"][o], read_len=length, avg_read_length=avg, comp_off=co, comp_member=cm,
                tree_off=to, tree_x=tx, tree_y=ty, purged_off=po, purged_read=pr)


CASES = {
    "long_k15": dict(genome_size=60000, divergence=0.02, k=15, read_len=2000, coverage=12, seed=21, error_rate=0.05, length_sigma=0.5, _min_size=5),
    "long_k19": dict(genome_size=120000, divergence=0.02, k=19, read_len=3000, coverage=15, seed=22, error_rate=0.03, length_sigma=0.5, _min_size=5),
    "short_ties": dict(genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7, error_rate=0.005, fmt="fastq", _min_size=30),
    "mid": dict(genome_size=100000, divergence=0.03, k=19, read_len=400, coverage=25, seed=10, error_rate=0.01, _min_size=30),
    # the survivors' merged k-mer lists make some tree "overlaps" longer than the reads: the uint64 distances wrap for some start
    # vertices of the first sweep (the tails then come out EMPTY) and not for others; both sides start at the smallest vertex id
    "wrapping_k17": dict(genome_size=80000, divergence=0.02, k=17, read_len=2500, coverage=12, seed=61, error_rate=0.04, length_sigma=0.5, _min_size=5),
    "k21": dict(genome_size=80000, divergence=0.03, k=21, read_len=1500, coverage=14, seed=62, error_rate=0.02, length_sigma=0.4, _min_size=5),
}


@pytest.mark.parametrize("name", list(CASES))
def test_tail_connections_match_the_reference(oracle, ref_driver, tmp_path, name):
    import hga_b200
    kw = dict(CASES[name])
    min_size = kw.pop("_min_size")
    paths, kp = datagen.make_diploid_case(str(tmp_path), **kw)
    ref = refdump.run_ref(ref_driver, paths, kp, enrich=20, full=True, min_size=min_size)
    assert ref["scaffold_components"] > 2 and ref["tail_connections"] > 0
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    a = _inputs(oracle, reads["seq"], reads["seq_off"], k, kmers, min_size)
    x, y, s = hga_b200.capi.host_tail_connections(amplification_min_score=40, **a)
    assert np.array_equal(x, ref["tconn_x"]) and np.array_equal(y, ref["tconn_y"]) and np.array_equal(s, ref["tconn_score"])
    # ... and chained into the spectral stage: the clusters the reference merges
    got = hga_b200.capi.spectral_clustering(x[s > 5], y[s > 5], s[s > 5], 16)
    so = ref["spectral_off"].astype(np.int64)
    want = [(ref["spectral_member"][so[i]:so[i + 1]].tolist(), int(ref["spectral_first"][i])) for i in range(len(so) - 1)]
    assert sorted((sorted(c.tolist()), int(c[0])) for c in got if len(c)) == want


def test_tail_connections_golden(oracle):
    import hga_b200
    z = np.load(os.path.join(golden_util.GOLDEN, "tails_a.npz"))
    a = _inputs(oracle, z["bases"].tobytes(), z["seq_off"], int(z["k"]), z["kmers"], int(z["min_size"]))
    x, y, s = hga_b200.capi.host_tail_connections(amplification_min_score=int(z["amplification_min_score"]), **a)
    assert np.array_equal(x, z["tconn_x"]) and np.array_equal(y, z["tconn_y"]) and np.array_equal(s, z["tconn_score"])


@pytest.mark.parametrize("name", ["long_k15", "short_ties"])
def test_whole_tail_spectral_block_reaches_the_reference_final_components(oracle, ref_driver, tmp_path, name):
    """run_clustering :764-794 INCLUDING the tail / spectral block: scaffold merge -> hga_host_tail_connections -> hga_spectral_clustering
    -> merge of the clusters (here on the oracle's engine state; the product's GPU form is hga_enrich_full, tests/test_zy_gpu_tail_block.py,
    and its rule tests/test_second_merge_rule.py) -> enrichment:
    final components and their ids equal to the reference's (ref_driver --full)"""
    import hga_b200
    import oracle_lib
    kw = dict(CASES[name])
    min_size = kw.pop("_min_size")
    paths, kp = datagen.make_diploid_case(str(tmp_path), **kw)
    ref = refdump.run_ref(ref_driver, paths, kp, enrich=20, full=True, min_size=min_size)
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    a = _inputs(oracle, reads["seq"], reads["seq_off"], k, kmers, min_size)
    x, y, s = hga_b200.capi.host_tail_connections(amplification_min_score=40, **a)
    clusters = [c for c in hga_b200.capi.spectral_clustering(x[s > 5], y[s > 5], s[s > 5], 16) if len(c)]
    res = oracle.run(reads["seq"], reads["seq_off"], k, kmers, min_size=min_size)
    eng = oracle_lib.Engine(oracle, res["row_off"], res["hit_kid"], len(kmers), res["inv_off"], res["inv_read"])
    try:
        eng.merge(res["comp"][0], res["comp"][1])                                    # :764
        if clusters:                                                                 # :774
            off = np.cumsum([0] + [len(c) for c in clusters]).astype(np.uint64)
            eng.merge(off, np.concatenate(clusters))
        cores = eng.ids(min_size)
        ex, ey, es = oracle.canonical_sort(*eng.connections(cores, 20))             # :787
        eo, em, _, _, _ = oracle.union_find(ex, ey, min_size=2, max_size=-1, restricted=cores)
        eng.merge(eo, em)
        final = eng.ids(min_size)
        got = sorted((int(np.sort(eng.component_reads(c))[0]), int(c), np.sort(eng.component_reads(c)).tolist()) for c in final)
    finally:
        eng.close()
    fo = ref["final_off"].astype(np.int64)
    want = [(int(ref["final_read"][fo[i]]), int(ref["final_id"][i]), ref["final_read"][fo[i]:fo[i + 1]].tolist()) for i in range(len(fo) - 1)]
    assert len(cores) == ref["cores"]
    assert got == want
