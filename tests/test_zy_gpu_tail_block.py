"""GPU: hga_enrich_full = everything run_clustering does after the scaffold union_find INCLUDING the tail / spectral block
(ReadClusteringEngine.cpp:764-794), through the C-ABI, against the dumps of the real reference (ref_driver --enrich 20
--full, fixtures tests/golden/full_*.npz): tail connections, spectral clusters, the state after the merge of the clusters (cores with
the reference's survivor ids, merged k-mer lists, twice purged index), enrichment connections, final components. The file sorts near the end
on purpose: it is the newest path (written after round 1's GPU minutes were spent; the same source passes on the host,
tests/test_kernel_bodies_on_host.py)."""
import os
import subprocess

import numpy as np
import pytest

import compare
import datagen
import golden_util

pytestmark = pytest.mark.gpu


def _run_full(h, c):
    h.scan(c["bases"], c["seq_off"])
    h.build_index()
    h.pair_count(min_score=1)
    h.select_edges(fraction=c["fraction"])
    h.enrich_full(c["seq_off"], min_size=c["min_size"], enrichment_min_score=c["enrich"], tail_amplification_min_score=40, spectral_dims=16)
    e = h.get_enrichment()
    ko, kk = h.get_core_kmers()
    po, pr = h.get_purged_index()
    co = e["core_off"].astype(np.int64); fo = e["final_off"].astype(np.int64); ko = ko.astype(np.int64)
    return dict(core_id=e["core_id"], core_kmers=[kk[ko[i]:ko[i + 1]] for i in range(len(ko) - 1)],
                core_reads=[e["core_read"][co[i]:co[i + 1]] for i in range(len(co) - 1)], purged_off=po, purged_read=pr,
                econn=(e["conn_x"], e["conn_y"], e["conn_score"].astype(np.uint64)), final_id=e["final_id"],
                final_reads=[e["final_read"][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)], assignment=e["assignment"], read_id_first=e["read_id_first"])


@pytest.mark.parametrize("name", golden_util.FULL_CASES)
def test_cuda_full_run_clustering_vs_reference_golden(name):
    import hga_b200
    c = golden_util.load_case(name)
    ref = c["ref"]
    with hga_b200.Handle(c["kmers"], c["k"]) as h:
        e = _run_full(h, c)
        t = h.get_tail_block()
        assert t["ran"] and t["n_scaffold_cores"] == ref["merged_scaffolds"]
        assert np.array_equal(t["conn_x"], ref["tconn_x"]) and np.array_equal(t["conn_y"], ref["tconn_y"]) and np.array_equal(t["conn_score"], ref["tconn_score"])
        so = ref["spectral_off"].astype(np.int64)
        want = [(ref["spectral_member"][so[i]:so[i + 1]].tolist(), int(ref["spectral_first"][i])) for i in range(len(so) - 1)]
        assert sorted((sorted(cl.tolist()), int(cl[0])) for cl in t["clusters"]) == want
        compare.check_enrichment(ref, e, c["kmers"])
        a = e["assignment"]
        for fid, reads in zip(e["final_id"], e["final_reads"]):
            assert np.all(a[reads - e["read_id_first"]] == fid)
        m = h.metrics()
        assert m["n_cores"] == ref["cores"] and m["n_final_components"] == ref["final_components"]
        # the same handle again without the block: the scaffold cores, more of them
        h.enrich(min_size=c["min_size"], enrichment_min_score=c["enrich"])
        assert len(h.get_enrichment()["core_id"]) == ref["merged_scaffolds"] and not h.get_tail_block()["ran"]


def test_full_equals_plain_when_the_block_does_not_run():
    """at most two scaffold components (:768): hga_enrich_full is hga_enrich_ex"""
    import hga_b200
    c = golden_util.load_case("full_long")
    big = 200                                     # only the largest components survive this min_size
    with hga_b200.Handle(c["kmers"], c["k"]) as h:
        h.scan(c["bases"], c["seq_off"])
        h.build_index()
        h.pair_count(min_score=1)
        h.select_edges(fraction=c["fraction"])
        h.enrich(min_size=big, enrichment_min_score=c["enrich"])
        a = h.get_enrichment()
        if len(a["core_id"]) > 2:
            pytest.skip("more than two components of that size")
        h.enrich_full(c["seq_off"], min_size=big, enrichment_min_score=c["enrich"])
        b = h.get_enrichment()
        assert not h.get_tail_block()["ran"]
        for key in ("core_id", "core_read", "conn_x", "conn_y", "conn_score", "final_id", "final_read", "assignment"):
            assert np.array_equal(a[key], b[key]), key


def test_engine_mirror_run_clustering_with_tail_block():
    import hga_b200
    c = golden_util.load_case("full_short")
    ref = c["ref"]

    class _Reader(hga_b200.SequenceRecords):
        def __init__(self):
            self.bases, self.seq_off = c["bases"], c["seq_off"]
            n = len(c["seq_off"]) - 1
            self.headers = [b"r%d" % i for i in range(n)]
            self.qualities = [b""] * n

    eng = hga_b200.ReadClusteringEngine(_Reader(), hga_b200.ReadClusteringConfig(scaffold_component_min_size=c["min_size"],
                                                                                 enrichment_connections_min_score=c["enrich"]))
    ids = eng.run_clustering(c["kmers"], c["k"])
    assert ids == [int(v) for v in ref["final_id"]]
    fo = ref["final_off"].astype(np.int64)
    for i, fid in enumerate(ids):
        assert np.array_equal(eng.final_components[fid], ref["final_read"][fo[i]:fo[i + 1]])
    eng.close()


def test_cli_tail_block_exports_the_reference_final_components(tmp_path):
    """categorization, default run (tail / spectral block included): files named after the reference's surviving component ids, holding the reference's reads"""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hybrid-genome-assembler_b200", "categorization")
    c = golden_util.load_case("full_long")
    ref = c["ref"]
    so = np.asarray(c["seq_off"]).astype(np.int64)
    rp, kp, outdir = str(tmp_path / "reads.fa"), str(tmp_path / "kmers.txt"), str(tmp_path / "out")
    with open(rp, "wb") as f:
        for i in range(len(so) - 1):
            f.write(b">r%d\n" % (i + 1) + c["bases"][so[i]:so[i + 1]] + b"\n")
    with open(kp, "w") as f:
        for v in c["kmers"]:
            f.write(datagen.kmer_to_str(v, c["k"]) + "\n")
    r = subprocess.run([exe, rp, "--kmers", kp, "-o", outdir, "--sc_min_size", str(c["min_size"])], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    fo = ref["final_off"].astype(np.int64)
    want = {f"#{int(fid)}.fa": [f"r{int(v)}" for v in ref["final_read"][fo[i]:fo[i + 1]]] for i, fid in enumerate(ref["final_id"])}
    assert sorted(os.listdir(outdir)) == sorted(want)
    for name, hdrs in want.items():
        lines = open(os.path.join(outdir, name)).read().split("\n")
        assert [l[1:] for l in lines[0::2] if l] == hdrs
    assert f"Exported {len(want)} components" in r.stdout


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_full_run_clustering_vs_oracle(oracle, tmp_path, seed):
    """random cases through hga_enrich_full against the C oracle's engine (scaffold merge, merge of the clusters, enrichment, final
    merge), with the tail connections and clusters the block itself reports - the chain scripts/fuzz_tail_block.py checks against
    the real reference on CPU"""
    import hga_b200
    import oracle_lib
    rng = np.random.default_rng(1000 + seed)
    long_ = seed != 2
    kw = dict(genome_size=int(rng.integers(30000, 80000)), divergence=float(rng.choice([0.02, 0.03])), k=int(rng.choice([15, 17, 19, 21])),
              read_len=int(rng.integers(1000, 3000)) if long_ else int(rng.integers(150, 400)), coverage=int(rng.integers(10, 14)) if long_ else int(rng.integers(20, 28)),
              seed=int(rng.integers(1, 10000)), error_rate=float(rng.choice([0.005, 0.02, 0.05])))
    if long_:
        kw["length_sigma"] = 0.5
    ms = 5 if long_ else 30
    paths, kp = datagen.make_diploid_case(str(tmp_path), **kw)
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    with hga_b200.Handle(kmers, k) as h:
        e = _run_full(h, dict(bases=reads["seq"], seq_off=reads["seq_off"], fraction=0.15, min_size=ms, enrich=20))
        t = h.get_tail_block()
    res = oracle.run(reads["seq"], reads["seq_off"], k, kmers, min_size=ms)
    eng = oracle_lib.Engine(oracle, res["row_off"], res["hit_kid"], len(kmers), res["inv_off"], res["inv_read"])
    try:
        ids = eng.merge(res["comp"][0], res["comp"][1])
        assert t["ran"] == (len(ids) > 2) and t["n_scaffold_cores"] == len(ids)
        clusters = [c for c in t["clusters"] if len(c)]
        if clusters:
            eng.merge(np.cumsum([0] + [len(c) for c in clusters]).astype(np.uint64), np.concatenate(clusters))
        cores = np.sort(eng.ids(ms))
        assert np.array_equal(e["core_id"], cores)
        for c, got_k, got_r in zip(cores, e["core_kmers"], e["core_reads"]):
            assert np.array_equal(np.sort(got_k), np.sort(eng.component_kmers(c))) and np.array_equal(got_r, np.sort(eng.component_reads(c)))
        po, pr = eng.index()
        assert np.array_equal(e["purged_off"], po) and np.array_equal(e["purged_read"], pr)
        ex, ey, es = oracle.canonical_sort(*eng.connections(cores, 20))
        assert np.array_equal(e["econn"][0], ex) and np.array_equal(e["econn"][1], ey) and np.array_equal(e["econn"][2], es)
        eo, em, _, _, _ = oracle.union_find(ex, ey, min_size=2, max_size=-1, restricted=cores)
        eng.merge(eo, em)
        want = sorted((int(np.sort(eng.component_reads(c))[0]), int(c), np.sort(eng.component_reads(c)).tolist()) for c in eng.ids(ms))
    finally:
        eng.close()
    assert [(int(r[0]), int(f), r.tolist()) for f, r in zip(e["final_id"], e["final_reads"])] == want
