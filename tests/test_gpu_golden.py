"""GPU: the CUDA path through the C-ABI against the committed golden vectors of the real reference."""
import numpy as np
import pytest

import compare
import golden_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_util.CASES)
def test_cuda_vs_reference_golden(name):
    import hga_b200
    c = golden_util.load_case(name)
    ref = c["ref"]
    with hga_b200.Handle(c["kmers"], c["k"]) as h:
        h.scan(c["bases"], c["seq_off"])
        row_off, kid, pos = h.get_hits()
        compare.check_hits(ref, row_off, kid, pos, c["kmers"])
        h.build_index()
        inv_off, inv_read = h.get_index()
        compare.check_index(ref, inv_off, inv_read, c["kmers"])
        h.pair_count(min_score=1)
        x, y, s, _ = h.get_pairs()
        ux, uy, us = compare.undirected(ref["conn_x"], ref["conn_y"], ref["conn_score"])
        assert np.array_equal(x, ux) and np.array_equal(y, uy) and np.array_equal(s.astype(np.uint64), us)
        compare.check_directed_symmetric(ref["conn_x"], ref["conn_y"], ref["conn_score"])
        h.select_edges(fraction=c["fraction"])
        sel = h.get_selection()
        assert sel["n_directed"] == ref["cut_n"] and sel["cut_score"] == ref["cut_score"]
        n = ref["cut_n"]
        want = {(min(a, b), max(a, b), sc) for a, b, sc in zip(ref["conn_x"][:n].tolist(), ref["conn_y"][:n].tolist(), ref["conn_score"][:n].tolist())}
        assert set(zip(sel["x"].tolist(), sel["y"].tolist(), sel["score"].tolist())) == want
        h.components(min_size=c["min_size"])
        comp = h.get_components()
        got = sorted(tuple(sorted(int(v) for v in (np.nonzero(comp["label"] == r)[0] + comp["read_id_first"]))) for r in comp["comp_label"])
        assert got == compare.components_partition(ref["comp_off"], ref["comp_read"])


def test_engine_mirror_on_golden():
    """the reference-shaped Python interface (construct_indices / get_all_connections / scaffold_components)"""
    import hga_b200
    c = golden_util.load_case("config1_mini")
    ref = c["ref"]

    class _Reader:
        bases, seq_off = c["bases"], c["seq_off"]
        n_reads = len(c["seq_off"]) - 1

    eng = hga_b200.ReadClusteringEngine(_Reader(), hga_b200.ReadClusteringConfig())
    eng.construct_indices(c["kmers"], c["k"])
    x, y, s = eng.get_all_connections(1)
    assert np.array_equal(x, ref["conn_x"]) and np.array_equal(y, ref["conn_y"]) and np.array_equal(s, ref["conn_score"])
    comps = eng.scaffold_components()
    assert sorted(tuple(cc.tolist()) for cc in comps) == compare.components_partition(ref["comp_off"], ref["comp_read"])
    ids = eng.component_ids()[:7]
    gx, gy, gs = eng.get_connections(ids, 20)
    m = np.isin(ref["conn_x"], ids) & (ref["conn_score"] >= 20)
    assert np.array_equal(gx, ref["conn_x"][m]) and np.array_equal(gy, ref["conn_y"][m]) and np.array_equal(gs, ref["conn_score"][m])
    eng.close()


def _gpu_enrichment(h, min_size, enrich_min):
    """hga_enrich result in the shape compare.check_enrichment expects"""
    h.enrich(min_size=min_size, enrichment_min_score=enrich_min)
    e = h.get_enrichment()
    ko, kk = h.get_core_kmers()
    po, pr = h.get_purged_index()
    co = e["core_off"].astype(np.int64); fo = e["final_off"].astype(np.int64); ko = ko.astype(np.int64)
    return dict(core_id=e["core_id"], core_kmers=[kk[ko[i]:ko[i + 1]] for i in range(len(ko) - 1)],
                core_reads=[e["core_read"][co[i]:co[i + 1]] for i in range(len(co) - 1)], purged_off=po, purged_read=pr,
                econn=(e["conn_x"], e["conn_y"], e["conn_score"].astype(np.uint64)), final_id=e["final_id"],
                final_reads=[e["final_read"][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)], assignment=e["assignment"], read_id_first=e["read_id_first"])


@pytest.mark.parametrize("name", golden_util.ENRICH_CASES)
def test_cuda_merge_enrichment_vs_reference_golden(name):
    """SURVEY §8f-1 through the C-ABI: cores (reference survivor ids), merged k-mer lists, purged index, enrichment connections,
    final components, against the dump of the real reference"""
    import hga_b200
    c = golden_util.load_case(name)
    ref = c["ref"]
    with hga_b200.Handle(c["kmers"], c["k"]) as h:
        h.scan(c["bases"], c["seq_off"])
        h.build_index()
        h.pair_count(min_score=1)
        h.select_edges(fraction=c["fraction"])
        sel = h.get_selection()
        assert sel["n_directed"] == ref["cut_n"] and sel["cut_score"] == ref["cut_score"]
        e = _gpu_enrichment(h, c["min_size"], c["enrich"])
        compare.check_enrichment(ref, e, c["kmers"])
        # assignment = final membership
        a = e["assignment"]
        for fid, reads in zip(e["final_id"], e["final_reads"]):
            assert np.all(a[reads - e["read_id_first"]] == fid)
        assert int((a != 0).sum()) == sum(len(r) for r in e["final_reads"])
        m = h.metrics()
        assert m["n_cores"] == ref["cores"] and m["n_final_components"] == ref["final_components"]


def test_engine_mirror_run_clustering(tmp_path):
    """reference-shaped run_clustering + export_components against the golden final components"""
    import hga_b200
    c = golden_util.load_case("enrich_short")
    ref = c["ref"]

    class _Reader(hga_b200.SequenceRecords):
        def __init__(self):
            self.bases, self.seq_off = c["bases"], c["seq_off"]
            n = len(c["seq_off"]) - 1
            self.headers = [b"r%d" % i for i in range(n)]
            self.qualities = [b""] * n

    eng = hga_b200.ReadClusteringEngine(_Reader(), hga_b200.ReadClusteringConfig(scaffold_component_min_size=c["min_size"],
                                                                                 enrichment_connections_min_score=c["enrich"]))
    ids = eng.run_clustering(c["kmers"], c["k"], tail_block=False)          # the fixture is a ref_driver --enrich dump (no tail / spectral block)
    assert ids == [int(v) for v in ref["final_id"]]
    fo = ref["final_off"].astype(np.int64)
    for i, fid in enumerate(ids):
        assert np.array_equal(eng.final_components[fid], ref["final_read"][fo[i]:fo[i + 1]])
    out = str(tmp_path / "clusters")
    eng.export_components(ids, out)
    import os
    assert sorted(os.listdir(out)) == sorted(f"#{i}.fa" for i in ids)
    first = open(os.path.join(out, f"#{ids[0]}.fa"), "rb").read().split(b"\n")
    assert first[0] == b">r%d" % (int(ref["final_read"][0]) - 1)
    eng.close()
