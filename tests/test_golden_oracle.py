"""CPU: the C restatement against the committed golden vectors (generated from the real reference by
tests/golden/make_golden.py). These run everywhere, including on the GPU box where /root/reference is absent."""
import os

import numpy as np
import pytest

import compare
import golden_util


def test_kmer_iterator_fixture(oracle):
    z = np.load(os.path.join(golden_util.GOLDEN, "kmeriter_fixture.npz"))
    off = z["off"]
    for i, (s, k) in enumerate(zip(z["seqs"].tolist(), z["ks"].tolist())):
        km, pos = oracle.kmer_windows(s.encode(), int(k))
        assert np.array_equal(km, z["kmers"][off[i]:off[i + 1]])
        assert np.array_equal(pos, z["pos"][off[i]:off[i + 1]])


def test_kmer_file_fixture(oracle):
    z = np.load(os.path.join(golden_util.GOLDEN, "kmers_fixture.npz"))
    vals, k = oracle.load_kmers(os.path.join(golden_util.GOLDEN, "kmers_fixture.txt"))
    assert k == int(z["k"]) == 19 and np.array_equal(vals, z["kmers"])


@pytest.mark.parametrize("tag,files", [("fq", ["records_a.fq"]), ("fq_fq", ["records_a.fq", "records_c.fq"]), ("fa", ["records_b.fa"])])
def test_record_stream_fixture(oracle, tag, files):
    metas, recs = golden_util.parse_records(os.path.join(golden_util.GOLDEN, f"records_{tag}.expected.txt"))
    rc, d = oracle.load_reads([os.path.join(golden_util.GOLDEN, f) for f in files])
    assert rc == 0 and d["n_reads"] == len(recs)
    for i, (rid, h, s, q) in enumerate(recs):
        assert rid == i + 1
        assert d["hdr"][int(d["hdr_off"][i]):int(d["hdr_off"][i + 1])].decode() == h
        assert d["seq"][int(d["seq_off"][i]):int(d["seq_off"][i + 1])].decode() == s
        assert d["qual"][int(d["qual_off"][i]):int(d["qual_off"][i + 1])].decode() == q
    agg = [m for m in metas if m[0] == "#AGG"][0]
    assert [int(v) for v in agg[2:]] == [d["a_records"], d["a_total"], d["a_avg"], d["a_max"], d["a_min"]]


@pytest.mark.parametrize("name", golden_util.CASES)
def test_hot_path_golden(oracle, name):
    c = golden_util.load_case(name)
    ref = c["ref"]
    res = oracle.run(c["bases"], c["seq_off"], c["k"], c["kmers"], fraction=c["fraction"], min_size=c["min_size"])
    compare.check_hits(ref, res["row_off"], res["hit_kid"], res["hit_pos"], c["kmers"])
    compare.check_index(ref, res["inv_off"], res["inv_read"], c["kmers"])
    sx, sy, ss = res["conn"]
    assert np.array_equal(sx, ref["conn_x"]) and np.array_equal(sy, ref["conn_y"]) and np.array_equal(ss, ref["conn_score"])
    assert res["cut_n"] == ref["cut_n"] and res["cut_score"] == ref["cut_score"]
    co, cm, to, tx, ty = res["comp"]
    assert compare.components_partition(co, cm) == compare.components_partition(ref["comp_off"], ref["comp_read"])
    assert compare.tree_edges(to, tx, ty) == compare.tree_edges(ref["tree_off"], ref["tree_x"], ref["tree_y"])
    assert len(compare.components_partition(co, cm)) == ref["scaffold_components"]


@pytest.mark.parametrize("name", golden_util.ENRICH_CASES)
def test_merge_enrichment_golden(oracle, name):
    """merge_components + enrichment + final merge against the dump of the real reference"""
    import oracle_lib
    c = golden_util.load_case(name)
    ref = c["ref"]
    res = oracle.run(c["bases"], c["seq_off"], c["k"], c["kmers"], fraction=c["fraction"], min_size=c["min_size"])
    assert res["cut_n"] == ref["cut_n"] and res["cut_score"] == ref["cut_score"]
    e = oracle_lib.enrich(oracle, res, len(c["kmers"]), min_size=c["min_size"], enrich_min=c["enrich"])
    compare.check_enrichment(ref, e, c["kmers"])
    assert len(e["final_id"]) == ref["final_components"] and ref["cores"] >= 2
