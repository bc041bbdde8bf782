"""Pins the C restatement (oracle/hga_oracle.c) against the REAL reference code (oracle/_ref/ref_driver =
unmodified /root/reference sources + shim). CPU only. Skipped when the driver binary is unavailable."""
import os

import numpy as np
import pytest

import compare
import datagen
import refdump


def _load_case(oracle, paths, kp):
    rc, reads = oracle.load_reads(paths)
    assert rc == 0
    kmers, k = oracle.load_kmers(kp)
    return reads, kmers, k


def test_kat1_kmer_iterator(oracle, ref_driver):
    # SURVEY §8c KAT 1 (N -> (0,0) rule)
    km, pos = oracle.kmer_windows(b"ACGTNACGTTTGACCAGTA", 3)
    assert list(zip(pos[:12].tolist(), km[:12].tolist())) == [(3, 6), (4, 6), (5, 1), (6, 48), (7, 1), (8, 6), (9, 6), (10, 1), (11, 0),
                                                              (12, 16), (13, 52), (14, 33)]
    rk, rp = refdump.ref_kmeriter(ref_driver, "ACGTNACGTTTGACCAGTA", 3)
    assert np.array_equal(km, rk) and np.array_equal(pos, rp)
    assert oracle.kmer_windows(b"AC", 3)[0].shape[0] == 0


@pytest.mark.parametrize("k", [1, 2, 15, 17, 19, 21, 31, 32])
def test_kmer_iterator_random(oracle, ref_driver, k):
    rng = np.random.default_rng(k)
    alphabet = np.frombuffer(b"ACGTACGTACGTNacgt\rX", dtype=np.uint8)
    for L in (0, 1, k - 1, k, k + 1, 200):
        if L < 0:
            continue
        s = alphabet[rng.integers(0, alphabet.shape[0], size=L)].tobytes()
        if b"\r" in s or L == 0:
            s = s.replace(b"\r", b"n")  # argv cannot carry these to the driver
        km, pos = oracle.kmer_windows(s, k)
        if L == 0:
            assert km.shape[0] == 0
            continue
        rk, rp = refdump.ref_kmeriter(ref_driver, s.decode(), k)
        assert np.array_equal(km, rk) and np.array_equal(pos, rp)


def test_load_kmers_matches_ref(oracle, ref_driver, tmp_path):
    p = str(tmp_path / "k.txt")
    with open(p, "w") as f:
        f.write("ACGTACGTACGTACGTACG\nCGTACGTACGTACGTACGT\nACGTACGTACGTACGTACG\nTTTTTTTTTTTTTTTTTTT\nACGTNNNNACGTACGTACG\nGATTACAGATTACAGATTA")
    a, k = oracle.load_kmers(p)
    b, kr = refdump.ref_canon(ref_driver, p)
    assert k == kr == 19 and np.array_equal(a, b)


def _full_compare(oracle, ref_driver, paths, kp, fraction=0.15, min_size=30, threads=1, enrich=0, sc_score=0, max_size=-1):
    reads, kmers, k = _load_case(oracle, paths, kp)
    ref = refdump.run_ref(ref_driver, paths, kp, fraction=fraction, min_size=min_size, threads=threads, enrich=enrich, sc_score=sc_score, max_size=max_size)
    res = oracle.run(reads["seq"], reads["seq_off"], k, kmers, fraction=fraction, min_size=min_size, sc_score=sc_score, max_size=max_size)
    assert ref["k"] == k and ref["n_kmers"] == kmers.shape[0] and ref["n_reads"] == reads["n_reads"]
    compare.check_hits(ref, res["row_off"], res["hit_kid"], res["hit_pos"], kmers)
    compare.check_index(ref, res["inv_off"], res["inv_read"], kmers)
    sx, sy, ss = res["conn"]
    assert np.array_equal(sx, ref["conn_x"]) and np.array_equal(sy, ref["conn_y"]) and np.array_equal(ss, ref["conn_score"])
    assert res["cut_n"] == ref["cut_n"] and res["cut_score"] == ref["cut_score"]
    co, cm, to, tx, ty = res["comp"]
    assert compare.components_partition(co, cm) == compare.components_partition(ref["comp_off"], ref["comp_read"])
    assert compare.tree_edges(to, tx, ty) == compare.tree_edges(ref["tree_off"], ref["tree_x"], ref["tree_y"])
    # root = element [0] of each component (ReadClusteringEngine.cpp:366)
    roots = sorted(int(cm[int(o)]) for o in co[:-1])
    assert roots == sorted(int(r) for r in ref["comp_root"])
    if enrich:
        import oracle_lib
        compare.check_enrichment(ref, oracle_lib.enrich(oracle, res, kmers.shape[0], min_size=min_size, enrich_min=enrich), kmers)
    return ref, res


def test_kat2_multiplicity(oracle, ref_driver, tmp_path):
    # SURVEY §8c KAT 2: r1=XYX r2=X r3=YXXX r4=Y r5=ACGT, SDKs = first five 15-mers of X
    rng = np.random.default_rng(7)
    X = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8)); Y = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8))
    reads = [X + Y + X, X, Y + X + X + X, Y, "ACGT"]
    rp = str(tmp_path / "r.fa"); kp = str(tmp_path / "k.txt")
    datagen.write_fasta(rp, reads)
    with open(kp, "w") as f:
        for i in range(5):
            f.write(X[i:i + 15] + "\n")
    ref, res = _full_compare(oracle, ref_driver, [rp], kp, fraction=1.0, min_size=1)
    und = compare.undirected(*res["conn"])
    assert list(zip(und[0].tolist(), und[1].tolist(), und[2].tolist())) == [(1, 2, 10), (1, 3, 30), (2, 3, 15)]
    io = res["inv_off"]
    for i in range(5):
        assert res["inv_read"][int(io[i]):int(io[i + 1])].tolist() == [1, 1, 2, 3, 3, 3]


@pytest.mark.parametrize("fmt,threads", [("fasta", 1), ("fastq", 1), ("fastq", 4)])
def test_config1_like_small(oracle, ref_driver, tmp_path, fmt, threads):
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=6000, divergence=0.03, k=19, read_len=150, coverage=20, seed=3,
                                          error_rate=0.005, fmt=fmt)
    ref, res = _full_compare(oracle, ref_driver, paths, kp, threads=threads)
    assert ref["directed_connections"] > 1000 and ref["scaffold_components"] >= 1


@pytest.mark.parametrize("k", [15, 17, 21])
def test_long_reads_k_sweep(oracle, ref_driver, tmp_path, k):
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=30000, divergence=0.02, k=k, read_len=2000, coverage=12, seed=k,
                                          error_rate=0.05, length_sigma=0.5)
    _full_compare(oracle, ref_driver, paths, kp, min_size=3)


@pytest.mark.parametrize("threads", [1, 4])
def test_merge_and_enrichment_short_reads(oracle, ref_driver, tmp_path, threads):
    # merge_components (unique union, index purge with its truncation quirk), get_connections over the merged
    # cores, restricted union-find, second merge: all stages against the real reference
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7,
                                          error_rate=0.005, fmt="fastq")
    ref, _ = _full_compare(oracle, ref_driver, paths, kp, threads=threads, enrich=20)
    assert ref["cores"] > 10 and ref["enrichment_connections"] > 100 and ref["final_components"] == ref["cores"]


@pytest.mark.parametrize("enrich,min_size", [(20, 5), (3, 3)])
def test_merge_and_enrichment_long_reads(oracle, ref_driver, tmp_path, enrich, min_size):
    # long noisy reads: k-mers occur several times per read, so the survivor keeps stale entries (quirk i)
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=30000, divergence=0.02, k=15, read_len=2000, coverage=12, seed=21,
                                          error_rate=0.05, length_sigma=0.5)
    ref, _ = _full_compare(oracle, ref_driver, paths, kp, min_size=min_size, enrich=enrich)
    assert ref["cores"] >= 1


def test_sc_score_mode_with_enrichment(oracle, ref_driver, tmp_path):
    # --sc_score S: pivot subset, score > S (a connection between a pivot and a non-pivot would exist in ONE direction only, which
    # decides whose root survives a size tie in union_find, :459-466; with min score = pivot threshold that needs repeated k-mers)
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7,
                                          error_rate=0.005, fmt="fastq")
    ref, res = _full_compare(oracle, ref_driver, paths, kp, min_size=10, enrich=20, sc_score=60)
    assert ref["cut_n"] > 100 and ref["cores"] >= 2


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_all_stages_against_the_reference(oracle, ref_driver, tmp_path, seed):
    # random k, genome, read length, coverage, error, repeats, fraction, component sizes, enrichment threshold; every stage from
    # the hit lists to the final components against the real reference
    haps, reads, k, fraction, min_size, enrich_min = datagen.fuzz_case(seed)
    rp = str(tmp_path / "r.fa"); kp = str(tmp_path / "k.txt")
    datagen.write_fasta(rp, reads)
    datagen.write_kmers(kp, np.random.default_rng(seed).permutation(datagen.discriminative_kmers(haps, k)), k)
    _full_compare(oracle, ref_driver, [rp], kp, fraction=fraction, min_size=min_size, enrich=enrich_min)


def test_reference_runs_its_whole_pipeline_in_the_driver(oracle, ref_driver, tmp_path):
    # SURVEY §8f-2 oracle (round-2 groundwork): ref_driver --full runs the reference's own tails, tail amplification, tail
    # connections, spectral clustering (lib/clustering compiled against the Eigen2 stand-in), merges and enrichment. No product
    # counterpart yet; this pins that the real-code oracle exists and is self-consistent.
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7,
                                          error_rate=0.005, fmt="fastq")
    ref = refdump.run_ref(ref_driver, paths, kp, enrich=20, full=True)
    assert ref["scaffold_components"] > 2 and ref["tail_connections"] > 0 and ref["spectral_clusters"] >= 1
    so = ref["spectral_off"].astype(np.int64)
    members = ref["spectral_member"]
    assert so[-1] == members.shape[0] and len(set(members.tolist())) == members.shape[0]            # clusters are disjoint
    scaffold_roots = set(int(r) for r in ref["comp_root"])
    assert set(members.tolist()) <= scaffold_roots                                                # ... and made of scaffold components
    assert set(ref["tconn_x"].tolist()) | set(ref["tconn_y"].tolist()) <= scaffold_roots
    merged_away = int(sum(max(0, int(so[i + 1] - so[i]) - 1) for i in range(len(so) - 1)))
    assert ref["cores"] == ref["scaffold_components"] - merged_away
    fo = ref["final_off"].astype(np.int64)
    assert ref["final_components"] == len(fo) - 1 == ref["cores"]
    assert len(set(ref["final_read"].tolist())) == ref["final_read"].shape[0]                      # a read is in at most one final component


@pytest.mark.parametrize("max_size", [40, 120])
def test_sc_max_size_with_enrichment(oracle, ref_driver, tmp_path, max_size):
    # --sc_max_size: unions that would exceed the limit are skipped (:457); the components then depend on the edge order
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7,
                                          error_rate=0.005, fmt="fastq")
    ref, res = _full_compare(oracle, ref_driver, paths, kp, min_size=10, enrich=20, max_size=max_size)
    co = res["comp"][0].astype(np.int64)
    assert ref["cores"] >= 2 and int(np.diff(co).max()) <= max_size


def test_config5_like_tetraploid(oracle, ref_driver, tmp_path):
    # BASELINE config 5 in small: four haplotype read files, dense discriminative set (k-mers absent from at least one haplotype)
    paths, kp = datagen.make_polyploid_case(str(tmp_path), genome_size=12000, divergence=0.02, k=19, read_len=1500, coverage=10, seed=31,
                                            error_rate=0.02, length_sigma=0.4)
    ref, _ = _full_compare(oracle, ref_driver, paths, kp, min_size=4, enrich=20)
    assert ref["n_reads"] == 4 * 80 and ref["scaffold_components"] >= 1


def test_non_acgt_and_crlf(oracle, ref_driver, tmp_path):
    a = datagen.random_genome(3000, 11)
    reads = [datagen.to_ascii(r) for r in datagen.sample_reads(a, 120, 100, 12)]
    rng = np.random.default_rng(13)
    mangled = []
    for i, r in enumerate(reads):
        r = list(r)
        for j in rng.integers(0, len(r), size=3):
            r[j] = "NnacgtRY*"[int(rng.integers(0, 9))]
        mangled.append("".join(r).lower() if i % 17 == 0 else "".join(r))
    rp = str(tmp_path / "r.fa"); kp = str(tmp_path / "k.txt")
    datagen.write_fasta(rp, mangled, newline="\r\n")
    sdk = np.unique(datagen.canonical_kmers(a, 15))[::3]
    datagen.write_kmers(kp, sdk, 15)
    _full_compare(oracle, ref_driver, [rp], kp, min_size=2)


def test_record_stream_matches_ref(oracle, ref_driver, tmp_path):
    a = datagen.random_genome(2000, 5)
    r1 = datagen.sample_reads(a, 7, 60, 1); r2 = datagen.sample_reads(a, 5, 80, 2)
    p1 = str(tmp_path / "a.fq"); p2 = str(tmp_path / "b.fq"); p3 = str(tmp_path / "c.fa")
    datagen.write_fastq(p1, r1, prefix="x"); datagen.write_fastq(p2, r2, prefix="y")
    with open(p2, "a") as f:
        f.write("\n")  # ONE trailing blank line is tolerated (SURVEY §8a parity item 4)
    datagen.write_fasta(p3, r2, prefix="z")
    for paths in ([p1], [p1, p2], [p3], [p3, p3]):
        rc, metas, recs = refdump.ref_records(ref_driver, paths)
        orc_rc, d = oracle.load_reads(paths)
        assert rc == 0 and orc_rc == 0
        assert d["n_reads"] == len(recs)
        for i, (rid, h, s, q) in enumerate(recs):
            assert rid == i + 1
            assert d["hdr"][int(d["hdr_off"][i]):int(d["hdr_off"][i + 1])].decode() == h
            assert d["seq"][int(d["seq_off"][i]):int(d["seq_off"][i + 1])].decode() == s
            assert d["qual"][int(d["qual_off"][i]):int(d["qual_off"][i + 1])].decode() == q
        agg = [m for m in metas if m[0] == "#AGG"][0]
        assert [int(v) for v in agg[2:]] == [d["a_records"], d["a_total"], d["a_avg"], d["a_max"], d["a_min"]]
        assert d["a_max"] == 0  # never updated in the reference (SequenceRecordIterator.cpp:59-62)
        per = [m for m in metas if m[0] == "#META"]
        for i, m in enumerate(per):
            assert [int(v) for v in m[2:]] == [d["f_records"][i], d["f_total"][i], d["f_avg"][i], d["f_max"][i], d["f_min"][i]]


def test_record_stream_errors(oracle, ref_driver, tmp_path):
    p = str(tmp_path / "two_blank.fa")
    with open(p, "w") as f:
        f.write(">a\nACGT\n\n\n")
    rc, _, _ = refdump.ref_records(ref_driver, [p])
    assert rc != 0  # reference aborts (substr(1) on an empty header)
    assert oracle.load_reads([p])[0] == 4
    q = str(tmp_path / "bad.txt")
    with open(q, "w") as f:
        f.write("hello\nworld\n")
    assert refdump.ref_records(ref_driver, [q])[0] != 0
    assert oracle.load_reads([q])[0] == 2
    assert refdump.ref_records(ref_driver, [str(tmp_path / "missing.fa")])[0] != 0
    assert oracle.load_reads([str(tmp_path / "missing.fa")])[0] == 1
