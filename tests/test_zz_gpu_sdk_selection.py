"""GPU: SDK selection: hga_count_kmers against exact counting in numpy and against the reference's own KmerIterator
(oracle/_ref/occ_driver count), and the jf_occurrences program end to end."""
import os
import subprocess

import numpy as np
import pytest

import datagen
import golden_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [11, 19, 32])
def test_cuda_kmer_counts_are_exact(k):
    """SURVEY §8f-4, the jellyfish step: hga_count_kmers against exact counting in numpy (canonical k-mers, count >= 2, ascending;
    windows with a non-ACGT byte skipped, lowercase accepted)"""
    import hga_b200
    from test_sdk_selection_cpu import exact_counts
    rng = np.random.default_rng(k)
    g = datagen.random_genome(4000, 100 + k)
    reads = [datagen.to_ascii(r) for r in datagen.sample_reads(g, 300, 180, 200 + k, error_rate=0.01)]
    reads[3] = reads[3][:40] + "N" + reads[3][41:]                 # a window breaker
    reads[5] = reads[5].lower()                                    # jellyfish counts lowercase bases
    reads[7] = reads[7][:k - 1]                                    # shorter than k
    reads[9] = ""
    seq = "".join(reads).encode()
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    km, ct = hga_b200.capi.count_kmers(seq, off, k, min_count=2)
    wk, wc = exact_counts(seq, off, k, 2)
    assert np.array_equal(km, wk) and np.array_equal(ct, wc) and len(km) > 100
    km1, ct1 = hga_b200.capi.count_kmers(seq, off, k, min_count=1)
    wk1, wc1 = exact_counts(seq, off, k, 1)
    assert np.array_equal(km1, wk1) and np.array_equal(ct1, wc1) and int(ct1.sum()) == int(wc1.sum())
    # the key-range passes of large inputs (quantiles from a sample, one pass per range, overflowing ranges split), forced with a tiny budget
    os.environ["HGA_COUNT_CHUNK"] = "4096"
    try:
        km2, ct2 = hga_b200.capi.count_kmers(seq, off, k, min_count=2)
    finally:
        del os.environ["HGA_COUNT_CHUNK"]
    assert np.array_equal(km2, wk) and np.array_equal(ct2, wc)


def test_key_range_passes_equal_the_single_pass_on_a_larger_input():
    """40 Mbases with a budget of 2^21 keys per pass (~30 key ranges; the range holding the poly-A k-mer overflows and is halved down to that
    single value, whose count is then the number of matches) against the single pass of the same input: identical k-mers and counts, ascending"""
    import hga_b200
    rng = np.random.default_rng(77)
    g = datagen.random_genome(400000, 501)
    g[1000:51000] = 0                                             # 50 kb of poly-A at 100x: ONE k-mer with ~5 M occurrences, more than a whole budget
    reads = datagen.sample_reads(g, 4000, 10000, 502, error_rate=0.02, length_sigma=0.3, max_len=60000)
    seq = b"".join(datagen.to_ascii(r).encode() for r in reads)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    assert off[-1] > 30_000_000
    km, ct = hga_b200.capi.count_kmers(seq, off, 19, min_count=2)
    os.environ["HGA_COUNT_CHUNK"] = str(1 << 21)
    try:
        km2, ct2 = hga_b200.capi.count_kmers(seq, off, 19, min_count=2)
    finally:
        del os.environ["HGA_COUNT_CHUNK"]
    assert len(km) > 100000 and np.all(km[1:] > km[:-1]) and km[0] == 0 and ct[0] > (1 << 21)
    assert np.array_equal(km, km2) and np.array_equal(ct, ct2)


def test_cli_jf_occurrences_exports_the_kmers_file(oracle, tmp_path):
    """jf_occurrences (the --kmers producer): per-file GPU counts -> merge -> specificity table -> export of a count range, against
    exact numpy counts pushed through the host functions the reference's reader pins (tests/test_sdk_selection_cpu.py)"""
    import hga_b200
    from test_sdk_selection_cpu import exact_counts
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hybrid-genome-assembler_b200", "jf_occurrences")
    k = 15
    paths, _ = datagen.make_diploid_case(str(tmp_path), genome_size=6000, divergence=0.03, k=k, read_len=300, coverage=8, seed=6, error_rate=0.01)
    per_file = []
    for p in paths:
        rc, reads = oracle.load_reads([p])
        per_file.append(exact_counts(reads["seq"], reads["seq_off"], k))
    km, total, largest, files = hga_b200.capi.sdk_merge(per_file)
    out = str(tmp_path / "sdk.txt")
    r = subprocess.run([exe] + paths + ["-k", str(k), "-o", out], input="3 12 1.0\n", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sel, n_sel, n_disc = hga_b200.capi.sdk_select(total, files, 3, 12)
    assert [l.strip() for l in open(out) if l.strip()] == [datagen.kmer_to_str(v, k) for v in km[sel]] and n_sel > 0
    assert f"{n_disc} out of {n_sel} exported kmers are discriminative" in r.stdout
    t, o, u = hga_b200.capi.sdk_specificity(total, largest)
    table = [tuple(l.split()) for l in r.stdout.splitlines() if len(l.split()) == 3 and l[0].isdigit()]
    assert [(float(a), int(b), int(c)) for a, b, c in table] == [(round(float(a), 2), int(b), int(c)) for a, b, c in zip(t, o, u)]


@pytest.mark.parametrize("k", [15, 19, 32])
def test_cuda_kmer_counts_match_the_reference_kmer_iterator(oracle, tmp_path, k):
    """hga_count_kmers against reference CODE: the reference's record stream + rolling canonical k-mer + std::map (occ_driver count) on
    ACGT-only reads, where its rule for other bytes cannot differ from jellyfish's (tests/test_sdk_selection_cpu.py pins the same)"""
    import hga_b200
    from test_sdk_selection_cpu import reference_counts
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "occ_driver")
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/occ_driver not built")
    g = datagen.random_genome(20000, 500 + k)
    reads = datagen.sample_reads(g, 600, 400, 600 + k, error_rate=0.02, length_sigma=0.4)
    p = str(tmp_path / "reads.fa")
    datagen.write_fasta(p, reads)
    rc, rd = oracle.load_reads([p])
    for mc in (1, 2):
        km, ct = hga_b200.capi.count_kmers(rd["seq"], rd["seq_off"], k, min_count=mc)
        rk, rcnt = reference_counts(drv, [p], k, mc)
        assert np.array_equal(km, rk) and np.array_equal(ct, rcnt) and len(rk) > 1000
