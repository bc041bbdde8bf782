"""CPU: the host half of the SDK selection (hga_host_sdk_merge / _specificity / _select) against the reference's own
occurrences/JellyfishOccurrenceReader.cpp compiled unmodified (oracle/_ref/occ_driver). jellyfish is not in the image: the
per-file dumps the reader merges (`<read file>_<k>-mers_sorted`, which make it skip jellyfish, :19-24) are written here from exact
numpy counts (canonical k-mers with count >= 2) - the same numbers hga_count_kmers has to produce on the GPU
(tests/test_zz_gpu_sdk_selection.py compares it with this counter)."""
import os
import subprocess

import numpy as np
import pytest

import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = np.full(256, 255, dtype=np.uint8)
for i, ch in enumerate(b"ACGT"):
    CODE[ch] = i
    CODE[ch + 32] = i


def exact_counts(seq, seq_off, k, min_count=2):
    """canonical k-mers of the reads (windows with a non-ACGT byte skipped) that occur >= min_count times, ascending, with counts"""
    seq_off = np.asarray(seq_off).astype(np.int64)
    allk = []
    raw = np.frombuffer(seq, dtype=np.uint8)
    for r in range(len(seq_off) - 1):
        codes = CODE[raw[seq_off[r]:seq_off[r + 1]]]
        if len(codes) < k:
            continue
        km = datagen.canonical_kmers(np.where(codes == 255, 0, codes).astype(np.uint8), k)
        bad = np.convolve((codes == 255).astype(np.int64), np.ones(k, dtype=np.int64), mode="valid") > 0
        allk.append(km[~bad])
    allk = np.concatenate(allk) if allk else np.zeros(0, np.uint64)
    u, c = np.unique(allk, return_counts=True)
    m = c >= min_count
    return u[m], c[m].astype(np.uint32)


@pytest.fixture(scope="module")
def occ_driver():
    path = os.path.join(ROOT, "oracle", "_ref", "occ_driver")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
        else:
            pytest.skip("oracle/_ref/occ_driver not built and /root/reference absent")
    return path


def _case(oracle, tmp_path, k, n_files, seed):
    if n_files == 2:
        paths, _ = datagen.make_diploid_case(str(tmp_path), genome_size=6000, divergence=0.03, k=k, read_len=300, coverage=8, seed=seed, error_rate=0.01)
    else:
        paths, _ = datagen.make_polyploid_case(str(tmp_path), genome_size=5000, divergence=0.03, k=k, read_len=400, coverage=6, seed=seed, n_haplotypes=n_files,
                                               error_rate=0.01)
    per_file = []
    for p in paths:
        rc, reads = oracle.load_reads([p])
        km, ct = exact_counts(reads["seq"], reads["seq_off"], k)
        per_file.append((km, ct))
        with open(f"{p}_{k}-mers_sorted", "w") as f:                      # what run_jellyfish.sh leaves behind
            for v, c in zip(km, ct):
                f.write(f"{datagen.kmer_to_str(v, k)} {int(c)}\n")
    return paths, per_file


@pytest.mark.parametrize("k,n_files,seed", [(11, 2, 5), (15, 2, 6), (13, 4, 7)])
def test_merge_specificity_and_export_match_the_reference_reader(oracle, occ_driver, tmp_path, k, n_files, seed):
    import hga_b200
    paths, per_file = _case(oracle, tmp_path, k, n_files, seed)
    km, total, largest, files = hga_b200.capi.sdk_merge(per_file)
    assert np.all(np.diff(km.astype(np.int64)) > 0) and len(km) == len(np.unique(np.concatenate([p[0] for p in per_file])))
    # specificity table
    r = subprocess.run([occ_driver, "specificity", str(k)] + paths, capture_output=True, text=True, check=True)
    want = [(float(a), int(b), int(c)) for a, b, c in (line.split() for line in r.stdout.splitlines() if line.strip())]
    t, o, u = hga_b200.capi.sdk_specificity(total, largest)
    assert [(round(float(a), 2), int(b), int(c)) for a, b, c in zip(t, o, u)] == want and len(want) > 3
    # export of a count range (percent = 1: every k-mer in range)
    lo, hi = int(np.percentile(total, 30)), int(np.percentile(total, 90))
    out = str(tmp_path / "exported.txt")
    r = subprocess.run([occ_driver, "export", str(k), str(lo), str(hi), "1.0", out] + paths, capture_output=True, text=True, check=True)
    sel, n_sel, n_disc = hga_b200.capi.sdk_select(total, files, lo, hi)
    exported = [line.strip() for line in open(out) if line.strip()]
    assert exported == [datagen.kmer_to_str(v, k) for v in km[sel]] and n_sel == len(exported) > 0
    assert f"{n_disc} out of {n_sel} exported kmers are discriminative" in r.stdout
    # ... and the exported file is a --kmers file the categorization loader takes (canonical, one per line)
    kk, k2 = oracle.load_kmers(out)
    assert k2 == k and np.array_equal(kk, km[sel])


def test_select_sampling_is_seeded():
    import hga_b200
    total = np.arange(1, 2001, dtype=np.uint32)
    files = np.ones(2000, dtype=np.uint32)
    a, na, _ = hga_b200.capi.sdk_select(total, files, 100, 1500, percent=0.25, seed=7)
    b, nb, _ = hga_b200.capi.sdk_select(total, files, 100, 1500, percent=0.25, seed=7)
    c, nc, _ = hga_b200.capi.sdk_select(total, files, 100, 1500, percent=0.25, seed=8)
    assert np.array_equal(a, b) and na == nb and not np.array_equal(a, c)
    assert not a[:99].any() and not a[1500:].any() and 250 < na < 450


def test_merge_rejects_unsorted_lists():
    import hga_b200
    with pytest.raises(hga_b200.HgaError):
        hga_b200.capi.sdk_merge([(np.array([5, 3], dtype=np.uint64), np.array([2, 2], dtype=np.uint32))])


def reference_counts(occ_driver, paths, k, min_count):
    """canonical k-mer counts by the reference's OWN code: SequenceRecordIterator + KmerIterator + std::map (occ_driver count)"""
    r = subprocess.run([occ_driver, "count", str(k), str(min_count)] + list(paths), capture_output=True, text=True, check=True)
    rows = [l.split() for l in r.stdout.replace("\r", "\n").splitlines() if len(l.split()) == 2 and l.split()[0].isdigit() and l.split()[1].isdigit()]
    return np.array([int(a) for a, _ in rows], dtype=np.uint64), np.array([int(b) for _, b in rows], dtype=np.uint32)


@pytest.mark.parametrize("k", [11, 19, 31, 32])
def test_exact_counts_match_the_reference_kmer_iterator(oracle, occ_driver, tmp_path, k):
    """the counting step pinned with reference code: on reads made of A C G T only, jellyfish's rule (skip windows with another byte) and
    KmerIterator's rule (another byte reads as code 0) coincide, so the reference's own iterator + std::map give jellyfish's counts. The
    numpy checker the GPU counting is compared with must agree with them, for every k-mer and every count."""
    g = datagen.random_genome(5000, 300 + k)
    reads = datagen.sample_reads(g, 250, 200, 400 + k, error_rate=0.01)
    p = str(tmp_path / "reads.fa")
    datagen.write_fasta(p, reads)
    rc, rd = oracle.load_reads([p])
    for mc in (1, 2):
        wk, wc = exact_counts(rd["seq"], rd["seq_off"], k, mc)
        rk, rcnt = reference_counts(occ_driver, [p], k, mc)
        assert np.array_equal(wk, rk) and np.array_equal(wc, rcnt) and len(rk) > 100
