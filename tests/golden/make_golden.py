"""Generates the golden fixtures under tests/golden/ from the REAL reference (oracle/_ref/ref_driver = unmodified
/root/reference sources + shim). Run in the build container (needs /root/reference to build the driver):

    python tests/golden/make_golden.py

Each <case>.npz holds the inputs (bases, seq_off, sorted canonical k-mers, k, fraction, min_size) and every stage
output the driver dumps (hit multisets, first positions, inverted index, canonical directed connections, cut,
components, spanning forests). records_*.txt hold the record stream of the small FASTA/FASTQ files next to them."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import datagen  # noqa: E402
import oracle_lib  # noqa: E402
import refdump  # noqa: E402

DRIVER = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "ref_driver")


def save_case(name, paths, kp, fraction, min_size, enrich=0, full=False):
    orc = oracle_lib.load()
    rc, reads = orc.load_reads(paths)
    assert rc == 0
    kmers, k = orc.load_kmers(kp)
    ref = refdump.run_ref(DRIVER, paths, kp, fraction=fraction, min_size=min_size, enrich=enrich, full=full)
    keep = {k_: v for k_, v in ref.items() if isinstance(v, np.ndarray)}
    if enrich:   # the stages before the merge are pinned by the other fixtures; keep these files small
        keep = {k_: v for k_, v in keep.items() if k_.split("_")[0] in ("core", "purged", "econn", "final", "comp", "tconn", "spectral")}
    scal = {k_: v for k_, v in ref.items() if not isinstance(v, np.ndarray) and not k_.endswith("_ms")}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), bases=np.frombuffer(reads["seq"], dtype=np.uint8), seq_off=reads["seq_off"],
                        kmers=kmers, k=np.int64(k), fraction=np.float64(fraction), min_size=np.int64(min_size), enrich=np.int64(enrich),
                        **{"ref_" + k_: v for k_, v in keep.items()}, **{"meta_" + k_: np.int64(v) for k_, v in scal.items()})
    print(name, {k_: v for k_, v in scal.items()})


def main_enrich():
    """SURVEY §8f-1 fixtures: merge_components + enrichment + final merge (ref_driver --enrich)."""
    with tempfile.TemporaryDirectory() as d:
        paths, kp = datagen.make_diploid_case(os.path.join(d, "e1"), genome_size=8000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7,
                                              error_rate=0.005, fmt="fastq")
        save_case("enrich_short", paths, kp, 0.15, 30, enrich=20)
        paths, kp = datagen.make_diploid_case(os.path.join(d, "e2"), genome_size=20000, divergence=0.02, k=15, read_len=2000, coverage=10,
                                              seed=21, error_rate=0.05, length_sigma=0.5)
        save_case("enrich_long", paths, kp, 0.15, 5, enrich=20)


def main_full():
    """SURVEY §8f-2 fixtures: the WHOLE of run_clustering after the scaffold union_find, tail / spectral block included (ref_driver
    --enrich 20 --full): tail connections, spectral clusters, and the state after the merge of the clusters (cores, merged k-mer lists,
    purged index), enrichment connections, final components."""
    with tempfile.TemporaryDirectory() as d:
        paths, kp = datagen.make_diploid_case(os.path.join(d, "f1"), genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7,
                                              error_rate=0.005, fmt="fastq")
        save_case("full_short", paths, kp, 0.15, 30, enrich=20, full=True)
        paths, kp = datagen.make_diploid_case(os.path.join(d, "f2"), genome_size=60000, divergence=0.02, k=15, read_len=2000, coverage=12,
                                              seed=21, error_rate=0.05, length_sigma=0.5)
        save_case("full_long", paths, kp, 0.15, 5, enrich=20, full=True)


def main_forced_spectral():
    """--spectral (run_clustering :739-746): get_all_connections(5), spectral clustering of the whole data set, merge, ids >= min size
    (ref_driver --force-spectral)."""
    orc = oracle_lib.load()
    with tempfile.TemporaryDirectory() as d:
        paths, kp = datagen.make_diploid_case(d, genome_size=20000, divergence=0.02, k=15, read_len=2000, coverage=10, seed=21, error_rate=0.05, length_sigma=0.5)
        ref = refdump.run_ref(DRIVER, paths, kp, min_size=5, force_spectral=True)
        rc, reads = orc.load_reads(paths)
        kmers, k = orc.load_kmers(kp)
    np.savez_compressed(os.path.join(HERE, "forced_spectral.npz"), bases=np.frombuffer(reads["seq"], dtype=np.uint8), seq_off=reads["seq_off"], kmers=kmers,
                        k=np.int64(k), min_size=np.int64(5), dims=np.int64(16), conn_x=ref["conn_x"], conn_y=ref["conn_y"], conn_score=ref["conn_score"],
                        final_id=ref["final_id"], final_off=ref["final_off"], final_read=ref["final_read"])
    print("forced_spectral", len(ref["conn_x"]), "directed connections,", ref["final_components"], "final components")


def main_spectral():
    """SURVEY §8f-2 fixtures: strong tail connections -> spectral clusters, from ref_driver --full (the reference's own
    spectral_clustering + lib/clustering, compiled against the Eigen2 stand-in of oracle/shim)."""
    cases = {"spectral_a": dict(genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7, error_rate=0.005, fmt="fastq"),
             "spectral_b": dict(genome_size=100000, divergence=0.03, k=19, read_len=400, coverage=25, seed=10, error_rate=0.01)}
    for name, kw in cases.items():
        with tempfile.TemporaryDirectory() as d:
            paths, kp = datagen.make_diploid_case(d, **kw)
            ref = refdump.run_ref(DRIVER, paths, kp, enrich=20, full=True)
        m = ref["tconn_score"] > 5
        np.savez_compressed(os.path.join(HERE, name + ".npz"), conn_x=ref["tconn_x"][m], conn_y=ref["tconn_y"][m], conn_score=ref["tconn_score"][m],
                            dims=np.int64(16), cluster_off=ref["spectral_off"], cluster_member=ref["spectral_member"], cluster_first=ref["spectral_first"])
        print(name, int(m.sum()), "connections,", len(ref["spectral_off"]) - 1, "clusters")


def main_tails():
    """SURVEY §8f-2 fixture: reads + k-mers -> tail connections between the scaffold components, from ref_driver --full."""
    orc = oracle_lib.load()
    with tempfile.TemporaryDirectory() as d:
        paths, kp = datagen.make_diploid_case(d, genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=7, error_rate=0.005, fmt="fastq")
        ref = refdump.run_ref(DRIVER, paths, kp, enrich=20, full=True, min_size=30)
        rc, reads = orc.load_reads(paths)
        kmers, k = orc.load_kmers(kp)
    np.savez_compressed(os.path.join(HERE, "tails_a.npz"), bases=np.frombuffer(reads["seq"], dtype=np.uint8), seq_off=reads["seq_off"], kmers=kmers, k=np.int64(k),
                        min_size=np.int64(30), amplification_min_score=np.int64(40), tconn_x=ref["tconn_x"], tconn_y=ref["tconn_y"], tconn_score=ref["tconn_score"])
    print("tails_a", len(ref["tconn_x"]), "tail connections,", ref["scaffold_components"], "scaffold components")


def main():
    if "--spectral-only" in sys.argv:
        return main_spectral()
    if "--tails-only" in sys.argv:
        return main_tails()
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-C", os.path.dirname(os.path.dirname(DRIVER)), "ref"], check=True)
    if "--enrich-only" in sys.argv:
        return main_enrich()
    if "--full-only" in sys.argv:
        return main_full()
    if "--forced-spectral-only" in sys.argv:
        return main_forced_spectral()
    main_enrich()
    main_full()
    main_forced_spectral()
    main_spectral()
    main_tails()
    with tempfile.TemporaryDirectory() as d:
        # KAT 2 of SURVEY §8c: multiplicity
        rng = np.random.default_rng(7)
        X = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8)); Y = datagen.to_ascii(rng.integers(0, 4, 40, dtype=np.uint8))
        rp = os.path.join(d, "kat2.fa"); kp = os.path.join(d, "kat2.txt")
        datagen.write_fasta(rp, [X + Y + X, X, Y + X + X + X, Y, "ACGT"])
        with open(kp, "w") as f:
            for i in range(5):
                f.write(X[i:i + 15] + "\n")
        save_case("kat2", [rp], kp, 1.0, 1)

        # config-1-like: short reads, FASTQ, two files
        paths, kp = datagen.make_diploid_case(os.path.join(d, "c1"), genome_size=4000, divergence=0.03, k=19, read_len=150, coverage=20, seed=3,
                                              error_rate=0.005, fmt="fastq")
        save_case("config1_mini", paths, kp, 0.15, 30)

        # long reads with errors, k sweep
        for k in (15, 21):
            paths, kp = datagen.make_diploid_case(os.path.join(d, f"lr{k}"), genome_size=20000, divergence=0.02, k=k, read_len=2000, coverage=10,
                                                  seed=k, error_rate=0.05, length_sigma=0.5)
            save_case(f"longreads_k{k}", paths, kp, 0.15, 3)

        # non-ACGT bytes, lowercase reads, CRLF line ends
        a = datagen.random_genome(3000, 11)
        reads = [datagen.to_ascii(r) for r in datagen.sample_reads(a, 120, 100, 12)]
        rng = np.random.default_rng(13)
        mangled = []
        for i, r in enumerate(reads):
            r = list(r)
            for j in rng.integers(0, len(r), size=3):
                r[j] = "NnacgtRY*"[int(rng.integers(0, 9))]
            mangled.append("".join(r).lower() if i % 17 == 0 else "".join(r))
        rp = os.path.join(d, "exc.fa"); kp = os.path.join(d, "exc.txt")
        datagen.write_fasta(rp, mangled, newline="\r\n")
        datagen.write_kmers(kp, np.unique(datagen.canonical_kmers(a, 15))[::3], 15)
        save_case("exceptions_crlf", [rp], kp, 0.15, 2)

        # record-stream fixtures (files are committed next to their expected record dump)
        g = datagen.random_genome(500, 5)
        fq = os.path.join(HERE, "records_a.fq"); fa = os.path.join(HERE, "records_b.fa"); fq2 = os.path.join(HERE, "records_c.fq")
        datagen.write_fastq(fq, datagen.sample_reads(g, 5, 40, 1), prefix="x")
        datagen.write_fasta(fa, datagen.sample_reads(g, 4, 55, 2), prefix="y")
        datagen.write_fastq(fq2, datagen.sample_reads(g, 3, 30, 3), prefix="z")
        with open(fq2, "a") as f:
            f.write("\n")   # one trailing blank line is tolerated
        for tag, paths in (("fq", [fq]), ("fq_fq", [fq, fq2]), ("fa", [fa])):
            sub = os.path.join(d, "rec_" + tag)
            os.makedirs(sub)
            subprocess.run([DRIVER, "records", sub] + paths, check=True, stdout=subprocess.DEVNULL)
            shutil.copy(os.path.join(sub, "records.txt"), os.path.join(HERE, f"records_{tag}.expected.txt"))

        # k-mer file fixture: duplicates, both strands, N inside a k-mer, no trailing newline
        kf = os.path.join(HERE, "kmers_fixture.txt")
        with open(kf, "w") as f:
            f.write("ACGTACGTACGTACGTACG\nCGTACGTACGTACGTACGT\nACGTACGTACGTACGTACG\nTTTTTTTTTTTTTTTTTTT\nACGTNNNNACGTACGTACG\nGATTACAGATTACAGATTA")
        vals, k = refdump.ref_canon(DRIVER, kf)
        np.savez(os.path.join(HERE, "kmers_fixture.npz"), kmers=vals, k=np.int64(k))
        # KmerIterator known answers (SURVEY §8c KAT 1 + random strings with exceptions)
        seqs, ks, outs_k, outs_p = [], [], [], []
        rng = np.random.default_rng(99)
        alphabet = np.frombuffer(b"ACGTACGTACGTNacgtX", dtype=np.uint8)
        cases = [("ACGTNACGTTTGACCAGTA", 3), ("AC", 3)]
        for k in (1, 2, 15, 19, 21, 31, 32):
            for L in (k, k + 1, 97):
                cases.append((alphabet[rng.integers(0, alphabet.shape[0], size=L)].tobytes().decode(), k))
        for s, k in cases:
            km, pos = refdump.ref_kmeriter(DRIVER, s, k)
            seqs.append(s); ks.append(k); outs_k.append(km); outs_p.append(pos)
        np.savez_compressed(os.path.join(HERE, "kmeriter_fixture.npz"), seqs=np.array(seqs), ks=np.array(ks, dtype=np.int64),
                            off=np.cumsum([0] + [len(o) for o in outs_k]).astype(np.int64), kmers=np.concatenate(outs_k), pos=np.concatenate(outs_p))


if __name__ == "__main__":
    main()
