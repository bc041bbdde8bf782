"""The C++ host `categorization` (hybrid-genome-assembler_b200/cli): record / k-mer loaders against the golden fixtures from
the real reference (CPU), and the whole executable against the oracle's scaffold components (GPU)."""
import os
import subprocess

import numpy as np
import pytest

import datagen  # noqa: E402
import golden_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "hybrid-genome-assembler_b200", "categorization")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        import importlib
        importlib.import_module("hybrid-genome-assembler_b200.build").build()
    assert os.path.exists(EXE)
    return EXE


@pytest.mark.parametrize("tag,files", [("fq", ["records_a.fq"]), ("fq_fq", ["records_a.fq", "records_c.fq"]), ("fa", ["records_b.fa"])])
def test_cli_record_stream_matches_reference(exe, tag, files):
    out = subprocess.run([exe, "--parse-only"] + [os.path.join(golden_util.GOLDEN, f) for f in files], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(golden_util.GOLDEN, f"records_{tag}.expected.txt")) as f:
        assert out == f.read()


def test_cli_kmer_loader_matches_reference(exe):
    z = np.load(os.path.join(golden_util.GOLDEN, "kmers_fixture.npz"))
    out = subprocess.run([exe, "--dump-kmers", "--kmers", os.path.join(golden_util.GOLDEN, "kmers_fixture.txt")], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == f"#K {int(z['k'])} {len(z['kmers'])}"
    assert [int(v) for v in out[1:] if v] == [int(v) for v in z["kmers"]]


@pytest.mark.parametrize("fmt,threads", [("fasta", 1), ("fastq", 1), ("fastq", 5)])
def test_cli_parallel_export_writes_the_reference_bytes(exe, tmp_path, fmt, threads):
    """export_components (ReadClusteringEngine.cpp:804-826): records in input order, '@h\\nseq\\n+\\nqual' or '>h\\nseq', newline after each"""
    g = datagen.random_genome(5000, 77)
    reads = [datagen.to_ascii(r) for r in datagen.sample_reads(g, 500, 300, 78, length_sigma=0.6, min_len=1)]
    p = str(tmp_path / ("r.fq" if fmt == "fastq" else "r.fa"))
    (datagen.write_fastq if fmt == "fastq" else datagen.write_fasta)(p, reads, prefix="x")
    out = str(tmp_path / "out")
    subprocess.run([exe, "--export-test", p, "-o", out, "--threads", str(threads)], check=True)
    want = {1: b"", 2: b"", 3: b""}
    for i, s in enumerate(reads):
        r = i + 1
        if r % 7 == 0:
            continue
        rec = (f"@x{i}\n{s}\n+\n{'I' * len(s)}\n" if fmt == "fastq" else f">x{i}\n{s}\n").encode()
        want[1 + r % 3] += rec
    assert sorted(os.listdir(out)) == ["#1.fa", "#2.fa", "#3.fa"]
    for c, data in want.items():
        assert open(os.path.join(out, f"#{c}.fa"), "rb").read() == data


def test_cli_option_styles(exe):
    """boost::program_options' default style: --name=value, -kvalue, unambiguous prefixes of long names"""
    z = np.load(os.path.join(golden_util.GOLDEN, "kmers_fixture.npz"))
    kf = os.path.join(golden_util.GOLDEN, "kmers_fixture.txt")
    want = subprocess.run([exe, "--dump-kmers", "--kmers", kf], capture_output=True, text=True, check=True).stdout
    assert want.startswith(f"#K {int(z['k'])} ")
    for args in (["--dump-kmers", f"--kmers={kf}"], ["--dump-kmers", f"-k{kf}"], ["--dump-kmers", "-k", kf], ["--dump-k", "--kmer", kf]):
        assert subprocess.run([exe] + args, capture_output=True, text=True, check=True).stdout == want
    r = subprocess.run([exe, "--s", "x"], capture_output=True, text=True)          # --sc_..., --spectral..., --scaffolds-only
    assert r.returncode != 0 and "ambiguous" in r.stderr
    r = subprocess.run([exe, "--nonsense"], capture_output=True, text=True)
    assert r.returncode != 0 and "unrecognised option" in r.stderr


def test_cli_errors(exe, tmp_path):
    bad = tmp_path / "bad.txt"
    bad.write_text("hello\nworld\n")
    r = subprocess.run([exe, "--parse-only", str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "Unrecognized file format" in r.stderr
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "You need to specify path to kmers" in r.stderr


def test_jf_occurrences_without_a_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked tests")
    exe2 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hybrid-genome-assembler_b200", "jf_occurrences")
    p = str(tmp_path / "x.fa")
    open(p, "w").write(">a\nACGTACGTACGTACGTACGT\n")
    r = subprocess.run([exe2, p, "-k", "5", "-o", str(tmp_path / "o.txt")], input="2 5 1\n", capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr and not os.path.exists(str(tmp_path / "o.txt"))
    assert "Size of kmer to analyze & select" in subprocess.run([exe2, "--help"], capture_output=True, text=True).stdout


@pytest.mark.gpu
def test_cli_end_to_end_components(exe, oracle, tmp_path):
    paths, kp = datagen.make_diploid_case(str(tmp_path / "c"), genome_size=30000, divergence=0.03, k=19, read_len=1200, coverage=12, seed=5,
                                          error_rate=0.01, fmt="fastq", length_sigma=0.3)
    outdir = str(tmp_path / "out")
    r = subprocess.run([exe] + paths + ["--kmers", kp, "-o", outdir, "--sc_min_size", "5", "--scaffolds-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for label in ("Index construction took", "Calculation of connections between reads took", "Union-find took", "Exported"):
        assert label in r.stdout
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    ref = oracle.run(reads["seq"], reads["seq_off"], k, kmers, fraction=0.15, min_size=5)
    co, cm, _, _, _ = ref["comp"]
    co = co.astype(np.int64)
    hdr_off = reads["hdr_off"].astype(np.int64)
    want = sorted(sorted(reads["hdr"][hdr_off[i - 1]:hdr_off[i]].decode() for i in cm[co[c]:co[c + 1]]) for c in range(len(co) - 1))
    got = []
    for f in sorted(os.listdir(outdir)):
        assert f.startswith("#") and f.endswith(".fa")
        lines = open(os.path.join(outdir, f)).read().split("\n")
        got.append(sorted(l[1:] for l in lines[0::4] if l))      # FASTQ records: 4 lines each
    assert sorted(got) == want
    assert f"Exported {len(want)} components" in r.stdout


@pytest.mark.gpu
def test_cli_end_to_end_final_components(exe, oracle, tmp_path):
    """run without the tail / spectral block (--no-tail-block): scaffold components -> merge -> enrichment -> export under the surviving
    component ids; the default run, block included, is tests/test_zy_gpu_tail_block.py"""
    import oracle_lib
    paths, kp = datagen.make_diploid_case(str(tmp_path / "c"), genome_size=20000, divergence=0.03, k=19, read_len=150, coverage=30, seed=9,
                                          error_rate=0.005, fmt="fastq")
    outdir = str(tmp_path / "out")
    r = subprocess.run([exe] + paths + ["--kmers", kp, "-o", outdir, "--no-tail-block"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Merging of initial components" in r.stdout and "Union-find took" in r.stdout
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    res = oracle.run(reads["seq"], reads["seq_off"], k, kmers, fraction=0.15, min_size=30)
    e = oracle_lib.enrich(oracle, res, len(kmers), min_size=30, enrich_min=20)
    assert len(e["final_id"]) > 2 and "scaffold components; the tail / spectral merge" in r.stderr
    hdr_off = reads["hdr_off"].astype(np.int64)
    want = {f"#{int(fid)}.fa": [reads["hdr"][hdr_off[i - 1]:hdr_off[i]].decode() for i in members] for fid, members in zip(e["final_id"], e["final_reads"])}
    assert sorted(os.listdir(outdir)) == sorted(want)
    for f, hdrs in want.items():
        lines = open(os.path.join(outdir, f)).read().split("\n")
        assert [l[1:] for l in lines[0::4] if l] == hdrs           # input order = ascending read id
    assert f"Exported {len(want)} components" in r.stdout


def _mangled_files(tmp_path, seed):
    """A few small FASTA / FASTQ files with the irregularities the reference's reader reacts to: format switches between files, a
    trailing blank line, CRLF line ends, a file without final newline, an empty header line, too few lines."""
    rng = np.random.default_rng(seed)
    g = datagen.random_genome(800, seed)
    paths = []
    for i in range(int(rng.integers(1, 4))):
        reads = [datagen.to_ascii(r) for r in datagen.sample_reads(g, int(rng.integers(1, 5)), int(rng.integers(5, 60)), seed * 10 + i)]
        fq = bool(rng.integers(0, 2))
        nl = "\r\n" if rng.integers(0, 4) == 0 else "\n"
        p = str(tmp_path / f"f{seed}_{i}.{'fq' if fq else 'fa'}")
        (datagen.write_fastq if fq else datagen.write_fasta)(p, reads, prefix=f"s{seed}_{i}_", newline=nl)
        tweak = int(rng.integers(0, 7))
        data = open(p, "rb").read()
        if tweak == 1:
            data += b"\n"                                   # one trailing blank line: tolerated at the end of the LAST file only
        elif tweak == 2:
            data = data.rstrip(b"\r\n")                      # no final newline
        elif tweak == 3:
            data += b"\n\n"                                  # two blank lines: header.substr(1) on an empty header
        elif tweak == 4 and fq:
            data = data[:data.rfind(b"+")]                   # truncated last record
        elif tweak == 5:
            data = data.replace(b"A", b"n", 3)               # non-ACGT bytes are just bytes to the reader
        open(p, "wb").write(data)
        paths.append(p)
    return paths


@pytest.mark.parametrize("seed", range(24))
def test_cli_reader_fuzz_against_the_reference(exe, ref_driver, tmp_path, seed):
    """record stream, per-file and aggregate meta data, and WHETHER the run fails, for mangled multi-file inputs: the C++ loader of
    the CLI against the reference's SequenceRecordIterator (ref_driver records)"""
    import refdump
    paths = _mangled_files(tmp_path, seed)
    rc, metas, recs = refdump.ref_records(ref_driver, paths)
    r = subprocess.run([exe, "--parse-only"] + paths, capture_output=True)
    if rc != 0:
        assert r.returncode != 0, "the reference aborts on this input, the CLI must fail too"
        return
    assert r.returncode == 0, r.stderr.decode()[-300:]
    got_meta, got_recs = [], []
    for line in r.stdout.decode().split("\n"):
        if line.startswith("#META ") or line.startswith("#AGG "):
            got_meta.append(line.split(" "))
        elif line:
            i, h, s, q = line.split("\t")
            got_recs.append((int(i), h, s, q))
    assert got_recs == recs
    assert got_meta == metas


def _mangled_kmer_file(tmp_path, seed):
    """k-mer file with lines of different lengths, duplicates, lowercase / N / CRLF bytes, an empty line, no final newline, a line > 32"""
    rng = np.random.default_rng(500 + seed)
    k = int(rng.integers(3, 33))
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTNacgt", dtype=np.uint8)
    lines = []
    for _ in range(int(rng.integers(1, 40))):
        L = k if rng.integers(0, 5) else int(rng.integers(1, 33))
        lines.append(alphabet[rng.integers(0, alphabet.shape[0], size=L)].tobytes())
    if rng.integers(0, 4) == 0:
        lines.insert(int(rng.integers(0, len(lines))), b"")
    if rng.integers(0, 6) == 0:
        lines.append(b"ACGT" * 9)                                   # 36 > 32: "Kmer size is too big"
    if lines and rng.integers(0, 3) == 0:
        lines.append(lines[0])                                      # duplicate
    nl = b"\r\n" if rng.integers(0, 5) == 0 else b"\n"
    data = nl.join(lines) + (b"" if rng.integers(0, 3) == 0 else nl)
    p = str(tmp_path / f"k{seed}.txt")
    open(p, "wb").write(data)
    return p


@pytest.mark.parametrize("seed", range(16))
def test_cli_kmer_loader_fuzz_against_the_reference(exe, ref_driver, tmp_path, seed):
    """k is the length of the LAST line, every line is canonicalised with its own length (read_clustering.cpp:18-33)"""
    import refdump
    p = _mangled_kmer_file(tmp_path, seed)
    r = subprocess.run([exe, "--dump-kmers", "--kmers", p], capture_output=True, text=True)
    try:
        want, want_k = refdump.ref_canon(ref_driver, p)
    except subprocess.CalledProcessError:
        assert r.returncode != 0, "the reference aborts on this k-mer file, the CLI must fail too"
        return
    assert r.returncode == 0, r.stderr[-300:]
    out = r.stdout.split("\n")
    assert out[0] == f"#K {want_k} {len(want)}"
    assert [int(v) for v in out[1:] if v] == [int(v) for v in want]
