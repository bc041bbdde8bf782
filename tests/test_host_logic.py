"""CPU: host-side logic of the package (k-mer file loader, record loader, sharding, NCCL-id plumbing over gloo)."""
import os
import sys

import numpy as np
import pytest

import datagen
import golden_util


def test_load_text_file_kmers_fixture():
    import hga_b200
    z = np.load(os.path.join(golden_util.GOLDEN, "kmers_fixture.npz"))
    vals, k = hga_b200.load_text_file_kmers(os.path.join(golden_util.GOLDEN, "kmers_fixture.txt"))
    assert k == int(z["k"]) and np.array_equal(vals, z["kmers"])


def test_load_text_file_kmers_matches_oracle(oracle, tmp_path):
    import hga_b200
    g = datagen.random_genome(3000, 3)
    for k in (15, 19, 21, 32):
        p = str(tmp_path / f"k{k}.txt")
        datagen.write_kmers(p, np.unique(datagen.canonical_kmers(g, k))[::5], k)
        a, ka = hga_b200.load_text_file_kmers(p)
        b, kb = oracle.load_kmers(p)
        assert ka == kb == k and np.array_equal(a, b)
    # ragged file: k = length of the LAST line, every line canonicalised with its own length; CRLF kept
    p = str(tmp_path / "ragged.txt")
    with open(p, "wb") as f:
        f.write(b"ACGTACGT\r\nACG\nTTTTTTTTTTTTTTTTTTT\n\nGATTACA")
    a, ka = hga_b200.load_text_file_kmers(p)
    b, kb = oracle.load_kmers(p)
    assert ka == kb == 7 and np.array_equal(a, b)
    assert hga_b200.load_text_file_kmers(str(tmp_path / "missing.txt"))[0].shape[0] == 0


@pytest.mark.parametrize("tag,files", [("fq", ["records_a.fq"]), ("fq_fq", ["records_a.fq", "records_c.fq"]), ("fa", ["records_b.fa"])])
def test_sequence_records_fixture(tag, files):
    import hga_b200
    metas, recs = golden_util.parse_records(os.path.join(golden_util.GOLDEN, f"records_{tag}.expected.txt"))
    r = hga_b200.SequenceRecords([os.path.join(golden_util.GOLDEN, f) for f in files])
    assert r.n_reads == len(recs)
    for i, (rid, h, s, q) in enumerate(recs):
        assert r.headers[i].decode() == h and r.sequence(rid).decode() == s and r.qualities[i].decode() == q
    agg = [m for m in metas if m[0] == "#AGG"][0]
    assert agg[1] == r.meta.filename
    assert [int(v) for v in agg[2:]] == [r.meta.records, r.meta.total_bases, r.meta.avg_read_length, r.meta.max_read_length, r.meta.min_read_length]
    per = [m for m in metas if m[0] == "#META"]
    for m, fm in zip(per, r.file_meta):
        assert m[1] == fm.filename
        assert [int(v) for v in m[2:]] == [fm.records, fm.total_bases, fm.avg_read_length, fm.max_read_length, fm.min_read_length]


def test_sequence_records_matches_oracle_and_errors(oracle, tmp_path):
    import hga_b200
    g = datagen.random_genome(2000, 9)
    p1 = str(tmp_path / "a.fa"); p2 = str(tmp_path / "b.fa")
    datagen.write_fasta(p1, datagen.sample_reads(g, 9, 70, 1), newline="\r\n")
    datagen.write_fasta(p2, datagen.sample_reads(g, 4, 30, 2))
    r = hga_b200.SequenceRecords([p1, p2])
    rc, d = oracle.load_reads([p1, p2])
    assert rc == 0 and r.bases == d["seq"] and np.array_equal(r.seq_off, d["seq_off"]) and np.array_equal(r.file_index, d["file_index"])
    bad = str(tmp_path / "bad.txt")
    open(bad, "w").write("hello\nworld\n")
    with pytest.raises(ValueError, match="Unrecognized file format"):
        hga_b200.SequenceRecords([bad])
    with pytest.raises(ValueError, match="does not exist"):
        hga_b200.SequenceRecords([str(tmp_path / "nope.fa")])
    two = str(tmp_path / "two_blank.fa")
    open(two, "w").write(">a\nACGT\n\n\n")
    with pytest.raises(IndexError):
        hga_b200.SequenceRecords([two])
    assert oracle.load_reads([two])[0] == 4


def test_shard_bounds():
    import hga_b200
    lens = np.random.default_rng(1).integers(100, 60000, size=5000)
    for world in (1, 2, 3, 4, 8):
        b = hga_b200.parallel.shard_bounds(lens, world)
        assert b[0] == 0 and b[-1] == 5000 and len(b) == world + 1 and all(x <= y for x, y in zip(b, b[1:]))
        per = [int(lens[b[i]:b[i + 1]].sum()) for i in range(world)]
        assert max(per) - min(per) <= 2 * 60000
    assert hga_b200.parallel.shard_bounds([], 4) == [0, 0, 0, 0, 0]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import hga_b200
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    uid = hga_b200.parallel.broadcast_unique_id(dist, rank, lambda: bytes(range(128)))
    lens = np.random.default_rng(5).integers(100, 20000, size=1000)
    b = hga_b200.parallel.shard_bounds(lens, world)
    q.put((rank, uid, b[rank], b[rank + 1]))
    dist.destroy_process_group()


def test_two_rank_plumbing_over_gloo():
    """world_size 2 on CPU: every rank gets the same NCCL id and disjoint contiguous read shards"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == bytes(range(128))
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 1000


@pytest.mark.parametrize("seed", range(24))
def test_record_readers_fuzz_against_the_reference(oracle, ref_driver, tmp_path, seed):
    """mangled multi-file inputs (format switches, blank lines, CRLF, truncation): the Python mirror and the C oracle accept / refuse
    exactly what the reference's SequenceRecordIterator accepts / dies on, and yield the same records"""
    import hga_b200
    import refdump
    from test_cli import _mangled_files
    paths = _mangled_files(tmp_path, seed)
    rc, metas, recs = refdump.ref_records(ref_driver, paths)
    orc_rc, d = oracle.load_reads(paths)
    assert (orc_rc != 0) == (rc != 0)
    if rc != 0:
        with pytest.raises((ValueError, IndexError, ZeroDivisionError)):
            hga_b200.SequenceRecords(paths)
        return
    r = hga_b200.SequenceRecords(paths)
    assert r.n_reads == len(recs) == d["n_reads"]
    for i, (rid, h, s, q) in enumerate(recs):
        assert r.headers[i].decode() == h and r.sequence(rid).decode() == s and r.qualities[i].decode() == q
        assert d["seq"][int(d["seq_off"][i]):int(d["seq_off"][i + 1])].decode() == s


@pytest.mark.parametrize("seed", range(16))
def test_kmer_loaders_fuzz_against_the_reference(oracle, ref_driver, tmp_path, seed):
    """the Python mirror and the C oracle of load_text_file_kmers against the reference on mangled k-mer files"""
    import subprocess
    import hga_b200
    import refdump
    from test_cli import _mangled_kmer_file
    p = _mangled_kmer_file(tmp_path, seed)
    try:
        want, want_k = refdump.ref_canon(ref_driver, p)
    except subprocess.CalledProcessError:
        with pytest.raises(ValueError):
            hga_b200.load_text_file_kmers(p)
        with pytest.raises(RuntimeError):
            oracle.load_kmers(p)
        return
    got, k = hga_b200.load_text_file_kmers(p)
    assert k == want_k and np.array_equal(got, want)
    og, ok = oracle.load_kmers(p)
    assert ok == want_k and np.array_equal(og, want)
