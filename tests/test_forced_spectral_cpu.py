"""CPU: --spectral (run_clustering :739-746). The product takes get_all_connections(5) from the GPU and runs the host stage
(engine.forced_spectral_components -> hga_spectral_clustering); here the connections come from the C oracle / the fixture and the
result is compared with the real reference (ref_driver --force-spectral: its own spectral_clustering + merge_components)."""
import os

import numpy as np

import golden_util


def _want(ref):
    fo = np.asarray(ref["final_off"]).astype(np.int64)
    return [(int(ref["final_id"][i]), ref["final_read"][fo[i]:fo[i + 1]].tolist()) for i in range(len(fo) - 1)]


def test_forced_spectral_golden():
    import hga_b200
    z = np.load(os.path.join(golden_util.GOLDEN, "forced_spectral.npz"))
    got = hga_b200.engine.forced_spectral_components(z["conn_x"], z["conn_y"], z["conn_score"], int(z["dims"]), int(z["min_size"]))
    assert [(fid, m.tolist()) for fid, m in got.items()] == _want(z)


def test_forced_spectral_live(oracle, ref_driver, tmp_path):
    import datagen
    import hga_b200
    import refdump
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=3000, divergence=0.03, k=19, read_len=150, coverage=20, seed=3, error_rate=0.005, fmt="fastq")
    ref = refdump.run_ref(ref_driver, paths, kp, min_size=30, force_spectral=True)
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    row_off, kid, pos = oracle.scan(reads["seq"], reads["seq_off"], k, kmers)
    inv_off, inv_read = oracle.index(row_off, kid, len(kmers))
    cx, cy, cs = oracle.canonical_sort(*oracle.connections(row_off, kid, inv_off, inv_read, min_score=5))
    assert np.array_equal(cx, ref["conn_x"]) and np.array_equal(cy, ref["conn_y"]) and np.array_equal(cs, ref["conn_score"])
    got = hga_b200.engine.forced_spectral_components(cx, cy, cs, 16, 30)
    assert [(fid, m.tolist()) for fid, m in got.items()] == _want(ref) and len(got) >= 1
