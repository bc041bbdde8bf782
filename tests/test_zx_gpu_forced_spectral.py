"""GPU: --spectral (run_clustering :739-746): connections with score >= 5 from the GPU, the reference's host spectral stage, merge;
engine mirror and CLI against the golden dump of the real reference (ref_driver --force-spectral). Only GPU-verified kernels are
involved (pair count with min_score = 5); sorts before the files that exercise code not yet run on a GPU."""
import os
import subprocess

import numpy as np
import pytest

import datagen
import golden_util

pytestmark = pytest.mark.gpu


def _forced_spectral_case():
    z = np.load(os.path.join(golden_util.GOLDEN, "forced_spectral.npz"))
    fo = z["final_off"].astype(np.int64)
    want = [(int(z["final_id"][i]), z["final_read"][fo[i]:fo[i + 1]].tolist()) for i in range(len(fo) - 1)]
    return z, want


def test_engine_mirror_forced_spectral():
    """--spectral (run_clustering :739-746): connections with score >= 5 from the GPU, the reference's spectral stage on the host"""
    import hga_b200
    z, want = _forced_spectral_case()

    class _Reader(hga_b200.SequenceRecords):
        def __init__(self):
            self.bases, self.seq_off = z["bases"].tobytes(), z["seq_off"]
            n = len(z["seq_off"]) - 1
            self.headers = [b"r%d" % i for i in range(n)]
            self.qualities = [b""] * n

    eng = hga_b200.ReadClusteringEngine(_Reader(), hga_b200.ReadClusteringConfig(scaffold_component_min_size=int(z["min_size"]), force_spectral=True,
                                                                                 spectral_dims=int(z["dims"])))
    ids = eng.run_clustering(z["kmers"], int(z["k"]))
    assert [(i, eng.final_components[i].tolist()) for i in ids] == want
    cx, cy, cs = eng.get_all_connections(5)
    assert np.array_equal(cx, z["conn_x"]) and np.array_equal(cy, z["conn_y"]) and np.array_equal(cs, z["conn_score"])
    eng.close()


def test_cli_forced_spectral(tmp_path):
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hybrid-genome-assembler_b200", "categorization")
    z, want = _forced_spectral_case()
    so = z["seq_off"].astype(np.int64)
    bases = z["bases"].tobytes()
    rp, kp, outdir = str(tmp_path / "reads.fa"), str(tmp_path / "kmers.txt"), str(tmp_path / "out")
    with open(rp, "wb") as f:
        for i in range(len(so) - 1):
            f.write(b">r%d\n" % (i + 1) + bases[so[i]:so[i + 1]] + b"\n")
    with open(kp, "w") as f:
        for v in z["kmers"]:
            f.write(datagen.kmer_to_str(v, int(z["k"])) + "\n")
    r = subprocess.run([exe, rp, "--kmers", kp, "-o", outdir, "--sc_min_size", str(int(z["min_size"])), "--spectral"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Forced spectral clustering took" in r.stdout
    assert sorted(os.listdir(outdir)) == sorted(f"#{fid}.fa" for fid, _ in want)
    for fid, members in want:
        lines = open(os.path.join(outdir, f"#{fid}.fa")).read().split("\n")
        assert [l[1:] for l in lines[0::2] if l] == [f"r{m}" for m in members]
