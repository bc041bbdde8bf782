"""CPU: the product's host-side spectral clustering (hga_spectral_clustering, SURVEY §8f-2 first piece) against the reference's own
spectral code: live through oracle/_ref/ref_driver --full (lib/clustering compiled unmodified against the Eigen2 stand-in, both
sides built without fused multiply-add) and against the committed golden fixtures generated from it."""
import os

import numpy as np
import pytest

import datagen
import golden_util
import refdump


def _clusters(cl):
    return sorted((sorted(c.tolist()), int(c[0])) for c in cl if len(c))


def _want(ref):
    so = ref["spectral_off"].astype(np.int64)
    return [(ref["spectral_member"][so[i]:so[i + 1]].tolist(), int(ref["spectral_first"][i])) for i in range(len(so) - 1)]


@pytest.mark.parametrize("seed,genome,read_len,k", [(7, 20000, 150, 19), (9, 40000, 200, 15), (13, 95000, 410, 19), (12, 90000, 390, 19)])
def test_spectral_clustering_matches_the_reference(ref_driver, tmp_path, seed, genome, read_len, k):
    import hga_b200
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=genome, divergence=0.03 if k == 19 else 0.02, k=k, read_len=read_len, coverage=25 if seed != 7 else 30,
                                          seed=seed, error_rate=0.005 if k == 19 else 0.01, fmt="fastq" if seed == 7 else "fasta")
    ref = refdump.run_ref(ref_driver, paths, kp, enrich=20, full=True)
    assert ref["scaffold_components"] > 2 and ref["strong_tail_connections"] > 0
    m = ref["tconn_score"] > 5                                     # strong_core_connections, ReadClusteringEngine.cpp:770
    got = hga_b200.capi.spectral_clustering(ref["tconn_x"][m], ref["tconn_y"][m], ref["tconn_score"][m], 16)
    # members of every cluster AND its element [0] (the component that survives merge_components, :366)
    assert _clusters(got) == _want(ref)
    # (seed 12 has a disconnected tail-connection graph: a repeated leading eigenvalue. What comes out then depends on which basis
    # of the degenerate eigenspace the eigen-solver returns - with a Jacobi solver on both sides some rows of the leading
    # eigenvectors were exactly zero, the reference's quality became NaN and NO cluster was formed; with the QL solver both sides
    # form 16 clusters. Product and reference agree either way because they share the solver.)


@pytest.mark.parametrize("name", ["spectral_a", "spectral_b"])
def test_spectral_clustering_golden(name):
    import hga_b200
    z = np.load(os.path.join(golden_util.GOLDEN, name + ".npz"))
    got = hga_b200.capi.spectral_clustering(z["conn_x"], z["conn_y"], z["conn_score"], int(z["dims"]))
    off = z["cluster_off"].astype(np.int64)
    want = [(z["cluster_member"][off[i]:off[i + 1]].tolist(), int(z["cluster_first"][i])) for i in range(len(off) - 1)]
    assert _clusters(got) == want


def test_spectral_clustering_edge_cases():
    import hga_b200
    assert hga_b200.capi.spectral_clustering(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint64), 16) == []
    # all scores equal (e.g. a single connection): the reference scales by (score - min) / (max - min) = 0 / 0 (:670-674), every
    # affinity is NaN, no quality comparison succeeds and no cluster is formed; the product keeps that behaviour
    cl = hga_b200.capi.spectral_clustering([5], [9], [7], 16)
    assert [v for c in cl for v in c] == []
    cl = hga_b200.capi.spectral_clustering([5, 5, 9], [9, 11, 11], [7, 7, 7], 16)
    assert [v for c in cl for v in c] == []
    # two different scores on a path of three components: everything is assigned exactly once
    cl = hga_b200.capi.spectral_clustering([5, 9], [9, 11], [7, 30], 16)
    assert sorted(int(v) for c in cl for v in c) == [5, 9, 11]


@pytest.mark.parametrize("n", [1, 2, 3, 8, 50, 300])
def test_eigen_solver_against_numpy(n):
    """the symmetric eigen-solver the spectral stage uses (Householder + implicit QL) against numpy.linalg.eigh: eigenvalues,
    residual and orthonormality; n = 8 is block diagonal (exact zeros off the diagonal blocks)"""
    import hga_b200
    rng = np.random.default_rng(n)
    b = rng.standard_normal((n, n))
    a = (b + b.T) / 2
    if n == 8:
        a[:4, 4:] = 0
        a[4:, :4] = 0
    w, v = hga_b200.capi.host_sym_eigen(a)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - np.linalg.eigvalsh(a)).max() < 1e-11
    assert np.abs(a @ v - v * w).max() < 1e-11 and np.abs(v.T @ v - np.eye(n)).max() < 1e-11
