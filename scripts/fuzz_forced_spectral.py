"""Differential fuzz of the --spectral path (run_clustering :739-746): connections with score >= 5 (C oracle) -> the product's
host stage (engine.forced_spectral_components -> hga_spectral_clustering) against the real reference (ref_driver --force-spectral).
Test infrastructure; needs the driver (build container).

    python scripts/fuzz_forced_spectral.py <seed> <cases>"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import datagen  # noqa: E402
import hga_b200  # noqa: E402
import oracle_lib  # noqa: E402
import refdump  # noqa: E402

drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
orc = oracle_lib.load()
rng = np.random.default_rng(int(sys.argv[1]))
for it in range(int(sys.argv[2])):
    long_ = rng.random() < 0.5
    kw = dict(genome_size=int(rng.integers(3000, 20000)), divergence=float(rng.choice([0.01, 0.02, 0.03])), k=int(rng.choice([15, 17, 19, 21])),
              read_len=int(rng.integers(800, 2500)) if long_ else int(rng.integers(120, 300)), coverage=int(rng.integers(6, 12)),
              seed=int(rng.integers(1, 10000)), error_rate=float(rng.choice([0.005, 0.02, 0.05])))
    if long_:
        kw["length_sigma"] = float(rng.choice([0.3, 0.5]))
    ms = int(rng.choice([2, 5, 20]))
    d = os.path.join(tempfile.gettempdir(), f"hga_fuzz_fs_{it}"); os.makedirs(d, exist_ok=True)
    paths, kp = datagen.make_diploid_case(d, **kw)
    ref = refdump.run_ref(drv, paths, kp, min_size=ms, force_spectral=True)
    rc, reads = orc.load_reads(paths); kmers, k = orc.load_kmers(kp)
    row_off, kid, pos = orc.scan(reads["seq"], reads["seq_off"], k, kmers)
    inv_off, inv_read = orc.index(row_off, kid, len(kmers))
    cx, cy, cs = orc.canonical_sort(*orc.connections(row_off, kid, inv_off, inv_read, min_score=5))
    got = hga_b200.engine.forced_spectral_components(cx, cy, cs, 16, ms)
    fo = ref["final_off"].astype(np.int64)
    want = [(int(ref["final_id"][i]), ref["final_read"][fo[i]:fo[i + 1]].tolist()) for i in range(len(fo) - 1)]
    ok = [(f, m.tolist()) for f, m in got.items()] == want
    print(it, "reads", ref["n_reads"], "connections", len(cx), "clusters", ref.get("spectral_clusters"), "final", len(want), "equal", ok, "" if ok else (kw, ms))
