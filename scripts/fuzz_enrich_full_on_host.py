"""Differential fuzz of hga_enrich_full's OWN source, run on the host (tests/test_kernel_bodies_on_host.py: hga_enrich.cu compiled
against a stand-in CUDA runtime / CUB), against the real reference (oracle/_ref/ref_driver --enrich 20 --full) on random cases.
Test infrastructure; needs the driver (build container).

    python scripts/fuzz_enrich_full_on_host.py <seed> <cases>"""
import os
import pathlib
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import compare  # noqa: E402
import datagen  # noqa: E402
import oracle_lib  # noqa: E402
import refdump  # noqa: E402
import test_kernel_bodies_on_host as T  # noqa: E402


class _Factory:
    def mktemp(self, name):
        return pathlib.Path(tempfile.mkdtemp(prefix=name))


lib = T.host_enrich.__wrapped__(_Factory())
orc = oracle_lib.load()
drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
rng = np.random.default_rng(int(sys.argv[1]))
bad = 0
for it in range(int(sys.argv[2])):
    long_ = rng.random() < 0.6
    kw = dict(genome_size=int(rng.integers(30000, 100000)), divergence=float(rng.choice([0.01, 0.02, 0.03])), k=int(rng.choice([15, 17, 19, 21])),
              read_len=int(rng.integers(1000, 3500)) if long_ else int(rng.integers(150, 500)), coverage=int(rng.integers(10, 16)) if long_ else int(rng.integers(20, 30)),
              seed=int(rng.integers(1, 10000)), error_rate=float(rng.choice([0.005, 0.02, 0.05])))
    if long_:
        kw["length_sigma"] = float(rng.choice([0.3, 0.5]))
    ms = 5 if long_ else 30
    d = tempfile.mkdtemp()
    paths, kp = datagen.make_diploid_case(d, **kw)
    ref = refdump.run_ref(drv, paths, kp, enrich=20, full=True, min_size=ms)
    rc, reads = orc.load_reads(paths)
    kmers, k = orc.load_kmers(kp)
    c = dict(bases=reads["seq"], seq_off=reads["seq_off"], k=k, kmers=kmers, fraction=0.15, min_size=ms, enrich=20)
    e, t = T._emu_run(lib, orc, c, with_tail=True)
    try:
        if "tconn_x" in ref:
            assert np.array_equal(t["conn_x"], ref["tconn_x"]) and np.array_equal(t["conn_score"], ref["tconn_score"])
        compare.check_enrichment(ref, e, kmers)
        ok = True
    except AssertionError as ex:
        ok = False
        bad += 1
        print("  ", ex)
    print(it, "scaffolds", ref["scaffold_components"], "-> cores", ref["cores"], "final", ref["final_components"], "block ran", t["ran"], "equal", ok, "" if ok else kw)
print("differing cases:", bad)
