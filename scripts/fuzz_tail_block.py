"""Differential fuzz of the tail / spectral block's host stages and of the second-merge rule against the real reference
(oracle/_ref/ref_driver --enrich 20 --full). Test infrastructure; needs the driver (build container).

    python scripts/fuzz_tail_block.py <seed> <cases>

Per case: random genome size, k, read length, coverage, error rate -> tail connections (hga_host_tail_connections), clusters
(hga_spectral_clustering) and the state after the second merge (tests/test_second_merge_rule.py) compared with the reference."""
import os
import sys
import tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, datagen, refdump, oracle_lib, hga_b200
import test_tails_cpu as T, test_second_merge_rule as R
drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
orc = oracle_lib.load()
rng = np.random.default_rng(int(sys.argv[1]))
for it in range(int(sys.argv[2])):
    long_ = rng.random() < 0.6
    kw = dict(genome_size=int(rng.integers(30000, 120000)), divergence=float(rng.choice([0.01, 0.02, 0.03])), k=int(rng.choice([15, 17, 19, 21])),
              read_len=int(rng.integers(1000, 3500)) if long_ else int(rng.integers(150, 500)), coverage=int(rng.integers(10, 16)) if long_ else int(rng.integers(20, 30)),
              seed=int(rng.integers(1, 10000)), error_rate=float(rng.choice([0.005, 0.02, 0.05])))
    if long_: kw["length_sigma"] = float(rng.choice([0.3, 0.5]))
    ms = 5 if long_ else 30
    d = os.path.join(tempfile.gettempdir(), f"hga_fuzz_tail_{it}"); os.makedirs(d, exist_ok=True)
    paths, kp = datagen.make_diploid_case(d, **kw)
    try:
        ref = refdump.run_ref(drv, paths, kp, enrich=20, full=True, min_size=ms)
    except Exception as e:
        print(it, kw, "REF FAILED", e); continue
    if ref["scaffold_components"] <= 2 or "tconn_x" not in ref:
        print(it, "skip: scaffolds", ref["scaffold_components"]); continue
    rc, reads = orc.load_reads(paths); kmers, k = orc.load_kmers(kp)
    a = T._inputs(orc, reads["seq"], reads["seq_off"], k, kmers, ms)
    x, y, s = hga_b200.capi.host_tail_connections(amplification_min_score=40, **a)
    ok_t = np.array_equal(x, ref["tconn_x"]) and np.array_equal(y, ref["tconn_y"]) and np.array_equal(s, ref["tconn_score"])
    ok_c = ok_r = None
    if "spectral_off" in ref:
        got = hga_b200.capi.spectral_clustering(x[s > 5], y[s > 5], s[s > 5], 16)
        so = ref["spectral_off"].astype(np.int64)
        want = [(ref["spectral_member"][so[i]:so[i + 1]].tolist(), int(ref["spectral_first"][i])) for i in range(len(so) - 1)]
        ok_c = sorted((sorted(c.tolist()), int(c[0])) for c in got if len(c)) == want
        cores, unions, po, pr = R._state_after_first_merge(orc, reads["seq"], reads["seq_off"], k, kmers, ms)
        try:
            R._check(ref, kmers, R.second_merge_rule(cores, unions, po, pr, R._clusters(ref))); ok_r = True
        except AssertionError:
            ok_r = False
    ok_f = None
    if ok_t and ok_c is not False:
        # the whole chain to the final components (second merge and enrichment on the oracle's engine state)
        clusters = [c for c in hga_b200.capi.spectral_clustering(x[s > 5], y[s > 5], s[s > 5], 16) if len(c)] if len(x) and (s > 5).any() else []
        res = orc.run(reads["seq"], reads["seq_off"], k, kmers, min_size=ms)
        eng = oracle_lib.Engine(orc, res["row_off"], res["hit_kid"], len(kmers), res["inv_off"], res["inv_read"])
        try:
            eng.merge(res["comp"][0], res["comp"][1])
            if clusters:
                eng.merge(np.cumsum([0] + [len(c) for c in clusters]).astype(np.uint64), np.concatenate(clusters))
            cores = eng.ids(ms)
            ex, ey, es = orc.canonical_sort(*eng.connections(cores, 20))
            eo, em, _, _, _ = orc.union_find(ex, ey, min_size=2, max_size=-1, restricted=cores)
            eng.merge(eo, em)
            got = sorted((int(np.sort(eng.component_reads(c))[0]), int(c), np.sort(eng.component_reads(c)).tolist()) for c in eng.ids(ms))
        finally:
            eng.close()
        fo = ref["final_off"].astype(np.int64)
        ok_f = got == [(int(ref["final_read"][fo[i]]), int(ref["final_id"][i]), ref["final_read"][fo[i]:fo[i + 1]].tolist()) for i in range(len(fo) - 1)]
    print(it, "final", ok_f, "scaffolds", ref["scaffold_components"], "tconn", len(x), "->cores", ref["cores"], "tails", ok_t, "clusters", ok_c, "rule", ok_r, "" if (ok_t and ok_c is not False and ok_r is not False and ok_f is not False) else kw)
