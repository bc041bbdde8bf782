"""CPU: the data-parallel form of the SECOND merge_components (merge of the spectral clusters, run_clustering :774) that
hga_enrich_full runs on the GPU (enr_merge2_keys_kernel + enr_purge2_kernel, csrc/hga_enrich.cu), restated in numpy and checked
against the real reference's state after that merge (fixtures full_*.npz from ref_driver --enrich 20 --full; live against the
driver when it is built). The state BEFORE the second merge comes from the C oracle's engine, which the other tests pin against the
reference. What this pins is the RULE (removal bounds, truncation, first-copy removal, unions); the kernels that apply it run in
tests/test_zy_gpu_tail_block.py."""
import numpy as np
import pytest

import golden_util
import oracle_lib


def second_merge_rule(core_id, unions, purged_off, purged_read, clusters):
    """core_id[c]: survivor read id of core c (ascending); unions[c]: sorted unique k-mer ids of core c; purged CSR by k-mer id;
    clusters: lists of survivor ids, element [0] survives. Returns (new core ids, new unions, new purged CSR)."""
    index_of = {int(s): i for i, s in enumerate(core_id)}
    n_cores = len(core_id)
    into = np.arange(n_cores)
    multi = np.zeros(n_cores, dtype=bool)
    for cl in clusters:
        if len(cl) < 2:                                   # ReadClusteringEngine.cpp:360-365
            continue
        for s in cl:
            into[index_of[int(s)]] = index_of[int(cl[0])]
            multi[index_of[int(s)]] = True
    n_kmers = len(purged_off) - 1
    bound = np.zeros(n_kmers, dtype=np.int64)             # R2(k): largest id on the k-mer's removal list + 1 (:385-389)
    for c in range(n_cores):
        if multi[c]:
            np.maximum.at(bound, unions[c], max(int(core_id[c]), int(core_id[into[c]])) + 1)
    live = [c for c in range(n_cores) if into[c] == c]
    new_unions = [np.unique(np.concatenate([unions[c] for c in range(n_cores) if into[c] == s])) for s in live]
    removed_once = set(int(core_id[c]) for c in range(n_cores) if multi[c])
    off = [0]
    rows = []
    purged_off = purged_off.astype(np.int64)
    for k in range(n_kmers):
        lst = purged_read[purged_off[k]:purged_off[k + 1]].tolist()
        if bound[k] == 0:
            rows.extend(lst)
        else:
            prev = None
            for e in lst:                                 # the two-pointer walk stops with the removal list (:405-416)
                if e + 1 >= bound[k]:
                    break
                if e not in removed_once or prev == e:
                    rows.append(e)
                prev = e
        off.append(len(rows))
    return core_id[live], new_unions, np.array(off, dtype=np.uint64), np.array(rows, dtype=np.uint32)


def _state_after_first_merge(oracle, bases, seq_off, k, kmers, min_size, fraction=0.15):
    res = oracle.run(bases, seq_off, k, kmers, min_size=min_size, fraction=fraction)
    eng = oracle_lib.Engine(oracle, res["row_off"], res["hit_kid"], len(kmers), res["inv_off"], res["inv_read"])
    try:
        eng.merge(res["comp"][0], res["comp"][1])
        cores = np.sort(eng.ids(min_size))
        unions = [np.sort(eng.component_kmers(c)).astype(np.int64) for c in cores]
        po, pr = eng.index()
    finally:
        eng.close()
    return cores, unions, po, pr


def _clusters(ref):
    so = ref["spectral_off"].astype(np.int64)
    out = []
    for i in range(len(so) - 1):
        first = int(ref["spectral_first"][i])
        out.append([first] + [int(v) for v in ref["spectral_member"][so[i]:so[i + 1]] if int(v) != first])
    return out


def _check(ref, kmers, got):
    ids, unions, off, rows = got
    assert np.array_equal(ids, ref["core_id"])
    ko = ref["core_kmer_off"].astype(np.int64)
    for i, u in enumerate(unions):
        assert np.array_equal(np.sort(kmers[u]), ref["core_kmer"][ko[i]:ko[i + 1]])
    assert np.array_equal(off, ref["purged_off"]) and np.array_equal(rows, ref["purged_read"])


@pytest.mark.parametrize("name", golden_util.FULL_CASES)
def test_second_merge_rule_golden(oracle, name):
    c = golden_util.load_case(name)
    ref = c["ref"]
    cores, unions, po, pr = _state_after_first_merge(oracle, c["bases"], c["seq_off"], c["k"], c["kmers"], c["min_size"], c["fraction"])
    assert len(cores) == ref["merged_scaffolds"] > ref["cores"]
    _check(ref, c["kmers"], second_merge_rule(cores, unions, po, pr, _clusters(ref)))


def test_second_merge_rule_live(oracle, ref_driver, tmp_path):
    """a case where one cluster swallows 115 of 121 scaffold components"""
    import datagen
    import refdump
    paths, kp = datagen.make_diploid_case(str(tmp_path), genome_size=100000, divergence=0.03, k=19, read_len=400, coverage=25, seed=10, error_rate=0.01)
    ref = refdump.run_ref(ref_driver, paths, kp, enrich=20, full=True, min_size=30)
    rc, reads = oracle.load_reads(paths)
    kmers, k = oracle.load_kmers(kp)
    cores, unions, po, pr = _state_after_first_merge(oracle, reads["seq"], reads["seq_off"], k, kmers, 30)
    assert len(cores) == ref["merged_scaffolds"] and ref["cores"] < 10
    _check(ref, kmers, second_merge_rule(cores, unions, po, pr, _clusters(ref)))
