"""Multi-GPU parity check (one process per GPU, launched by torchrun; see tests/test_multi_gpu.py):
every rank scans its contiguous shard of the reads, the ranks exchange the inverted index over NCCL, and the union of
the ranks' results is compared with the oracle run on the whole input. Bit-exact: hits, owner-partitioned inverted index (every list complete on exactly one rank),
pair scores, cut (n, s*), selected edge set, components.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch
    import torch.distributed as dist
    import hga_b200
    from hga_b200 import parallel
    import datagen
    import oracle_lib
    import compare

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    for case, (k, gsize, n_reads, rlen, sigma, err, min_size) in enumerate([(19, 60000, 500, 1500, 0.5, 0.03, 5), (15, 20000, 3000, 150, 0.0, 0.005, 10)]):
        a = datagen.random_genome(gsize, 700 + case)
        b = datagen.mutate(a, 0.02, 800 + case)
        reads = datagen.sample_reads(a, n_reads, rlen, 11 + case, error_rate=err, length_sigma=sigma, max_len=gsize) + \
            datagen.sample_reads(b, n_reads, rlen, 21 + case, error_rate=err, length_sigma=sigma, max_len=gsize)
        reads.insert(7, np.zeros(0, dtype=np.uint8))            # an empty read inside a shard
        seqs = [datagen.to_ascii(r).encode() for r in reads]
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        kmers = datagen.discriminative_kmers([a, b], k)
        bounds = parallel.shard_bounds(lens, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        bases = b"".join(seqs[lo:hi])
        off = np.zeros(hi - lo + 1, dtype=np.uint64)
        np.cumsum(lens[lo:hi], out=off[1:])

        h = hga_b200.Handle(kmers, k, device=local)
        uid = parallel.broadcast_unique_id(dist, rank, hga_b200.capi.comm_unique_id)
        h.comm_init(uid, rank, world, len(seqs))
        h.scan(bases, off, read_id_base=lo + 1)
        row_off, kid, pos = h.get_hits()
        h.build_index()
        inv_off, inv_read = h.get_index()
        h.pair_count(min_score=1)
        x, y, s, _ = h.get_pairs()
        h.select_edges(fraction=0.15)
        sel = h.get_selection()
        h.components(min_size=min_size)
        comp = h.get_components()
        m = h.metrics()
        # the stages after the scaffold union_find run on rank 0 after hga_comm_gather_root: the gathered handle must behave as a single-GPU
        # handle that scanned everything (hits, by-slot index, selection, components) and hga_enrich_full on it must give the single-GPU result
        h.comm_gather_root()
        if rank == 0:
            allb = b"".join(seqs)
            alloff = np.zeros(len(seqs) + 1, dtype=np.uint64)
            np.cumsum(lens, out=alloff[1:])
            h1 = hga_b200.Handle(kmers, k, device=local)
            h1.scan(allb, alloff, read_id_base=1)
            h1.build_index(); h1.pair_count(min_score=1); h1.select_edges(fraction=0.15); h1.components(min_size=min_size)
            for a1, a2 in zip(h.get_hits(), h1.get_hits()):
                assert np.array_equal(a1, a2), f"case {case}: gathered hits differ from the single-GPU hits"
            for a1, a2 in zip(h.get_index(), h1.get_index()):
                assert np.array_equal(a1, a2), f"case {case}: index rebuilt on rank 0 differs from the single-GPU index"
            s0, s1 = h.get_selection(), h1.get_selection()
            for key in ("x", "y", "score"):
                assert np.array_equal(s0[key], s1[key]), f"case {case}: gathered selection differs ({key})"
            for hh in (h, h1):
                hh.enrich_full(alloff, min_size=min_size, enrichment_min_score=5, tail_amplification_min_score=10, spectral_dims=8)
            e0, e1 = h.get_enrichment(), h1.get_enrichment()
            for key in ("core_id", "core_off", "core_read", "conn_x", "conn_y", "conn_score", "final_id", "final_off", "final_read", "assignment"):
                assert np.array_equal(e0[key], e1[key]), f"case {case}: enrichment on the gathered handle differs from the single-GPU run ({key})"
            print(f"gathered handle ok: case {case}, {len(e0['core_id'])} cores, {len(e0['final_id'])} final components")
            h1.close()
        h.close()

        mine = dict(row_off=row_off, kid=kid, pos=pos, x=x, y=y, s=s, sx=sel["x"], sy=sel["y"], ss=sel["score"], n_directed=sel["n_directed"],
                    cut=sel["cut_score"], label=comp["label"], comp_label=comp["comp_label"], comp_size=comp["comp_size"], inv_off=inv_off,
                    inv_read=inv_read, exchange_ms=m["exchange_ms"])
        box = [None] * world
        dist.all_gather_object(box, mine)
        if rank == 0:
            orc = oracle_lib.load()
            allb = b"".join(seqs)
            alloff = np.zeros(len(seqs) + 1, dtype=np.uint64)
            np.cumsum(lens, out=alloff[1:])
            ref = orc.run(allb, alloff, k, kmers, fraction=0.15, min_size=min_size)
            ro = ref["row_off"].astype(np.int64)
            for r, res in enumerate(box):
                a0, a1 = bounds[r], bounds[r + 1]
                assert np.array_equal(res["row_off"].astype(np.int64), ro[a0:a1 + 1] - ro[a0]), f"case {case}: row offsets of rank {r} differ"
                assert np.array_equal(res["kid"], ref["hit_kid"][ro[a0]:ro[a1]]) and np.array_equal(res["pos"], ref["hit_pos"][ro[a0]:ro[a1]]), f"hits of rank {r} differ"
                # the index is partitioned by k-mer owner (whole table buckets dealt round robin, an internal choice): a rank holds the COMPLETE
                # lists of the k-mers it owns and nothing of the others; below: every non-empty list lives on exactly one rank
                len_ref = np.diff(ref["inv_off"].astype(np.int64)); len_res = np.diff(res["inv_off"].astype(np.int64))
                owned = len_res > 0
                assert np.array_equal(len_res[owned], len_ref[owned]), f"case {case}: list lengths on rank {r} differ"
                assert np.array_equal(res["inv_read"], ref["inv_read"][np.repeat(owned, len_ref)]), f"case {case}: inverted lists of rank {r}'s k-mers differ"
            holders = sum((np.diff(res["inv_off"].astype(np.int64)) > 0).astype(np.int64) for res in box)
            assert np.array_equal(holders, (np.diff(ref["inv_off"].astype(np.int64)) > 0).astype(np.int64)), f"case {case}: a list is missing or held twice"
            ux, uy, us = compare.undirected(*ref["conn"])
            gx = np.concatenate([r["x"] for r in box]); gy = np.concatenate([r["y"] for r in box]); gs = np.concatenate([r["s"] for r in box])
            o = np.lexsort((gy, gx))
            assert np.array_equal(gx[o], ux) and np.array_equal(gy[o], uy) and np.array_equal(gs[o].astype(np.uint64), us), f"case {case}: pair scores differ"
            for r in box:
                assert r["n_directed"] == ref["cut_n"] and r["cut"] == ref["cut_score"], f"case {case}: cut differs"
            got = set()
            for r in box:
                got |= set(zip(r["sx"].tolist(), r["sy"].tolist(), r["ss"].tolist()))
            want = set()
            cx, cy, cs = ref["conn"]
            for p, q, w in zip(cx[:ref["cut_n"]].tolist(), cy[:ref["cut_n"]].tolist(), cs[:ref["cut_n"]].tolist()):
                want.add((min(p, q), max(p, q), w))
            assert got == want, f"case {case}: selected edge set differs ({len(got)} vs {len(want)})"
            assert sum(len(r["sx"]) for r in box) == len(want), "an edge was selected on two ranks"
            co, cm, _, _, _ = ref["comp"]
            wantc = compare.components_partition(co, cm)
            for r in box:
                label = r["label"]
                gotc = sorted(tuple(sorted(int(v) for v in (np.nonzero(label == c)[0] + 1))) for c in r["comp_label"])
                assert gotc == wantc, f"case {case}: components differ"
            print(f"multi-GPU parity ok: case {case}, {world} ranks, {len(seqs)} reads, {len(ux)} pairs, {len(want)} selected, {len(wantc)} components, "
                  f"exchange {max(r['exchange_ms'] for r in box):.3f} ms")
        dist.barrier()

    # a rank whose local phase fails before an exchange (HGA_FAULT test hook) must not leave the others waiting in NCCL: every rank leaves the stage with an
    # error (the failing one with its own, the others with "rank r failed before the exchange"), and the same handles' communicator is still usable afterwards
    a = datagen.random_genome(20000, 900)
    b = datagen.mutate(a, 0.02, 901)
    reads = datagen.sample_reads(a, 100, 1000, 31) + datagen.sample_reads(b, 100, 1000, 32)
    seqs = [datagen.to_ascii(r).encode() for r in reads]
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    kmers = datagen.discriminative_kmers([a, b], 19)
    bounds = parallel.shard_bounds(lens, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    off = np.zeros(hi - lo + 1, dtype=np.uint64)
    np.cumsum(lens[lo:hi], out=off[1:])
    h = hga_b200.Handle(kmers, 19, device=local)
    uid = parallel.broadcast_unique_id(dist, rank, hga_b200.capi.comm_unique_id)
    h.comm_init(uid, rank, world, len(seqs))
    for stage in ("index", "partials"):
        os.environ["HGA_FAULT"] = f"{stage}:{world - 1}"
        h.scan(b"".join(seqs[lo:hi]), off, read_id_base=lo + 1)
        failed = None
        try:
            h.build_index()
            h.pair_count(min_score=1)
        except hga_b200.capi.HgaError as e:
            failed = str(e)
        finally:
            del os.environ["HGA_FAULT"]
        assert failed is not None, f"rank {rank}: the injected {stage} fault went unnoticed"
        assert ("injected fault" in failed) == (rank == world - 1), (rank, failed)
    # ... and without the fault the same handles still work
    h.scan(b"".join(seqs[lo:hi]), off, read_id_base=lo + 1)
    h.build_index(); h.pair_count(min_score=1); h.select_edges(fraction=0.15); h.components(min_size=5)
    n_pairs = torch.tensor([h.get_pairs()[0].shape[0]], device=torch.device("cuda", local))
    dist.all_reduce(n_pairs)
    assert int(n_pairs.item()) > 0
    h.close()
    if rank == 0:
        print(f"fault injection ok: {world} ranks left the index and the partial-pair exchange together")

    # the Python mirror of the engine with dist = torch.distributed: every rank runs run_clustering, all get rank 0's final components,
    # and they are the single-GPU engine's
    import tempfile
    from hga_b200 import engine
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, f"reads_rank{rank}.fa")
        datagen.write_fasta(fa, reads, prefix="r")
        cfg = engine.ReadClusteringConfig(scaffold_component_min_size=5, enrichment_connections_min_score=5, tail_amplification_min_score=10, spectral_dims=8)
        eng = engine.ReadClusteringEngine(engine.SequenceRecords([fa]), cfg, device=local, dist=dist)
        ids = eng.run_clustering(kmers, 19)
        comps = {int(k_): np.asarray(v) for k_, v in eng.final_components.items()}
        eng.close()
        if rank == 0:
            one = engine.ReadClusteringEngine(engine.SequenceRecords([fa]), cfg, device=local)
            ids1 = one.run_clustering(kmers, 19)
            assert sorted(ids) == sorted(ids1) and len(ids1) > 0, (ids, ids1)
            for fid in ids1:
                assert np.array_equal(comps[fid], one.final_components[fid]), f"final component {fid} differs between {world} GPUs and one"
            assert np.array_equal(eng.assignment, one.assignment)
            one.close()
            print(f"engine mirror ok: {world} ranks, {len(ids1)} final components")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
