#!/usr/bin/env python
"""End-to-end run of the C++ `categorization` executable on generated FASTA files, timed by wall clock, next to the REAL reference
(oracle/_ref/ref_driver --enrich, all host threads) on the same files; the exported components are compared with the reference's
final components (read sets AND surviving component ids = file names). Test / measurement infrastructure.

    python scripts/cli_e2e.py --genome-mbp 4 --out gpurun_out/cli_e2e.json
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402

EXE = os.path.join(ROOT, "hybrid-genome-assembler_b200", "categorization")
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def write_kmers_fast(path, values, k):
    v = np.asarray(values, dtype=np.uint64)
    arr = np.empty((v.shape[0], k + 1), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(k):
        arr[:, i] = lut[((v >> np.uint64(2 * (k - 1 - i))) & np.uint64(3)).astype(np.int64)]
    arr[:, k] = 10
    with open(path, "wb") as f:
        f.write(arr.tobytes())


def write_fasta_fast(path, reads, prefix):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">%s%d\n" % (prefix, i))
            f.write(lut[r].tobytes())
            f.write(b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mbp", type=float, default=4.0)
    ap.add_argument("--coverage", type=float, default=50.0)
    ap.add_argument("--mean-len", type=int, default=10000)
    ap.add_argument("--divergence", type=float, default=0.01)
    ap.add_argument("--error", type=float, default=0.05)
    ap.add_argument("--k", type=int, default=19)
    ap.add_argument("--seed", type=int, default=77)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    G = int(args.genome_mbp * 1e6)
    res = {"genome_mbp": args.genome_mbp, "coverage": args.coverage, "k": args.k}
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        a = datagen.random_genome(G, args.seed)
        b = datagen.mutate(a, args.divergence, args.seed + 1)
        n_per = max(1, int(args.coverage * G / args.mean_len))
        paths, bases = [], 0
        for i, hp in enumerate((a, b)):
            reads = datagen.sample_reads(hp, n_per, args.mean_len, args.seed + 10 + i, error_rate=args.error, length_sigma=0.5, min_len=100,
                                         max_len=min(60000, G))
            bases += sum(len(r) for r in reads)
            p = os.path.join(d, f"hap{i}.fa")
            write_fasta_fast(p, reads, b"h%d_" % i)
            paths.append(p)
            del reads
        sdk = datagen.discriminative_kmers([a, b], args.k)
        kp = os.path.join(d, f"{args.k}-mers.txt")
        write_kmers_fast(kp, np.random.default_rng(args.seed + 5).permutation(sdk), args.k)
        res.update(bases=int(bases), reads=2 * n_per, kmers=int(sdk.shape[0]), generate_s=time.perf_counter() - t0)

        out = os.path.join(d, "clusters")
        t0 = time.perf_counter()
        r = subprocess.run([EXE] + paths + ["--kmers", kp, "-o", out], capture_output=True, text=True)
        res["cli_wall_s"] = time.perf_counter() - t0
        res["cli_rc"] = r.returncode
        res["cli_stdout_stages"] = [l for l in r.stdout.split("\n") if " took " in l or l.startswith("Exported")]
        res["cli_stderr_tail"] = r.stderr.strip().split("\n")[-12:]
        if r.returncode != 0:
            print(json.dumps(res)); return 1
        got = {}
        for f in os.listdir(out):
            with open(os.path.join(out, f), "rb") as fh:
                got[f] = [l[1:].decode() for l in fh.read().split(b"\n")[0::2] if l]
        res["cli_components"] = len(got)
        res["cli_gbases_per_s_wall"] = bases / res["cli_wall_s"] / 1e9

        if not args.no_reference and os.path.exists(DRIVER):
            rout = os.path.join(d, "ref")
            os.makedirs(rout)
            t0 = time.perf_counter()
            subprocess.run([DRIVER, "run", "--kmers", kp, "--out", rout, "--threads", str(os.cpu_count() or 1), "--enrich", "20", "--full"] + paths, check=True,
                           stdout=subprocess.DEVNULL)
            res["reference_wall_s"] = time.perf_counter() - t0
            res["reference_threads"] = os.cpu_count() or 1
            meta = dict(l.strip().split("=") for l in open(os.path.join(rout, "meta.txt")))
            res["reference_stage_ms"] = {k_: float(v) for k_, v in meta.items() if k_.endswith("_ms")}
            fid = np.fromfile(os.path.join(rout, "final_id.u32"), dtype=np.uint32)
            fo = np.fromfile(os.path.join(rout, "final_off.u64"), dtype=np.uint64).astype(np.int64)
            fr = np.fromfile(os.path.join(rout, "final_read.u32"), dtype=np.uint32).astype(np.int64)
            # read id -> header: ids run over the files in order
            hdr = [f"h0_{i}" for i in range(n_per)] + [f"h1_{i}" for i in range(n_per)]
            want = {f"#{int(c)}.fa": [hdr[j - 1] for j in fr[fo[i]:fo[i + 1]]] for i, c in enumerate(fid)}
            res["reference_components"] = len(want)
            res["final_components_identical"] = bool(want == got)
            res["speedup_wall"] = res["reference_wall_s"] / res["cli_wall_s"]
    line = json.dumps(res)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")
    return 0 if res.get("final_components_identical", True) else 1


if __name__ == "__main__":
    sys.exit(main())
