#!/usr/bin/env python
"""bench.py — categorization hot path on B200: Gbases/s scanned + read-pairs/s counted (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--genome-mbp G] ...

A "step" is one full pass of the hot path (scan -> inverted index -> pair count -> edge selection -> components)
over one batch of synthetic reads. N=1 workload = BASELINE.json configs[3] ("config 4"): synthetic 100 Mbp diploid,
1 % divergence, 50x long reads (~10 kb mean, ~10 Gbp), k = 19, discriminative set = canonical 19-mers present in
exactly one haplotype. Synthetic data is generated ON the GPU with torch (plumbing only); every kernel inside the
timed region belongs to libhga_b200.so and is reached through its C-ABI.

Output: ONE JSON line on rank 0 (see the keys at the bottom of main()).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 19
METRIC = "categorization_hot_path_gbases_per_s"

# BASELINE.json configs (SURVEY.md 8d "Synthetic inputs"); 4 is the one the metric is quoted on and the driver's default
CONFIGS = {
    1: dict(name="config1: 500 kbp diploid, 3% divergence, 150 bp reads 30x per haplotype, 0.5% substitution errors (the reference's own CPU-runnable case)",
            genome_mbp=0.5, divergence=0.03, coverage=30.0, mean_len=150.0, len_sigma=0.0, error=0.005, ks=[19], haplotypes=2),
    2: dict(name="config2: E. coli sized 2.6 Mbp diploid, 2% divergence, 75x long reads (lognormal, mean 7.8 kb), 10% substitution errors",
            genome_mbp=2.6, divergence=0.02, coverage=75.0, mean_len=7800.0, len_sigma=0.5, error=0.10, ks=[19], haplotypes=2),
    3: dict(name="config3: k sweep on the config-2 reads",
            genome_mbp=2.6, divergence=0.02, coverage=75.0, mean_len=7800.0, len_sigma=0.5, error=0.10, ks=[15, 17, 21], haplotypes=2),
    4: dict(name="config4", genome_mbp=100.0, divergence=0.01, coverage=50.0, mean_len=10000.0, len_sigma=0.5, error=0.05, ks=[19], haplotypes=2),
    5: dict(name="config5: tetraploid, 4 read files (haplotypes) of 50 Mbp, 40x each, dense SDK = k-mers absent from at least one haplotype",
            genome_mbp=50.0, divergence=0.01, coverage=40.0, mean_len=10000.0, len_sigma=0.5, error=0.05, ks=[19], haplotypes=4),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-mbp", type=float, default=100.0, help="haplotype size in Mbp (config 4: 100)")
    ap.add_argument("--coverage", type=float, default=50.0)
    ap.add_argument("--mean-len", type=float, default=10000.0)
    ap.add_argument("--divergence", type=float, default=0.01)
    ap.add_argument("--error", type=float, default=0.05, help="per-base substitution error of the reads")
    ap.add_argument("--seed", type=int, default=4000)
    ap.add_argument("--cpu-sample-mbp", type=float, default=0.25, help="haplotype size of the CPU-baseline sample (same generator)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-enrich", action="store_true")
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS), help="BASELINE.json config (1-based); 4 = the headline workload")
    ap.add_argument("--k", type=int, default=0, help="k-mer length (default: the config's)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the comparison with a single-GPU run of the same data on rank 0")
    a = ap.parse_args()
    a.len_sigma, a.haplotypes, a.ks = 0.5, 2, [a.k or K]
    if a.config != 4:
        c = CONFIGS[a.config]
        for key in ("genome_mbp", "divergence", "coverage", "mean_len", "len_sigma", "error", "haplotypes"):
            setattr(a, key, c[key])
        a.ks = [a.k] if a.k else c["ks"]
    return a


# ----------------------------------------------------------------------------------------------------------
# synthetic data on the GPU (torch = plumbing)
# ----------------------------------------------------------------------------------------------------------
def make_haplotypes(torch, dev, genome_size, divergence, seed, n_hap=2):
    """base haplotype + (n_hap - 1) independently mutated copies"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    a = torch.randint(0, 4, (genome_size,), dtype=torch.uint8, device=dev, generator=g)
    haps = [a]
    for _ in range(n_hap - 1):
        mut = (torch.rand(genome_size, device=dev, generator=g) < divergence).to(torch.uint8)
        shift = torch.randint(1, 4, (genome_size,), dtype=torch.uint8, device=dev, generator=g)
        haps.append((a + mut * shift) % 4)        # scripts/read_generator.py:158-162
    return tuple(haps)


def canonical_kmers(torch, codes, k):
    n = codes.shape[0] - k + 1
    fwd = torch.zeros(n, dtype=torch.int64, device=codes.device)
    rev = torch.zeros(n, dtype=torch.int64, device=codes.device)
    for j in range(k):
        c = codes[j:j + n].to(torch.int64)
        fwd |= c << (2 * (k - 1 - j))
        rev |= (3 - c) << (2 * j)
    return torch.minimum(fwd, rev)


def discriminative_set(torch, haps, k):
    """two haplotypes: canonical k-mers present in exactly one of them; more: k-mers absent from at least one (the dense set of
    config 5). Sorted, int64 on the GPU."""
    uniq = [torch.unique(canonical_kmers(torch, h, k)) for h in haps]
    allk, counts = torch.unique(torch.cat(uniq), return_counts=True)
    del uniq
    return allk[counts == 1] if len(haps) == 2 else allk[counts < len(haps)]


def read_plan(torch, dev, genome_size, n_reads_per_hap, mean_len, seed, sigma=0.5, n_hap=2):
    """lengths / starts / strand of every read of all haplotypes (identical on every rank)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 7)
    n = n_hap * n_reads_per_hap
    mu = torch.log(torch.tensor(mean_len)) - 0.5 * sigma * sigma
    lens = torch.exp(mu + sigma * torch.randn(n, device=dev, generator=g)).to(torch.int64).clamp_(100, min(60000, genome_size))
    if sigma == 0:
        lens = torch.full_like(lens, int(mean_len))
    starts = (torch.rand(n, device=dev, generator=g, dtype=torch.float64) * (genome_size - lens + 1).to(torch.float64)).to(torch.int64)
    flip = torch.rand(n, device=dev, generator=g) < 0.5
    hap = torch.div(torch.arange(n, device=dev), n_reads_per_hap, rounding_mode="floor").to(torch.int64)
    return lens, starts, flip, hap


def _mix64(torch, x):
    """murmur3 finaliser on int64 tensors (wrapping arithmetic, logical shifts emulated)"""
    x = x ^ ((x >> 33) & 0x7FFFFFFF)
    x = x * -49064778989728563          # 0xff51afd7ed558ccd
    x = x ^ ((x >> 33) & 0x7FFFFFFF)
    x = x * -4265267296055464877        # 0xc4ceb9fe1a85ec53
    return x ^ ((x >> 33) & 0x7FFFFFFF)


def synth_reads(torch, dev, haps, lens, starts, flip, hap, lo, hi, error, seed, base0=0):
    """ASCII bytes of reads [lo, hi) back to back + their offsets (uint64-compatible int64). Substitution errors are a
    function of the GLOBAL base index (base0 = index of the shard's first base), so the data does not depend on how the
    reads are sharded over ranks."""
    L = lens[lo:hi]
    off = torch.zeros(hi - lo + 1, dtype=torch.int64, device=dev)
    torch.cumsum(L, 0, out=off[1:])
    total = int(off[-1].item())
    out = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    genome = torch.stack(haps)          # [n_hap, G]
    G = genome.shape[1]
    chunk_bytes = 1 << 27
    r0 = 0
    n = hi - lo
    off_cpu = off.cpu()
    while r0 < n:
        r1 = int(torch.searchsorted(off_cpu, off_cpu[r0] + chunk_bytes, right=True).item()) - 1
        r1 = max(r1, r0 + 1)
        r1 = min(r1, n)
        b0, b1 = int(off_cpu[r0]), int(off_cpu[r1])
        rid = torch.repeat_interleave(torch.arange(r0, r1, device=dev), L[r0:r1])
        j = torch.arange(b0, b1, device=dev) - off[rid]
        gr = rid + lo
        fl = flip[gr]
        src = torch.where(fl, starts[gr] + lens[gr] - 1 - j, starts[gr] + j)
        code = genome.view(-1)[hap[gr] * G + src]
        code = torch.where(fl, 3 - code, code)
        if error > 0:
            hsh = _mix64(torch, torch.arange(b0, b1, device=dev, dtype=torch.int64) + (base0 + seed * 1000003))
            e = (hsh & 0xFFFFFF).to(torch.float32) < error * float(1 << 24)
            sh = (((hsh >> 24) & 0xFFFF) % 3 + 1).to(torch.uint8)
            code = torch.where(e, (code + sh) % 4, code)
            del hsh, e, sh
        out[b0:b1] = lut[code.to(torch.int64)]
        del rid, j, gr, fl, src, code
        r0 = r1
    return out, off, total


# ----------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None

    def _run(self):
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.samples.append([c.strip() for c in r.stdout.strip().split("\n")[0].split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=10)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for nm, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU reference arm (the unmodified reference compiled into oracle/_ref/ref_driver)
# ----------------------------------------------------------------------------------------------------------
def cpu_sample_data(args, k):
    """the bounded sample both arms can run: the same generator scaled to args.cpu_sample_mbp per haplotype (numpy, tests/datagen.py)"""
    import datagen
    gsize = int(min(args.cpu_sample_mbp, args.genome_mbp) * 1e6)
    n_per_hap = max(1, int(args.coverage * gsize / args.mean_len))
    base = datagen.random_genome(gsize, args.seed)
    haps = [base] + [datagen.mutate(base, args.divergence, args.seed + 1 + i) for i in range(args.haplotypes - 1)]
    reads_per_hap = [datagen.sample_reads(hp, n_per_hap, int(args.mean_len), args.seed + 10 + i, error_rate=args.error, length_sigma=args.len_sigma,
                                          min_len=min(100, int(args.mean_len)), max_len=min(60000, gsize)) for i, hp in enumerate(haps)]
    if args.haplotypes == 2:
        sdk = datagen.discriminative_kmers(haps, k)
    else:
        import numpy as np
        uniq = [np.unique(datagen.canonical_kmers(hp, k)) for hp in haps]
        allk, cnt = np.unique(np.concatenate(uniq), return_counts=True)
        sdk = allk[cnt < len(haps)]
    return haps, reads_per_hap, sdk


def run_cpu_reference(args, k, steps=3, warmup=0, data=None):
    import datagen
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        return None
    cores = os.cpu_count() or 1
    haps, reads_per_hap, sdk = data if data is not None else cpu_sample_data(args, k)
    with tempfile.TemporaryDirectory() as d:
        paths = []
        total = 0
        for i, reads in enumerate(reads_per_hap):
            total += sum(len(r) for r in reads)
            p = os.path.join(d, f"hap{i}.fa")
            datagen.write_fasta(p, reads, prefix=f"h{i}_")
            paths.append(p)
        kp = os.path.join(d, f"{k}-mers.txt")
        datagen.write_kmers(kp, sdk, k)
        best = None
        for it in range(warmup + steps):
            out = os.path.join(d, f"out{it}")
            os.makedirs(out)
            t0 = time.perf_counter()
            subprocess.run([driver, "run", "--kmers", kp, "--out", out, "--threads", str(cores), "--no-dump"] + paths, check=True,
                           stdout=subprocess.DEVNULL)
            wall = time.perf_counter() - t0
            meta = {}
            with open(os.path.join(out, "meta.txt")) as f:
                for line in f:
                    kk, v = line.strip().split("=")
                    meta[kk] = float(v)
            # the reference's own stage timers; the driver's canonical re-sort (test harness work, not the reference's) is excluded
            hot_ms = meta["index_ms"] + meta["connections_ms"] + meta["union_find_ms"]
            rec = dict(hot_ms=hot_ms, index_ms=meta["index_ms"], connections_ms=meta["connections_ms"], union_find_ms=meta["union_find_ms"],
                       wall_s=wall, bases=total, pairs=meta["directed_connections"] / 2, kmers=int(meta["n_kmers"]))
            if it >= warmup and (best is None or rec["hot_ms"] < best["hot_ms"]):
                best = rec
    best["cores"] = cores
    best["sample"] = (f"same generator scaled to a {min(args.cpu_sample_mbp, args.genome_mbp):g} Mbp x {args.haplotypes} haplotypes ({best['bases'] / 1e6:.1f} Mbases of reads, "
                      f"{best['kmers']} {k}-mers), reference stage timers index + connections + union-find, --threads {cores}, best of {steps}")
    return best


def reference_arm(args, rank):
    if rank != 0:
        return
    res = run_cpu_reference(args, args.ks[0], steps=max(1, args.steps), warmup=0)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (reference sources absent at build time)"}))
        return
    value = res["bases"] / (res["hot_ms"] * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["hot_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args, args.ks[0]), "cpu_sample": res["sample"]},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                         "scan_gbases_per_s": res["bases"] / (res["index_ms"] * 1e-3) / 1e9,
                         "pairs_per_s": res["pairs"] / (res["connections_ms"] * 1e-3) if res["connections_ms"] > 0 else None},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args, k):
    if args.config != 4:
        return f"{CONFIGS[args.config]['name']}, k={k}, synthetic (seed {args.seed})"
    return (f"config4: synthetic {args.genome_mbp:g} Mbp diploid, {args.divergence * 100:g}% divergence, {args.coverage:g}x long reads "
            f"(lognormal, mean {args.mean_len:g} bp), {args.error * 100:g}% substitution errors, k={k}, SDK = canonical k-mers in exactly one haplotype")


def result_invariants(np, h, dist, world, dev, torch):
    """order-independent digest of a run: pair count, score sum, cut (n, s*), selected edges, component count and size multiset.
    With a communicator the pair / selection parts are summed over the ranks (every pair lives on exactly one rank)."""
    x, y, sc, _ = h.get_pairs()
    sel = h.get_selection()
    comp = h.get_components()
    part = np.array([x.shape[0], int(sc.astype(np.uint64).sum()), int((x.astype(np.uint64) * 1000003 + y.astype(np.uint64) * 7 + sc).sum() % (1 << 55)),
                     sel["x"].shape[0], int(sel["score"].astype(np.uint64).sum())], dtype=np.int64)
    if world > 1:
        t = torch.from_numpy(part).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        part = t.cpu().numpy()
        part[2] %= (1 << 55)
    sizes = np.sort(comp["comp_size"].astype(np.int64))
    return {"pairs": int(part[0]), "score_sum": int(part[1]), "pair_digest": int(part[2]), "n_directed": int(sel["n_directed"]), "cut_score": int(sel["cut_score"]),
            "selected": int(part[3]), "selected_score_sum": int(part[4]), "components": int(sizes.shape[0]),
            "component_size_digest": int((sizes * (np.arange(sizes.shape[0]) + 1)).sum() % (1 << 55))}


# ----------------------------------------------------------------------------------------------------------
def run_config(args, k, env):
    torch, dist, np, hga_b200 = env["torch"], env["dist"], env["np"], env["hga_b200"]
    rank, world, local_rank, dev = env["rank"], env["world"], env["local_rank"], env["dev"]

    # ---- data (untimed) ------------------------------------------------------------------------------------
    gsize = int(args.genome_mbp * 1e6)
    n_per_hap = max(1, int(args.coverage * gsize / args.mean_len))
    haps = make_haplotypes(torch, dev, gsize, args.divergence, args.seed, args.haplotypes)
    sdk = discriminative_set(torch, haps, k)
    kmers_host = sdk.cpu().numpy().astype(np.uint64)
    del sdk
    lens, starts, flip, hap = read_plan(torch, dev, gsize, n_per_hap, args.mean_len, args.seed, args.len_sigma, args.haplotypes)
    n_reads_total = int(lens.shape[0])
    if os.environ.get("HGA_BENCH_TRUE_ORDER") and world == 1:
        # experiment (profiles/r2s): the pair count's pivots in TRUE genome order (known to the generator only), the upper bound for any pivot order
        order = torch.argsort(hap * gsize + starts).to(torch.int32).cpu().numpy().astype(np.uint32)
        order.tofile("/tmp/hga_true_order.u32")
        os.environ["HGA_PAIR_ORDER"] = "3"
        os.environ["HGA_PAIR_ORDER_FILE"] = "/tmp/hga_true_order.u32"
    # contiguous shards balanced by bases
    csum = torch.cumsum(lens, 0).cpu()
    total_bases_all = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(torch.searchsorted(csum, total_bases_all * r // world).item()))
    bounds.append(n_reads_total)
    lo, hi = bounds[rank], bounds[rank + 1]
    base0 = int(csum[lo - 1]) if lo > 0 else 0
    d_bases, d_off, n_bases = synth_reads(torch, dev, list(haps), lens, starts, flip, hap, lo, hi, args.error, args.seed, base0)
    n_reads = hi - lo
    parity_n = None
    want_parity = world > 1 and not args.no_parity
    if not (want_parity and rank == 0):
        del haps
    torch.cuda.empty_cache()
    torch.cuda.synchronize()

    h = hga_b200.Handle(kmers_host, k, device=local_rank)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        uid = [hga_b200.capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(uid[0], rank, world, n_reads_total)

    def step_device():
        h.scan_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases, read_id_base=lo + 1)
        h.build_index()
        h.pair_count(min_score=1)
        h.select_edges(fraction=0.15)
        h.components(min_size=30)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        stage = []
        for _ in range(steps):
            fn()
            stage.append(h.metrics())
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, stage

    for _ in range(args.warmup):
        step_device()
    launches0 = h.metrics()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, wall_ms, stages = timed(step_device, args.steps)
    launches = h.metrics()["kernel_launches"] - launches0
    m = stages[-1]

    # ---- N > 1: the N-rank result against a single-GPU run of the SAME data on rank 0 (untimed) ---------------
    if want_parity:
        inv_n = result_invariants(np, h, dist, world, dev, torch)
        if rank == 0:
            f_bases, f_off, f_nb = synth_reads(torch, dev, list(haps), lens, starts, flip, hap, 0, n_reads_total, args.error, args.seed, 0)
            del haps
            h1 = hga_b200.Handle(kmers_host, k, device=local_rank)
            h1.set_stream(torch.cuda.current_stream().cuda_stream)
            h1.scan_device(f_bases.data_ptr(), f_off.data_ptr(), n_reads_total, f_nb, read_id_base=1)
            h1.build_index(); h1.pair_count(min_score=1); h1.select_edges(fraction=0.15); h1.components(min_size=30)
            inv_1 = result_invariants(np, h1, None, 1, dev, torch)
            h1.close()
            del f_bases, f_off
            torch.cuda.empty_cache()
            parity_n = {"status": "ok" if inv_1 == inv_n else "MISMATCH", "ranks": world, "single_gpu": inv_1, "n_gpu": inv_n}
        barrier()

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # pinned host copies of this rank's inputs (untimed setup); the timed step starts from HOST memory
        hb = torch.empty(n_bases + 64, dtype=torch.uint8, pin_memory=True)
        ho = torch.empty(n_reads + 1, dtype=torch.int64, pin_memory=True)
        hb[:n_bases].copy_(d_bases[:n_bases]); ho.copy_(d_off)
        torch.cuda.synchronize()
        d2h_bytes = [0]

        def step_host():
            h.scan_host_ptr(hb.data_ptr(), ho.data_ptr(), n_reads, read_id_base=lo + 1)
            h.build_index()
            h.pair_count(min_score=1)
            h.select_edges(fraction=0.15)
            h.components(min_size=30)
            comp = h.get_components()                                    # D2H: label per read + component list
            d2h_bytes[0] = comp["label"].nbytes + comp["comp_label"].nbytes + comp["comp_size"].nbytes

        step_host()
        e2e_ms, _, _ = timed(step_host, max(1, min(args.steps, 2)))
        e2e_steps = max(1, min(args.steps, 2))
        e2e = {"value": total_bases_all * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(n_bases + (n_reads + 1) * 8), "d2h_bytes_per_step": int(d2h_bytes[0]), "ms_per_step": e2e_ms / e2e_steps}
        del hb, ho
    clocks = sampler.stop() if rank == 0 else None

    # ---- SURVEY §8f-1 stage, reported beside the headline (not part of `value`): merge of the scaffold components + enrichment
    enrich = None
    if world == 1 and not args.no_enrich:
        step_device()
        h.enrich(min_size=30, enrichment_min_score=20)
        t0 = time.perf_counter()
        h.enrich(min_size=30, enrichment_min_score=20)
        me = h.metrics()
        enrich = {"ms": me["enrich_ms"], "wall_ms": (time.perf_counter() - t0) * 1e3, "cores": me["n_cores"], "connections": me["n_enrich_connections"],
                  "final_components": me["n_final_components"],
                  "note": "hga_enrich after the timed steps: host replay of the union_find roots + GPU merge / purge / enrichment connections + host restricted union_find"}

    # gather per-rank stage numbers (max over ranks)
    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    scan_ms = maxr(sum(s["scan_ms"] for s in stages) / len(stages))
    index_ms = maxr(sum(s["index_ms"] for s in stages) / len(stages))
    pair_ms = maxr(sum(s["pair_ms"] for s in stages) / len(stages))
    select_ms = maxr(sum(s["select_ms"] for s in stages) / len(stages))
    cc_ms = maxr(sum(s["components_ms"] for s in stages) / len(stages))
    exchange_ms = maxr(m["exchange_ms"])
    hits_total = sumr(m["n_hits"])
    pairs_total = sumr(m["n_pairs"])
    incr_total = sumr(m["n_increments"])       # every rank counts the increments of ITS k-mers' lists: the sum is the job's total

    # ---- the CPU reference and this library on the SAME bounded sample (rank 0, N = 1) -------------------------
    cpu = same = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        data = cpu_sample_data(args, k)
        cpu = run_cpu_reference(args, k, steps=3, data=data)
        if cpu:
            import datagen
            _, reads_per_hap, sdk_s = data
            seqs = [datagen.to_ascii(r).encode() for reads in reads_per_hap for r in reads]
            s_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
            np.cumsum([len(q) for q in seqs], out=s_off[1:])
            s_bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
            hs = hga_b200.Handle(np.asarray(sdk_s, dtype=np.uint64), k, device=local_rank)
            best = None
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                hs.scan(s_bases, s_off)
                hs.build_index(); hs.pair_count(min_score=1); hs.select_edges(fraction=0.15); hs.components(min_size=30)
                hs.get_components()
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            hs.close()
            same = {"gpu_gbases_per_s": int(s_off[-1]) / best / 1e9, "cpu_gbases_per_s": cpu["bases"] / (cpu["hot_ms"] * 1e-3) / 1e9,
                    "ratio": (int(s_off[-1]) / best) / (cpu["bases"] / (cpu["hot_ms"] * 1e-3)), "same_config": True,
                    "note": "both arms on the identical sample, GPU arm through hga_scan with HOST buffers (H2D + D2H inside), wall clock, best of 4"}

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        hpb = hits_total / total_bases_all
        # algorithmic bytes of the fused pack+scan kernel: 1 B/base ASCII read + 8 B per hit written
        scan_bytes_per_rank = (1.0 + 8.0 * hpb) * (total_bases_all / world)
        achieved = scan_bytes_per_rank / (scan_ms * 1e-3) / 1e9
        # measured DRAM traffic of the scan kernel: from the committed ncu capture of the SAME workload (profiles/scan_traffic.json), else null
        traffic_gb, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
                tr = json.load(f)
            if world == 1 and tr.get("workload") == workload_name(args, k):
                traffic_gb, traffic_src = tr["dram_gb_per_launch"], tr["capture"]
        except Exception:
            pass
        ms_per_step = ms_total / args.steps
        line = {
            "metric": METRIC, "value": total_bases_all / (ms_per_step * 1e-3) / 1e9, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, k), "bases": total_bases_all, "reads": n_reads_total, "kmers": int(kmers_host.shape[0]),
                       "hits_per_base": hpb, "pairs": pairs_total, "increments": incr_total,
                       "l2": "inputs (ASCII reads, hit lists, inverted index) are far larger than the 126 MB L2" if total_bases_all > (1 << 30) else
                             "inputs larger than the 126 MB L2 only in part: see bases; no L2 flush between steps",
                       "parallelism": f"reads sharded over {world} GPU(s), inverted index partitioned by k-mer owner, partial pair scores reduced at the owner of x"},
            "stages_ms": {"scan": scan_ms, "index": index_ms, "pair_count": pair_ms, "select": select_ms, "components": cc_ms},
            "scan_gbases_per_s": total_bases_all / (scan_ms * 1e-3) / 1e9,
            "pairs_per_s": pairs_total / (pair_ms * 1e-3) if pair_ms > 0 else None,
            "increments_per_s": incr_total / (pair_ms * 1e-3) if pair_ms > 0 else None,
            "diagnostics": {"filter_candidates_per_base": m["n_candidates"] / max(1, m["n_bases"]), "table_overflow_keys": m["table_overflow_keys"],
                            "table_bytes": m["table_bytes"], "filter_bytes": m["filter_bytes"], "pair_redo_rows": m["redo_pivots"], "pair_mid_rows": m["mid_pivots"],
                            "pair_heavy_rows": m["heavy_pivots"], "exchange_ms": exchange_ms},
            "roofline": {"kernel": "scan_probe_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_gb, "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                         "algorithmic_gb_per_launch": scan_bytes_per_rank / 1e9, "peak_source": peak_src, "bytes_per_base": 1.0 + 8.0 * hpb,
                         "note": "achieved = (1 + 8*hits/base) B/base x bases per GPU / scan stage time (CUDA events on the launch stream; the stage also "
                                 "holds the 1/64 sampling pre-pass, the segment reorder and the row fix-up, so the kernel alone is slightly faster)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
        }
        if parity_n:
            line["parity_n"] = parity_n["status"]
            line["parity_detail"] = parity_n
        if enrich:
            line["enrich"] = enrich
        if cpu:
            line["cpu_baseline"] = {"value": cpu["bases"] / (cpu["hot_ms"] * 1e-3) / 1e9, "unit": "Gbases/s", "cores": cpu["cores"], "kind": "reference",
                                    "sample": cpu["sample"], "scan_gbases_per_s": cpu["bases"] / (cpu["index_ms"] * 1e-3) / 1e9,
                                    "pairs_per_s": cpu["pairs"] / (cpu["connections_ms"] * 1e-3) if cpu["connections_ms"] > 0 else None}
        if same:
            line["same_sample"] = same
        print(json.dumps(line), flush=True)
    h.close()
    del d_bases, d_off
    torch.cuda.empty_cache()
    return line


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import hga_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = dict(torch=torch, dist=dist, np=np, hga_b200=hga_b200, rank=rank, world=world, local_rank=local_rank, dev=dev)
    bad = False
    for k in args.ks:                       # one JSON line per k (config 3 sweeps k; every other config has one)
        line = run_config(args, k, env)
        if line and line.get("parity_n") == "MISMATCH":
            bad = True
    if world > 1:
        dist.destroy_process_group()
    if bad:
        raise SystemExit("bench.py: the N-rank result differs from the single-GPU result of the same data (parity_detail in the JSON line)")


if __name__ == "__main__":
    main()
