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
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# synthetic data on the GPU (torch = plumbing)
# ----------------------------------------------------------------------------------------------------------
def make_haplotypes(torch, dev, genome_size, divergence, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    a = torch.randint(0, 4, (genome_size,), dtype=torch.uint8, device=dev, generator=g)
    mut = (torch.rand(genome_size, device=dev, generator=g) < divergence).to(torch.uint8)
    shift = torch.randint(1, 4, (genome_size,), dtype=torch.uint8, device=dev, generator=g)
    b = (a + mut * shift) % 4        # scripts/read_generator.py:158-162
    return a, b


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
    """canonical k-mers present in exactly one haplotype (sorted, int64 on the GPU)"""
    uniq = [torch.unique(canonical_kmers(torch, h, k)) for h in haps]
    allk, counts = torch.unique(torch.cat(uniq), return_counts=True)
    return allk[counts == 1]


def read_plan(torch, dev, genome_size, n_reads_per_hap, mean_len, seed):
    """lengths / starts / strand of every read of both haplotypes (identical on every rank)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 7)
    n = 2 * n_reads_per_hap
    sigma = 0.5
    mu = torch.log(torch.tensor(mean_len)) - 0.5 * sigma * sigma
    lens = torch.exp(mu + sigma * torch.randn(n, device=dev, generator=g)).to(torch.int64).clamp_(100, min(60000, genome_size))
    starts = (torch.rand(n, device=dev, generator=g, dtype=torch.float64) * (genome_size - lens + 1).to(torch.float64)).to(torch.int64)
    flip = torch.rand(n, device=dev, generator=g) < 0.5
    hap = (torch.arange(n, device=dev) >= n_reads_per_hap).to(torch.int64)
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
    genome = torch.stack(haps)          # [2, G]
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
def run_cpu_reference(args, steps=1, warmup=0):
    import numpy as np
    import datagen
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        return None
    cores = os.cpu_count() or 1
    gsize = int(args.cpu_sample_mbp * 1e6)
    n_per_hap = max(1, int(args.coverage * gsize / args.mean_len))
    with tempfile.TemporaryDirectory() as d:
        a = datagen.random_genome(gsize, args.seed)
        b = datagen.mutate(a, args.divergence, args.seed + 1)
        paths = []
        total = 0
        for i, hp in enumerate((a, b)):
            reads = datagen.sample_reads(hp, n_per_hap, int(args.mean_len), args.seed + 10 + i, error_rate=args.error, length_sigma=0.5,
                                         min_len=100, max_len=min(60000, gsize))
            total += sum(len(r) for r in reads)
            p = os.path.join(d, f"hap{i}.fa")
            datagen.write_fasta(p, reads, prefix=f"h{i}_")
            paths.append(p)
        sdk = datagen.discriminative_kmers([a, b], K)
        kp = os.path.join(d, f"{K}-mers.txt")
        datagen.write_kmers(kp, sdk, K)
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
    best["sample"] = (f"same generator scaled to a {args.cpu_sample_mbp} Mbp diploid ({best['bases'] / 1e6:.1f} Mbases of reads, {best['kmers']} "
                      f"{K}-mers), reference stage timers index + connections + union-find, --threads {cores}")
    return best


def reference_arm(args, rank):
    if rank != 0:
        return
    res = run_cpu_reference(args, steps=max(1, args.steps), warmup=0)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built (reference sources absent at build time)"}))
        return
    value = res["bases"] / (res["hot_ms"] * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["hot_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_sample": res["sample"]},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                         "scan_gbases_per_s": res["bases"] / (res["index_ms"] * 1e-3) / 1e9,
                         "pairs_per_s": res["pairs"] / (res["connections_ms"] * 1e-3) if res["connections_ms"] > 0 else None},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args):
    return (f"config4: synthetic {args.genome_mbp:g} Mbp diploid, {args.divergence * 100:g}% divergence, {args.coverage:g}x long reads "
            f"(lognormal, mean {args.mean_len:g} bp), {args.error * 100:g}% substitution errors, k={K}, SDK = canonical k-mers in exactly one haplotype")


# ----------------------------------------------------------------------------------------------------------
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

    # ---- data (untimed) ------------------------------------------------------------------------------------
    gsize = int(args.genome_mbp * 1e6)
    n_per_hap = max(1, int(args.coverage * gsize / args.mean_len))
    haps = make_haplotypes(torch, dev, gsize, args.divergence, args.seed)
    sdk = discriminative_set(torch, haps, K)
    kmers_host = sdk.cpu().numpy().astype(np.uint64)
    del sdk
    lens, starts, flip, hap = read_plan(torch, dev, gsize, n_per_hap, args.mean_len, args.seed)
    n_reads_total = int(lens.shape[0])
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
    del haps
    torch.cuda.empty_cache()
    torch.cuda.synchronize()

    h = hga_b200.Handle(kmers_host, K, device=local_rank)
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
    hits_total = sumr(m["n_hits"])
    pairs_total = sumr(m["n_pairs"])
    incr_total = sumr(m["n_increments"])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_reference(args)

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
        # measured DRAM traffic of the scan kernel for the default workload on one GPU (ncu --set full, profiles/r05z_scan_pair_ncu_full.csv)
        default_cfg = (args.genome_mbp, args.coverage, args.mean_len, args.divergence, args.error, args.seed) == (100.0, 50.0, 10000.0, 0.01, 0.05, 4000)
        traffic_gb = 120.0 if (default_cfg and world == 1) else None
        ms_per_step = ms_total / args.steps
        line = {
            "metric": METRIC, "value": total_bases_all / (ms_per_step * 1e-3) / 1e9, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "bases": total_bases_all, "reads": n_reads_total, "kmers": int(kmers_host.shape[0]),
                       "hits_per_base": hpb, "pairs": pairs_total, "increments": incr_total,
                       "l2": "inputs (ASCII reads, hit lists, inverted index) are far larger than the 126 MB L2", "parallelism": f"reads sharded over {world} GPU(s)"},
            "stages_ms": {"scan": scan_ms, "index": index_ms, "pair_count": pair_ms, "select": select_ms, "components": cc_ms},
            "scan_gbases_per_s": total_bases_all / (scan_ms * 1e-3) / 1e9,
            "pairs_per_s": pairs_total / (pair_ms * 1e-3) if pair_ms > 0 else None,
            "increments_per_s": incr_total / (pair_ms * 1e-3) if pair_ms > 0 else None,
            "diagnostics": {"filter_candidates_per_base": m["n_candidates"] / max(1, m["n_bases"]), "table_overflow_keys": m["table_overflow_keys"],
                            "table_bytes": m["table_bytes"], "filter_bytes": m["filter_bytes"], "pair_redo_rows": m["redo_pivots"], "pair_mid_rows": m["mid_pivots"],
                            "pair_heavy_rows": m["heavy_pivots"], "exchange_ms": m["exchange_ms"]},
            "roofline": {"kernel": "scan_probe_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_gb, "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu capture profiles/r05z, same workload)",
                         "algorithmic_gb_per_launch": scan_bytes_per_rank / 1e9, "peak_source": peak_src, "bytes_per_base": 1.0 + 8.0 * hpb,
                         "note": "achieved = (1 + 8*hits/base) B/base x bases per GPU / scan stage time (CUDA events on the launch stream; the stage also "
                                 "holds the 1/64 sampling pre-pass, the segment reorder and the row fix-up, so the kernel alone is slightly faster)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
        }
        if enrich:
            line["enrich"] = enrich
        if cpu:
            line["cpu_baseline"] = {"value": cpu["bases"] / (cpu["hot_ms"] * 1e-3) / 1e9, "unit": "Gbases/s", "cores": cpu["cores"], "kind": "reference",
                                    "sample": cpu["sample"], "scan_gbases_per_s": cpu["bases"] / (cpu["index_ms"] * 1e-3) / 1e9,
                                    "pairs_per_s": cpu["pairs"] / (cpu["connections_ms"] * 1e-3) if cpu["connections_ms"] > 0 else None}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
