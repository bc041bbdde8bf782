#!/usr/bin/env python
"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python scripts/launch_shares.py gpurun_out/launches_X.csv [skip_launches]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; ix = {n: j for j, n in enumerate(hdr)}
tot = collections.defaultdict(float); cnt = collections.Counter(); n = 0
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = r[ix["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].split("<")[0][-48:]
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1000 if u.startswith("us") else v / 1e6 if u.startswith("ns") else v * 1000 if u in ("s", "second") else v
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print(f"kernel,launches,total_ms,share_pct   # {n - skip} launches, {s:.3f} ms")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k},{cnt[k]},{v:.3f},{v / s * 100:.1f}")
