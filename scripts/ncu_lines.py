#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python scripts/ncu_lines.py <rep> <kernel regex> [launch-skip] [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = ""
hdr = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; ix = {}
        for j, n in enumerate(hdr):
            ix.setdefault(n, j)
        continue
    if hdr is None or len(r) < len(hdr) or r[ix["Address"]] != "-":
        continue
    try:
        ie = int(r[ix["Instructions Executed"]]); smp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    l2 = r[ix["L2 Theoretical Sectors Global"]]; shw = r[ix["L1 Wavefronts Shared"]]
    out.append((ie, smp, cur_file, r[0], r[1].strip()[:120], l2, shw))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print(f"total warp-inst {tot}  samples {ts}")
key = (lambda o: -o[0]) if (len(sys.argv) > 5 and sys.argv[5] == "inst") else (lambda o: -o[1])
for o in sorted(out, key=key)[:top]:
    print(f"{o[0]/tot*100:5.1f}%inst {o[1]/ts*100:5.1f}%smp L2sec={o[5]:>12} shw={o[6]:>11} {o[2]}:{o[3]:>4} {o[4]}")
