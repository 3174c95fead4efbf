#!/usr/bin/env python
"""Warp instructions executed + stall samples per CUDA source line of the first kernel in an .ncu-rep (read here, no GPU)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if "# Samples" in r)
h = rows[hdr]
ss, ie = h.index("# Samples"), h.index("Instructions Executed")
agg = {}
done = False
for r in rows[hdr + 1:]:
    if len(r) != len(h):
        continue
    if not r[0].strip().isdigit():
        continue
    key = (r[0], r[1])
    if r[ie] in ("", "-"):
        continue
    a = agg.setdefault(key, [0, 0, r])
    try:
        a[0] += int(r[ie] or 0)
        a[1] += int(r[ss] or 0)
    except ValueError:
        pass
print(h[:4])
ti = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
print(f"{ti} warp instructions, {ts} samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / max(ti, 1):5.1f}% inst {100 * a[1] / max(ts, 1):5.1f}% samp  {k[0][:12]:>12} {k[1].strip()[:120]}")
