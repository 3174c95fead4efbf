#!/usr/bin/env python
"""Top stalled SASS instructions + stall-reason totals of the first kernel in an .ncu-rep (read here, no GPU)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
start = hdr_idx[0]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
h = rows[start]
body = [r for r in rows[start + 1:end] if len(r) == len(h)]
si, ss, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ss] or 0) for r in body)
print(f"{rep}: {len(body)} SASS instructions, {tot} samples, {sum(int(r[ie] or 0) for r in body)} warp instructions executed")
agg = {h[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print("mix:", end=" ")
mix = {}
for r in body:
    op = r[si].split()[0] if r[si].split() else "?"
    if op.startswith("@"):
        op = r[si].split()[1]
    op = op.split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[ie] or 0)
ti = sum(mix.values())
print(", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
for r in sorted(body, key=lambda r: -int(r[ss] or 0))[:top]:
    reasons = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{100 * int(r[ss]) / max(tot, 1):5.1f}%  {r[si][:70]:70s} {reasons}")
