#!/usr/bin/env python
"""Hot SASS lines of an ncu report (run here): python scripts/ncu_hotlines.py <file.ncu-rep> [min share %]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one block per kernel: a "Kernel Name" row, a header row, then the lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "body": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["body"].append(r)
for b in blocks:
    hdr = b["hdr"]
    si, ei, ai = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[si] or 0) for r in b["body"])
    print("==", b["name"][:80], "total samples", tot)
    for n, r in enumerate(b["body"]):
        s = int(r[si] or 0)
        if tot and s >= tot * thr / 100:
            print(f"{n:5d} {s:7d} {100 * s / tot:5.1f}%  exec {r[ei]:>10}  {r[ai].strip()[:90]}")
