#!/usr/bin/env python
"""Print the interesting keys of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} {d['unit']}  ms/frame {d['ms_per_step']:.4f}  samples/frame {d['samples_per_frame']:.0f}  launches {d['gpu_launches']}")
r = d["ess_rebuild_ms"]
print("rebuild median", round(r["median"], 4), {k: round(v, 4) for k, v in r.get("stages", {}).items()})
for k, v in d.get("modes", {}).items():
    print(f"  mode {k:22s} {v['ms_per_frame']:.4f} ms  {v['msamples_per_s']:.0f} Ms/s")
e = d["e2e"]
print(f"e2e {e['value']:.0f} {e['unit']}  {e['ms_per_frame']:.4f} ms/frame")
rf = d["roofline"]
print(f"roofline raycast: achieved {rf['achieved']:.1f} peak {rf['peak']:.1f} frac {rf['frac']:.4f}")
for k, v in d.get("rooflines_hbm", {}).items():
    print(f"  {k:32s} {v['ms']:.4f} ms  {v['achieved']:.0f} GB/s  frac {v['frac']:.3f}")
print("cpu", d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"), "clocks", d.get("clocks"))
