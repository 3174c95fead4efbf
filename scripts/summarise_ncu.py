#!/usr/bin/env python
"""Summarise gpurun_out/<tag>_*.ncu-rep + <tag>_launches.csv into profiles/<tag>_summary.md (run here, no GPU needed)."""
import collections
import csv
import subprocess
import sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r1a"
out = Path("profiles") / f"{tag}_summary.md"
src = Path("gpurun_out")
lines = [f"# ncu summary `{tag}`", "",
         "Command: `scripts/gpu_profile.sh " + tag + "` (bench.py --steps 6 --warmup 3 --tf-changes 3 --quick --no-cpu-baseline under ncu, 1 GPU, --clock-control none).",
         "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", ""]
launch = src / f"{tag}_launches.csv"
if launch.exists():
    rows = [l for l in launch.read_text().splitlines() if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(rows):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines += ["## Launch list (gpu__time_duration.sum)", "", "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
    lines.append("")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_tex_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall"]
for rep in sorted(src.glob(f"{tag}_*.ncu-rep")):
    p = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    lines += [f"## `{rep.name}` (ncu --set full)", ""]
    for r in rows[2:]:
        lines.append(f"**{r[hdr.index('Kernel Name')][:100]}**")
        lines.append("")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"- {w}: {r[i]} {units[i]}")
        lines.append("")
# dram traffic per launch of every captured kernel -> profiles/traffic.json (bench.py's roofline.traffic)
import json
import re
traffic = {}
for rep in sorted(src.glob(f"{tag}_*.ncu-rep")):
    p = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    if len(rows) < 3 or "dram__bytes_read.sum" not in rows[0]:
        continue
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in rows[2:]:
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1)
        vals.append(tot)
    name = re.sub(r"^void |<.*$|\(.*$", "", rows[2][hdr.index("Kernel Name")]).split("::")[-1]
    traffic[name] = sum(vals) / len(vals)
if traffic:
    tj = Path("profiles") / "traffic.json"
    old = json.loads(tj.read_text()) if tj.exists() else {}
    old.update(traffic)
    old["_source"] = f"ncu --set full captures of round tag {tag} (scripts/gpu_profile.sh), dram__bytes_read.sum + dram__bytes_write.sum per launch"
    tj.write_text(json.dumps(old, indent=1))
out.parent.mkdir(exist_ok=True)
out.write_text("\n".join(lines))
print("wrote", out)
