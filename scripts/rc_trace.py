#!/usr/bin/env python
"""Debug probe (GPU box): per-warp timeline of one ray-cast launch of the headline frame.

Runs config 2's view `--view` with VKV_RC_TRACE set (raycast.cu dumps {start ns, end ns, loop iterations} per warp),
prints where the kernel's time goes — critical path (longest warp) vs. throughput — and saves the raw trace under gpurun_out/.
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from vkvolume_b200 import capi, scene  # noqa: E402
from vkvolume_b200.capi import RenderOptions, VolumeOptions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--view", type=int, default=0)
    ap.add_argument("--skip", type=int, default=None)
    ap.add_argument("--tag", default="trace")
    a = ap.parse_args()
    import torch
    wl = bench.WORKLOADS[a.workload]
    W, H, D = wl["dim"]
    FW, FH = wl["frame"]
    skip = wl["skip"] if a.skip is None else a.skip
    ctx = capi.Context(0)
    vol = capi.Volume(ctx, W, H, D, block_size=4)
    capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), 0)
    vol.upload_device(vol.device_voxels(), 0)
    opt = VolumeOptions(**wl["tf"])
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(tfu, 0)
    vol.update_transfer_function(opt, skip)
    it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
    ropt = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1)
    fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cu, ru = vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(a.view, 72, wl), aspect=FW / FH), it, wl["clip"])
    for _ in range(3):
        vol.render(cu, ru, tfu, ropt, FW, FH, fb.data_ptr(), 0, 0, 0)
    torch.cuda.synchronize()
    flush.fill_(1)
    torch.cuda.synchronize()
    out = ROOT / "gpurun_out" / f"{a.tag}_{a.workload}_v{a.view}_s{skip}.u64"
    out.parent.mkdir(exist_ok=True)
    os.environ["VKV_RC_TRACE"] = str(out)
    vol.render(cu, ru, tfu, ropt, FW, FH, fb.data_ptr(), 0, 0, 0)
    torch.cuda.synchronize()
    del os.environ["VKV_RC_TRACE"]
    t = np.fromfile(out, dtype=np.uint64).reshape(-1, 8)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    start, end = (t[:, 0] - t0).astype(np.int64), (t[:, 1] - t0).astype(np.int64)
    w = t[:, 2]
    it_, it_d, it_r, it_m, lanes = [((w >> np.uint64(s)) & np.uint64(0xfff)).astype(np.int64) for s in (0, 12, 24, 36, 48)]
    dur = end - start
    print(f"warps traced {len(t)}  kernel span {end.max() / 1e3:.1f} us  last start {start.max() / 1e3:.1f} us")
    print(f"warp duration us: mean {dur.mean() / 1e3:.2f} p50 {np.median(dur) / 1e3:.2f} p99 {np.percentile(dur, 99) / 1e3:.2f} max {dur.max() / 1e3:.2f}")
    k = np.argsort(-dur)[:8]
    for i in k:
        print(f"  long warp: start {start[i] / 1e3:7.1f} us  dur {dur[i] / 1e3:7.1f} us  iterations {it_[i]}  ns/iter {dur[i] / max(it_[i], 1):.0f}"
              f"  with skip load {it_d[i]}  with tex batch {it_r[i]}  mixed {it_m[i]}  mean live lanes {lanes[i]}"
              f"  | kcycles head {t[i, 3] / 1e3:.1f} issue {t[i, 4] / 1e3:.1f} skip branch {t[i, 5] / 1e3:.1f} sample branch {t[i, 6] / 1e3:.1f}")
    act = it_ > 0
    for thr in (16, 24, 32, 48, 64, 96, 128, 160):
        m = it_ >= thr
        print(f"  warps with >= {thr:3d} iterations: {int(m.sum()):6d}   live lanes (mean x warps) {int(lanes[m].sum()):7d}   iterations in them {int(it_[m].sum())}")
    print(f"all marching warps: iterations {it_[act].sum()}  with skip load {it_d[act].sum()}  with tex batch {it_r[act].sum()}  mixed {it_m[act].sum()}  mean live lanes {(lanes[act] * it_[act]).sum() / it_[act].sum():.1f}")
    print(f"marching warps {act.sum()}  mean iterations {it_[act].mean():.1f}  mean ns/iter {dur[act].sum() / it_[act].sum():.0f}")
    # concurrency over time (warps in flight, whole GPU) in 5 us buckets
    edges = np.arange(0, end.max() + 5000, 5000)
    conc = [(int(((start < e + 5000) & (end > e)).sum())) for e in edges[:-1]]
    print("warps in flight per 5 us bucket:", conc)


if __name__ == "__main__":
    main()
