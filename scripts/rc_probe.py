#!/usr/bin/env python
"""Debug probe (GPU box): where a ray-cast frame's time goes — set-up only (TEST_RAY_ENTRY view: entry geometry + epilogue, no
march) against the full frame, per skip mode, 24 orbit views with the L2 flushed."""
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
    ap.add_argument("--skips", default="2")
    ap.add_argument("--tests", default="1,0")
    a = ap.parse_args()
    import torch
    wl = bench.WORKLOADS[a.workload]
    W, H, D = wl["dim"]
    FW, FH = wl["frame"]
    ctx = capi.Context(0)
    vol = capi.Volume(ctx, W, H, D, block_size=4)
    capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), 0)
    vol.upload_device(vol.device_voxels(), 0)
    opt = VolumeOptions(**wl["tf"])
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(tfu, 0)
    it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
    fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    views = [vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(v, 72, wl), aspect=FW / FH), it, wl["clip"]) for v in range(0, 72, 3)]
    for skip in [int(x) for x in a.skips.split(",")]:
        vol.update_transfer_function(opt, skip)
        for test in [int(x) for x in a.tests.split(",")]:
            ropt = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1, test=test)
            ts = []
            counts.zero_()
            for rep in range(2):
                for cu, ru in views:
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    vol.render(cu, ru, tfu, ropt, FW, FH, fb.data_ptr(), 0, counts.data_ptr() if rep == 0 else 0, 0)
                    e1.record()
                    e1.synchronize()
                    if rep == 1:
                        ts.append(e0.elapsed_time(e1))
            c = counts.tolist()
            print(f"skip {skip} test {test}: mean {np.mean(ts):.4f} ms  min {np.min(ts):.4f}  max {np.max(ts):.4f}  covered/frame {c[3] / len(views):.0f} samples/frame {(c[0] + c[1]) / len(views):.0f}", flush=True)


if __name__ == "__main__":
    main()
