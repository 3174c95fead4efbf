#!/usr/bin/env python
"""Debug probe (GPU box): A/B the ray caster's VKV_RC_FLAGS on the headline frame (24 orbit views, L2 flushed)."""
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
    ap.add_argument("--flags", default="0,1,2,3")
    ap.add_argument("--skips", default="2,1,3,0")
    ap.add_argument("--env", default="VKV_RC_FLAGS")
    ap.add_argument("--view-step", type=int, default=3, help="orbit views between consecutive frames (bench.py: 1 = 5 degrees)")
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
    views = [vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(v, 72, wl), aspect=FW / FH), it, wl["clip"]) for v in range(0, 24 * a.view_step, a.view_step)]
    for skip in [int(x) for x in a.skips.split(",")]:
        vol.update_transfer_function(opt, skip)
        ropt = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1)
        ref = None
        for fl in a.flags.split(","):
            os.environ[a.env] = fl
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
            frame = fb.clone()
            same = "ref" if ref is None else ("same frame" if torch.equal(frame, ref[0]) and c == ref[1] else f"DIFFERENT ({int((frame != ref[0]).sum())} bytes, counts {c} vs {ref[1]})")
            if ref is None:
                ref = (frame, c)
            print(f"skip {skip} {a.env}={fl}: mean {np.mean(ts):.4f} ms  min {np.min(ts):.4f}  max {np.max(ts):.4f}   samples/frame {(c[0] + c[1]) / len(views):.0f}  [{same}]", flush=True)


if __name__ == "__main__":
    main()
