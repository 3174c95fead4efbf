#!/usr/bin/env python
"""Debug probe (GPU box): a few frames of one workload / skip mode with per-frame event times (env knobs apply)."""
import argparse, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from vkvolume_b200 import capi, scene  # noqa: E402
from vkvolume_b200.capi import RenderOptions, VolumeOptions  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--skip", type=int, default=2)
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--test", type=int, default=0)
a = ap.parse_args()
import torch
wl = bench.WORKLOADS[a.workload]
W, H, D = wl["dim"]; FW, FH = wl["frame"]
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), 0)
vol.upload_device(vol.device_voxels(), 0)
opt = VolumeOptions(**wl["tf"]); tfu = capi.transfer_function_uniform(opt)
vol.compute_gradient_map(tfu, 0)
vol.update_transfer_function(opt, a.skip)
it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
counts = torch.zeros(4, dtype=torch.int64, device="cuda")
ropt = RenderOptions(skipping_type=a.skip, clip_distance=wl["clip"], early_ray_termination=1, test=a.test)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in range(a.frames):
    cu, ru = vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(3 * k, 72, wl), aspect=FW / FH), it, wl["clip"])
    flush.fill_(1)
    counts.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vol.render(cu, ru, tfu, ropt, FW, FH, fb.data_ptr(), 0, counts.data_ptr(), 0)
    e1.record(); e1.synchronize()
    print(f"frame {k}: {e0.elapsed_time(e1):.4f} ms counts {counts.tolist()}", flush=True)
