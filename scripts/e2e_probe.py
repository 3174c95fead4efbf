#!/usr/bin/env python
"""GPU-box probe: e2e frame time of vkv_render_to_host on config 2 for several band counts (VKV_E2E_BANDS)."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

import bench
from vkvolume_b200 import capi, scene
from vkvolume_b200.capi import RenderOptions, VolumeOptions

wl = bench.WORKLOADS["c2"]
W, H, D = wl["dim"]
FW, FH = wl["frame"]
ctx = capi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), stream)
vol.upload_device(vol.device_voxels(), stream)
opt = VolumeOptions(**wl["tf"])
tfu = capi.transfer_function_uniform(opt)
vol.compute_gradient_map(tfu, stream)
vol.update_transfer_function(opt, 2, stream=stream)
it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
ropt = RenderOptions(skipping_type=2, clip_distance=wl["clip"], early_ray_termination=1)
host_fb = torch.empty((FH, FW, 4), dtype=torch.uint8).pin_memory()
host_np = host_fb.numpy()
unis = [vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(s, 72, wl), aspect=FW / FH), it, wl["clip"]) for s in range(72)]
for bands in (1, 2, 3, 4, 6, 8):
    os.environ["VKV_E2E_BANDS"] = str(bands)
    for s in range(5):
        vol.render_to_host(*unis[s], tfu, ropt, FW, FH, out=host_np, stream=stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 60
    for s in range(n):
        vol.render_to_host(*unis[s % 72], tfu, ropt, FW, FH, out=host_np, stream=stream)
    dt = (time.perf_counter() - t0) / n
    # same without the counters (pageable ctypes struct on the old path)
    print(f"bands {bands}: {dt * 1e3:.4f} ms/frame", flush=True)
# raw D2H of the frame alone
fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    host_fb.copy_(fb, non_blocking=True)
    torch.cuda.synchronize()
print(f"raw D2H 8.3 MB pinned: {(time.perf_counter() - t0) / 50 * 1e3:.4f} ms")
