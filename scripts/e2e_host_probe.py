#!/usr/bin/env python
"""GPU-box probe: where the pipelined e2e frame time goes on the host side — Python camera maths, the C call (enqueue only), and the
PCIe floor (back-to-back 8.3 MB D2H copies)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from vkvolume_b200 import capi, scene
from vkvolume_b200.capi import RenderOptions, VolumeOptions

wl = bench.WORKLOADS["c2"]
W, H, D = wl["dim"]; FW, FH = wl["frame"]
ctx = capi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), stream)
vol.upload_device(vol.device_voxels(), stream)
opt = VolumeOptions(**wl["tf"]); tfu = capi.transfer_function_uniform(opt)
vol.compute_gradient_map(tfu, stream)
vol.update_transfer_function(opt, 2, stream=stream)
it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
ropt = RenderOptions(skipping_type=2, clip_distance=wl["clip"], early_ray_termination=1)
def uniforms(s):
    return vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(s, 72, wl), aspect=FW / FH), it, wl["clip"])
n = 72
t0 = time.perf_counter()
for s in range(n): uniforms(s)
print(f"python uniforms(): {(time.perf_counter() - t0) / n * 1e6:.1f} us per frame", flush=True)
unis = [uniforms(s) for s in range(n)]
ring = torch.empty((3, FH, FW, 4), dtype=torch.uint8).pin_memory()
cnt = torch.zeros((n, 4), dtype=torch.int64).pin_memory()
for label, pre in (("with python maths", False), ("uniforms precomputed", True)):
    for s in range(3): vol.render_to_host_async(*unis[s], tfu, ropt, FW, FH, ring[s % 3].data_ptr(), cnt[s].data_ptr(), stream)
    vol.render_to_host_wait(stream); torch.cuda.synchronize()
    t0 = time.perf_counter(); tc = 0.0
    for s in range(n):
        cu, ru = unis[s] if pre else uniforms(s)
        a = time.perf_counter()
        vol.render_to_host_async(cu, ru, tfu, ropt, FW, FH, ring[s % 3].data_ptr(), cnt[s].data_ptr(), stream)
        tc += time.perf_counter() - a
    vol.render_to_host_wait(stream)
    dt = time.perf_counter() - t0
    print(f"{label}: {dt / n * 1e3:.4f} ms per frame; the async call itself {tc / n * 1e6:.1f} us per frame", flush=True)
fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(60): ring[k % 3].copy_(fb, non_blocking=True)
torch.cuda.synchronize()
print(f"PCIe floor: {(time.perf_counter() - t0) / 60 * 1e3:.4f} ms per 8.3 MB frame back to back", flush=True)
