#!/usr/bin/env python
"""GPU-box check: on a volume above the flat-walk limit (2 G voxels) the row-task gradient kernel (default there) and the flat walk
(VKV_GRAD_FLAT=1) produce identical maps; prints both times."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402
from vkvolume_b200 import capi  # noqa: E402
from vkvolume_b200.capi import VolumeOptions  # noqa: E402

W, H, D = 2048, 2048, 1024
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, 3, 0x5EED0005, W, H, D, vol.device_voxels(), 0)
vol.upload_device(vol.device_voxels(), 0)
tfu = capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2))
g = torch.as_tensor(bench._DevPtr(capi.lib().vkv_volume_device_gradient(vol.handle), W * H * D), device="cuda")
maps = []
for env in ({}, {"VKV_GRAD_FLAT": "1"}):
    os.environ.pop("VKV_GRAD_FLAT", None)
    os.environ.update(env)
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); vol.compute_gradient_map(tfu, 0); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    maps.append(g.clone())
    print(env or "default (row-task kernel)", " ".join(f"{t:.3f}" for t in ts), "ms", flush=True)
print("identical maps:", bool(torch.equal(maps[0], maps[1])), " nonzero:", int((maps[0] > 0).sum().item()))
