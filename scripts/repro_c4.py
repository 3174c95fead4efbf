#!/usr/bin/env python
"""Debug probe (GPU box): config-4 sized rebuild with a gradient transfer function, stage by stage with a sync after each."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from vkvolume_b200 import capi  # noqa: E402
from vkvolume_b200.capi import VolumeOptions  # noqa: E402

W, H, D = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024x1024x1024").split("x")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, 0, 0x5EED0004, W, H, D, vol.device_voxels(), 0)
vol.upload_device(vol.device_voxels(), 0)
torch.cuda.synchronize(); print("upload ok", flush=True)
opt = VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.05, gradient_max=0.25)
tfu = capi.transfer_function_uniform(opt)
vol.compute_gradient_map(tfu, 0)
torch.cuda.synchronize(); print("gradient ok", flush=True)
vol.update_transfer_function_texture(opt, 0)
torch.cuda.synchronize(); print("tf ok", flush=True)
vol.compute_occupancy_slab(tfu, skip, 0, vol.map_extent[2], stream=0)
torch.cuda.synchronize(); print("occupancy ok", flush=True)
vol.compute_distance_from_occupancy(skip, 0)
torch.cuda.synchronize(); print("distance ok", flush=True)
n = vol.compute_occupied_voxel_count(tfu, 0)
print("count", n, flush=True)
