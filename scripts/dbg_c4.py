import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import bench
from vkvolume_b200 import capi
from vkvolume_b200.capi import VolumeOptions
ctx = capi.Context(0)
class P:
    def __init__(s, p, n): s.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 3}
for dims in ((256, 256, 256), (512, 512, 512), (1024, 1024, 1024)):
    W, H, D = dims
    vol = capi.Volume(ctx, W, H, D)
    capi.synth_volume(ctx, 0, 0x5EED0004, W, H, D, vol.device_voxels())
    torch.cuda.synchronize()
    v = torch.as_tensor(P(vol.device_voxels(), W * H * D), device="cuda")
    print(dims, "max", int(v.max()), "frac>12", float((v > 12).float().mean()), flush=True)
    vol.upload_device(vol.device_voxels())
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    for k in (0, 1):
        opt = bench.sweep_options(None, k)
        n = vol.update_transfer_function(opt, 1, count=True)
        O = vol.download_distance_map(0)
        print("  k", k, "count", n, "occ frac", float((O == 0).mean()), flush=True)
    vol.close()
