import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import bench
from vkvolume_b200 import capi
from vkvolume_b200.capi import VolumeOptions
ctx = capi.Context(0)
for dims in ((512, 512, 512), (1024, 1024, 1024), (1024, 1024, 1024)):
    W, H, D = dims
    vol = capi.Volume(ctx, W, H, D)
    capi.synth_volume(ctx, 0, 0x5EED0004, W, H, D, vol.device_voxels())
    vol.upload_device(vol.device_voxels())
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    if len(sys.argv) > 1 and sys.argv[1] == "sync":
        torch.cuda.synchronize()
    opt = bench.sweep_options(None, 0)
    vol.update_transfer_function(opt, 1, count=False)
    O2 = vol.download_distance_map(0)
    vol.update_transfer_function(opt, 1, count=False)
    O3 = vol.download_distance_map(0)
    n = vol.update_transfer_function(opt, 1, count=True)
    O1 = vol.download_distance_map(0)
    print(dims, "first nocount %.4f second nocount %.4f count %.4f" % ((O2 == 0).mean(), (O3 == 0).mean(), (O1 == 0).mean()), flush=True)
    vol.close()
