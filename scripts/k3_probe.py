#!/usr/bin/env python
"""Debug probe (GPU box): K3 alone on a map of a given extent (a block_size = 1 volume whose occupancy is a thresholded synthetic
volume): event times of the distance transform, isotropic and anisotropic; run under ncu for the per-kernel split."""
import argparse, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vkvolume_b200 import capi
from vkvolume_b200.capi import VolumeOptions
ap = argparse.ArgumentParser()
ap.add_argument("--dims", default="1024,1024,512")
ap.add_argument("--kind", type=int, default=3)
ap.add_argument("--imin", type=float, default=0.3)
ap.add_argument("--modes", default="2,3")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
import torch
W, H, D = [int(x) for x in a.dims.split(",")]
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=1, use_precomputed_gradient=False)
capi.synth_volume(ctx, a.kind, 0x5EED0005, W, H, D, vol.device_voxels())
vol.upload_device(vol.device_voxels())
opt = VolumeOptions(intensity_min=a.imin, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
tfu = capi.transfer_function_uniform(opt)
vol.update_transfer_function_texture(opt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in [int(x) for x in a.modes.split(",")]:
    ts = []
    for r in range(a.reps):
        vol.compute_occupancy_slab(tfu, mode, 0, D)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vol.compute_distance_from_occupancy(mode)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    M = W * H * D
    by = (6 if mode == 2 else 28) * M
    print(f"map {W}x{H}x{D} mode {mode}: distance transform {np.median(ts):.4f} ms (min {min(ts):.4f})  -> {by / np.median(ts) / 1e6:.1f} GB/s algorithmic", flush=True)
    if a.check:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_api as orc
        vol.compute_occupancy_slab(tfu, 1, 0, D)
        O = vol.download_distance_map(0)
        vol.compute_occupancy_slab(tfu, mode, 0, D)
        vol.compute_distance_from_occupancy(mode)
        t0 = time.time()
        if mode == 2:
            ok = np.array_equal(vol.download_distance_map(0), orc.distance_map(O))
        else:
            want = orc.distance_map_anisotropic(O)
            ok = all(np.array_equal(vol.download_distance_map(i), want[i]) for i in range(8))
        print(f"   occupied {float((O == 0).mean()):.4f}  bit-exact vs oracle: {ok}  ({time.time() - t0:.1f} s)", flush=True)
