#!/usr/bin/env python
"""GPU-box probe: the occupancy pass alone (K2a, + fused count) on a bench workload, with and without a gradient TF, L2 flushed."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from vkvolume_b200 import capi  # noqa: E402
from vkvolume_b200.capi import VolumeOptions  # noqa: E402

wl = bench.WORKLOADS[os.environ.get("OCC_PROBE_WORKLOAD", "c2")]
W, H, D = wl["dim"]
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), 0)
vol.upload_device(vol.device_voxels(), 0)
vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)), 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
Db = vol.map_extent[2]
N = W * H * D
for name, tf in (("no gradient", dict(intensity_min=0.086, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)),
                 ("gradient", dict(intensity_min=0.086, intensity_max=1.0, gradient_min=0.1, gradient_max=0.3)),
                 ("gradient wide", dict(intensity_min=0.05, intensity_max=1.0, gradient_min=0.05, gradient_max=0.25))):
    opt = VolumeOptions(**tf)
    tfu = capi.transfer_function_uniform(opt)
    vol.update_transfer_function_texture(opt)
    for count in (False, True):
        ts = []
        for _ in range(7):
            flush.fill_(1)
            cnt.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            vol.compute_occupancy_slab(tfu, capi.SKIP_DISTANCE, 0, Db, count_dev=cnt.data_ptr() if count else 0)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts[2:]))
        nbytes = (2 if tfu.use_gradient else 1) * N
        print(f"{name:14s} count={int(count)}  {t:.4f} ms  {nbytes / t / 1e6:.0f} GB/s", flush=True)
