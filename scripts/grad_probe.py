#!/usr/bin/env python
"""GPU-box probe: gradient-map stage time (kernel + array sync) on config 2 under the A/B environment switches."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

import bench
from vkvolume_b200 import capi
from vkvolume_b200.capi import VolumeOptions

wl = bench.WORKLOADS[os.environ.get("GRAD_PROBE_WORKLOAD", "c2")]
W, H, D = wl["dim"]
ctx = capi.Context(0)
stream = torch.cuda.current_stream().cuda_stream
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), stream)
vol.upload_device(vol.device_voxels(), stream)
tfu = capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
VARIANTS = {"v1_int_surf": {"VKV_GRAD_V1": "1"}, "flat_nosurf": {"VKV_GRAD_NOSURF": "1"}, "flat_surf": {},
            "flat_surf_c2": {"VKV_GRAD_CTAS": "2"}, "flat_surf_c4": {"VKV_GRAD_CTAS": "4"}, "flat_surf_c3": {"VKV_GRAD_CTAS": "3"},
            "flat_surf_c1": {"VKV_GRAD_CTAS": "1"}}
only = os.environ.get("GRAD_PROBE_ONLY")
for name, env in VARIANTS.items():
    if only and name not in only.split(","):
        continue
    for k in ("VKV_GRAD_FP32", "VKV_GRAD_NOSURF", "VKV_GRAD_V1", "VKV_GRAD_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        vol.compute_gradient_map(tfu, stream)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    print(name, " ".join(f"{t:.3f}" for t in ts), flush=True)
