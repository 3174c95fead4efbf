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
VARIANTS = {"walk": {}, **{f"steps{3 * n}": {"VKV_GRAD_STEPS": str(n)} for n in (3, 4, 5, 6, 7, 9, 10, 12, 14)}, "walk_nosurf": {"VKV_GRAD_NOSURF": "1"}, "walk_s2": {"VKV_GRAD_STEPS": "2"}, "walk_s4": {"VKV_GRAD_STEPS": "4"},
            "walk_s8": {"VKV_GRAD_STEPS": "8"}, "walk_s16": {"VKV_GRAD_STEPS": "16"}, "walk_s32": {"VKV_GRAD_STEPS": "32"},
            "dbg1": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "1"}, "dbg2": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "2"}, "dbg3": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "3"}, "dbg4": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "4"}, "dbg6": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "6"}, "dbg7": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "7"}, "dbg16": {"VKV_GRAD_DBG": "16"}, "dbg32": {"VKV_GRAD_DBG": "32"}, "dbg48": {"VKV_GRAD_DBG": "48"}, "dbg9": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "9"}, "dbg11": {"VKV_GRAD_STEPS": "4", "VKV_GRAD_DBG": "11"},
            "v1_int_surf": {"VKV_GRAD_V1": "1"}, "flat_nosurf": {"VKV_GRAD_NOSURF": "1", "VKV_GRAD_FLAT": "1"}, "flat_surf": {"VKV_GRAD_FLAT": "1"}}
only = os.environ.get("GRAD_PROBE_ONLY")
for name, env in VARIANTS.items():
    if only and name not in only.split(","):
        continue
    for k in ("VKV_GRAD_FP32", "VKV_GRAD_NOSURF", "VKV_GRAD_V1", "VKV_GRAD_CTAS", "VKV_GRAD_FLAT", "VKV_GRAD_STEPS", "VKV_GRAD_DBG"):
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
