#!/usr/bin/env python
"""GPU-box probe (one GPU): what ONE rank of an N-way tile-sharded 8K frame costs, without any peer — vkv_render_tiles with
tile_first = 0, tile_stride = N into local memory, N = 1, 2, 4, 8, under the ray caster's A/B knobs.  Separates the part of the
tile-mode scaling loss that is local (long-ray pass, launch overheads) from the part that is the gather over NVLink."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from vkvolume_b200 import capi, scene  # noqa: E402
from vkvolume_b200.capi import RenderOptions, VolumeOptions  # noqa: E402

wl = bench.WORKLOADS[os.environ.get("TILES_PROBE_WORKLOAD", "c5s")]
W, H, D = wl["dim"]
FW, FH = wl["frame"]
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), 0)
vol.upload_device(vol.device_voxels(), 0)
opt = VolumeOptions(**wl["tf"])
tfu = capi.transfer_function_uniform(opt)
vol.compute_gradient_map(tfu, 0)
vol.update_transfer_function(opt, wl["skip"])
it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
ropt = RenderOptions(skipping_type=wl["skip"], clip_distance=wl["clip"], early_ray_termination=1)
fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
views = [vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(k, 72, wl), aspect=FW / FH), it, wl["clip"]) for k in range(13)]
VARIANTS = {"default": {}, "no_handover": {"VKV_RC_LONG_T": "0"}, "handover_always": {"VKV_RC_LONG_ALWAYS": "1"}, "no_history": {"VKV_RC_NO_HISTORY": "1"}}
for t in (96, 112, 128, 144, 160, 192):
    VARIANTS[f"always_T{t}"] = {"VKV_RC_LONG_ALWAYS": "1", "VKV_RC_LONG_T": str(t)}
only = os.environ.get("TILES_PROBE_ONLY")
for name, env in VARIANTS.items():
    if only and name not in only.split(","):
        continue
    for k in ("VKV_RC_LONG_T", "VKV_RC_LONG_ALWAYS", "VKV_RC_NO_HISTORY"):
        os.environ.pop(k, None)
    os.environ.update(env)
    row = []
    for n in (1, 2, 4, 8):
        per_first = []
        for first in sorted({0, n // 2, n - 1}):
            ts = []
            for k, (cu, ru) in enumerate(views):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                vol.render_tiles(cu, ru, tfu, ropt, FW, FH, bench.TILE_W, bench.TILE_H, first, n, fb.data_ptr(), 0, 0, 0)
                b.record()
                b.synchronize()
                if k >= 3:
                    ts.append(a.elapsed_time(b))
            per_first.append(float(np.median(ts)))
        row.append((n, max(per_first)))
    t1 = row[0][1]
    print(name, "  ".join(f"N={n}: {t:.4f} ms (x{t1 / t:.2f})" for n, t in row), flush=True)
