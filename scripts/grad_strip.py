#!/usr/bin/env python
"""GPU-box probe: gradient map on a big volume, slice order vs strips of rows (VKV_GRAD_STRIP), byte-checked on z-crops against the oracle."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import bench
from vkvolume_b200 import capi
from vkvolume_b200.capi import VolumeOptions
name = sys.argv[1] if len(sys.argv) > 1 else "c5s"
wl = bench.WORKLOADS[name]
W, H, D = wl["dim"]
ctx = capi.Context(0)
vol = capi.Volume(ctx, W, H, D, block_size=4)
capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), 0)
vol.upload_device(vol.device_voxels(), 0)
tfu = capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
class P:
    def __init__(s, p, n): s.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 3}
G = torch.as_tensor(P(vol.device_gradient(), W * H * D), device="cuda").view(D, H, W)
V = torch.as_tensor(P(vol.device_voxels(), W * H * D), device="cuda").view(D, H, W)
ref = None
for strip in sys.argv[2:] or ["0", "16", "32", "64", "128"]:
    os.environ["VKV_GRAD_STRIP"] = strip
    os.environ["VKV_GRAD_V1"] = "1"
    ts = []
    for _ in range(3):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); vol.compute_gradient_map(tfu, 0); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ck = int(G[::7, ::5].to(torch.int64).sum().item())
    if ref is None:
        ref = ck
        import oracle_api as orc
        z0 = D // 2
        want = orc.gradient_map(V[z0 - 1:z0 + 5].cpu().numpy(), True)[1:5]
        ok = np.array_equal(G[z0:z0 + 4].cpu().numpy(), want)
        print("   crop vs oracle:", ok)
    print(f"{name} strip {strip}: {np.median(ts):.3f} ms  ({2 * W * H * D / np.median(ts) / 1e6:.0f} GB/s algorithmic)  checksum {'same' if ck == ref else 'DIFFERENT'}", flush=True)
