#!/usr/bin/env python
"""Run under compute-sanitizer (GPU box): one small invocation of every kernel family that stages data through shared memory with
mbarriers / bulk copies, of the ray caster (march + long-ray pass) and of the rest of the TF-change path.  Results are compared with
the oracle so that a tool-induced timing change that exposed a race would also show up as a wrong answer."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle_api as orc
from vkvolume_b200 import capi, scene
from vkvolume_b200.capi import RenderOptions, VolumeOptions

ctx = capi.Context(0)
ok = True
def check(name, cond):
    global ok
    print(("ok   " if cond else "FAIL ") + name, flush=True)
    ok = ok and bool(cond)

# (0) gradient column walk on a volume full of exact ties (per-warp tie queue in shared memory, passes split around a row of the walk,
# indexed shuffles between the lanes of a 256-byte run, rows wider than a run)
V = (128 + np.random.default_rng(3).integers(-2, 3, size=(9, 31, 528))).astype(np.uint8)
vol = capi.Volume(ctx, 528, 31, 9)
vol.upload(V)
vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
Gref = orc.gradient_map(V, True)
check("gradient map (column walk, ties) == oracle", np.array_equal(vol.download_gradient(), Gref))
check("gradient texture array == oracle", np.array_equal(vol.download_gradient_texture(), Gref))
vol.close()

# (1) occupancy_tma_kernel with and without the gradient map on a 1024-wide volume; fused count
W, H, D = 1024, 24, 16
V = np.random.default_rng(1).integers(0, 256, size=(D, H, W), dtype=np.uint8)
vol = capi.Volume(ctx, W, H, D)
vol.upload(V)
for tfo in (dict(intensity_min=0.5, intensity_max=1.0, gradient_min=0.1, gradient_max=0.6), dict(intensity_min=0.9, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)):
    opt = VolumeOptions(**tfo)
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    G = vol.download_gradient()
    check("gradient map == oracle", np.array_equal(G, orc.gradient_map(V, True)))
    n = vol.update_transfer_function(opt, capi.SKIP_BLOCK, count=True)
    tf = orc.transfer_function_texture(opt)
    check(f"occupancy (gradient {'on' if tfu.use_gradient else 'off'}) == oracle", np.array_equal(vol.download_distance_map(0), orc.occupancy_map(V, G, tf, 4, bool(tfu.use_gradient))))
    check("count == oracle", n == orc.occupied_voxel_count(V, G, tfu))
vol.close()

# (2) distance maps: y sweep over rows of three strips (edge exchange through shared memory every 16 rows), z walk with its two
# walker warps meeting in the middle (named barrier), all 8 octant maps
O = np.where(np.random.default_rng(2).random((72, 100, 544)) < 0.0005, 0, 255).astype(np.uint8)
vol = capi.Volume(ctx, 544, 100, 72, block_size=1)
vol.upload(np.where(O == 0, 255, 0).astype(np.uint8))
opt = VolumeOptions(intensity_min=0.5, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
tfu = capi.transfer_function_uniform(opt)
vol.update_transfer_function_texture(opt)
vol.compute_distance_map(tfu, capi.SKIP_DISTANCE)
check("isotropic distance map == oracle", np.array_equal(vol.download_distance_map(0), orc.distance_map(O)))
vol.compute_distance_map(tfu, capi.SKIP_ANISOTROPIC_DISTANCE)
want = orc.distance_map_anisotropic(O)
check("8 octant maps == oracle", all(np.array_equal(vol.download_distance_map(i), want[i]) for i in range(8)))
vol.close()

# (3) ray caster: march + long-ray pass (hand-over after 6 trips so that most rays go through the queue), tile history on the second frame
shape = (40, 48, 64)
D, H, W = shape
V = scene.blobs_volume(shape, seed=7)
opt = VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2)
tfu = capi.transfer_function_uniform(opt)
vol = capi.Volume(ctx, W, H, D)
vol.upload(V)
vol.compute_gradient_map(tfu)
vol.update_transfer_function(opt, capi.SKIP_DISTANCE)
G = orc.gradient_map(V); tf = orc.transfer_function_texture(opt)
Dm = orc.distance_map(orc.occupancy_map(V, G, tf, 4, True))
it = scene.image_transform((0.004,) * 3, (W, H, D))
width, height = 256, 192
cu, ru = vol.make_uniforms(scene.look_at_camera((30, 20, 44), aspect=width / height), it, 5.0)
os.environ["VKV_RC_LONG_T"] = "6"; os.environ["VKV_RC_LONG_ALWAYS"] = "1"
for filt in (capi.FILTER_HARDWARE, capi.FILTER_EXACT):
    ropt = RenderOptions(skipping_type=capi.SKIP_DISTANCE, clip_distance=5.0, filter=filt)
    for rep in range(2):
        img, counts = vol.render_to_host(cu, ru, tfu, ropt, width, height)
    ref, rc, _, _ = orc.render(V, G, tf, Dm, vol.map_extent, cu, ru, tfu, ropt, width, height)
    d = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=2)
    check(f"frame (filter {filt}) within the bar of the oracle", (d <= 1).mean() >= 0.999 and counts.covered_pixels == rc.covered_pixels)
vol.close()
print("SANITIZE_PROBE_" + ("OK" if ok else "FAILED"), flush=True)
sys.exit(0 if ok else 1)
