"""CPU-only coverage of the N > 1 host logic: slab / tile arithmetic, and a world_size-2 gloo run of the z-slab occupancy
exchange (per-rank slab computed with the oracle, slab rows all-gathered, count all-reduced, distance transform on the
gathered map) checked against the single-rank result."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from vkvolume_b200 import sharding

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("Db,world", [(7, 2), (8, 2), (1, 2), (3, 4), (124, 8), (512, 8), (5, 8)])
def test_slabs_partition_the_map(Db, world):
    covered = []
    for r in range(world):
        z0, zc = sharding.slab_range(r, world, Db)
        assert 0 <= z0 <= Db and 0 <= zc <= sharding.slab_size(Db, world)
        covered += list(range(z0, z0 + zc))
    assert covered == list(range(Db))


@pytest.mark.parametrize("frame,world", [((1920, 1080), 2), ((7680, 4320), 8), ((160, 120), 4), ((64, 32), 8)])
def test_tiles_are_dealt_exactly_once(frame, world):
    n = sharding.n_tiles(*frame, 64, 32)
    seen = sorted(t for r in range(world) for t in sharding.tiles_of_rank(r, world, n))
    assert seen == list(range(n))
    sizes = [len(sharding.tiles_of_rank(r, world, n)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world,steps", [(2, 5), (8, 9), (1, 4)])
def test_frames_are_dealt_exactly_once_and_slots_do_not_overlap(world, steps):
    views = sorted(sharding.view_of_rank(s, r, world) for s in range(steps) for r in range(world))
    assert views == list(range(steps * world))
    w, h = 1920, 1080
    offs = [sharding.frame_slot_offset(r, w, h) for r in range(world)]
    assert offs == [r * w * h * 4 for r in range(world)] and len(set(offs)) == world


FRAMES_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["VKV_ROOT"]); sys.path.insert(0, os.path.join(os.environ["VKV_ROOT"], "tests"))
import oracle_api as orc
from vkvolume_b200 import scene, sharding
from vkvolume_b200.capi import RenderOptions, VolumeOptions
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# frames decomposition with the oracle standing in for the ray caster: each rank renders its view of every step into its slot;
# the ring gathered on rank 0 must equal rank 0's own render of views step*world .. step*world + world - 1
shape = (20, 24, 28); D, H, W = shape; fw, fh = 48, 32
V = scene.blobs_volume(shape, seed=9, n_blobs=5)
opt = VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
tfu = orc.transfer_function_uniform(opt); tf = orc.transfer_function_texture(opt)
(dim_b), _ = orc.map_extent((W, H, D), 4)
it = scene.image_transform((0.004,) * 3, (W, H, D))
ropt = RenderOptions(skipping_type=0, clip_distance=2.0)
def frame(view):
    cam = scene.look_at_camera((14 + 2 * view, 9, 20 - view), aspect=fw / fh)
    cu, ru = orc.make_uniforms((W, H, D), dim_b, cam, it, 2.0)
    return orc.render(V, None, tf, None, dim_b, cu, ru, tfu, ropt, fw, fh)[0]
for step in range(2):
    mine = torch.from_numpy(frame(sharding.view_of_rank(step, rank, world)).reshape(-1).copy())
    ring = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(ring, mine)        # stands in for the peer stores into rank 0's ring
    if rank == 0:
        flat = torch.cat(ring).numpy()
        for r in range(world):
            o = sharding.frame_slot_offset(r, fw, fh)
            assert np.array_equal(flat[o:o + fw * fh * 4].reshape(fh, fw, 4), frame(step * world + r)), (step, r)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["VKV_ROOT"]); sys.path.insert(0, os.path.join(os.environ["VKV_ROOT"], "tests"))
import oracle_api as orc
from vkvolume_b200 import scene, sharding
from vkvolume_b200.capi import VolumeOptions
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
shape = (26, 20, 24)        # D, H, W: 7 block slices of 4 (ragged last block), 2 ranks -> slabs of 4 and 3
V = scene.blobs_volume(shape, seed=5, n_blobs=6)
opt = VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2)
tfu = orc.transfer_function_uniform(opt); tf = orc.transfer_function_texture(opt)
G = orc.gradient_map(V)
D, H, W = shape
(Wb, Hb, Db), _ = orc.map_extent((W, H, D), 4)
O_full = orc.occupancy_map(V, G, tf, 4, True)
z0, zc = sharding.slab_range(rank, world, Db)
# this rank's slab only: the voxel slices [4 z0, 4 (z0 + zc)) of the volume
vz0, vz1 = 4 * z0, min(D, 4 * (z0 + zc))
mine = orc.occupancy_map(V[vz0:vz1], G[vz0:vz1], tf, 4, True) if zc else np.zeros((0, Hb, Wb), np.uint8)
full = torch.full((Wb * Hb * Db,), 77, dtype=torch.uint8)        # poison: every cell must be overwritten
full[z0 * Wb * Hb:(z0 + zc) * Wb * Hb] = torch.from_numpy(mine.reshape(-1).copy())
sharding.all_gather_occupancy(full, rank, world, (Wb, Hb, Db))
cnt = torch.tensor([orc.occupied_voxel_count(V[vz0:vz1], G[vz0:vz1], tfu) if zc else 0], dtype=torch.int64)
sharding.all_reduce_count(cnt)
got = full.numpy().reshape(Db, Hb, Wb)
assert np.array_equal(got, O_full), "gathered occupancy map differs from the single-rank map"
assert int(cnt.item()) == orc.occupied_voxel_count(V, G, tfu), "all-reduced voxel count differs"
assert np.array_equal(orc.distance_map(got.copy()), orc.distance_map(O_full.copy()))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


K3_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["VKV_ROOT"]); sys.path.insert(0, os.path.join(os.environ["VKV_ROOT"], "tests"))
import oracle_api as orc
from vkvolume_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# The sharded isotropic distance-map build of csrc/group.cu, with numpy standing in for the three 1-D passes:
# x and y passes on the rank's z-slab, all-gather of the xy-intermediate slabs, z pass on the rank's block rows, all-gather of
# the rows — must equal the oracle's transform of the whole map (shaders/distance_map.comp:44-109).
def line_pass(h, axis):
    """D(t) = min(255, min_j max(|j - t|, h(j))) along `axis` (the operator of every pass; on 0/255 input it is the x sweep)."""
    h = np.moveaxis(h.astype(np.int32), axis, -1)
    L = h.shape[-1]
    out = h.copy()
    for n in range(1, min(L, 255)):
        lo = np.full_like(h, 255); hi = np.full_like(h, 255)
        lo[..., n:] = h[..., :-n]; hi[..., :-n] = h[..., n:]
        out = np.minimum(out, np.maximum(n, np.minimum(lo, hi)))
    return np.moveaxis(np.minimum(out, 255), -1, axis).astype(np.uint8)
Db, Hb, Wb = 13, 10, 12        # 13 slices over 2 ranks: slabs of 7 and 6; 10 rows: 5 and 5
rng = np.random.default_rng(17)
O = np.where(rng.random((Db, Hb, Wb)) < 0.01, 0, 255).astype(np.uint8)
want = orc.distance_map(O.copy())
z0, zc = sharding.slab_range(rank, world, Db)
xy = line_pass(line_pass(O[z0:z0 + zc], 2), 1)                     # x then y, slice-local
s = sharding.slab_size(Db, world)
mine = torch.full((s * Hb * Wb,), 255, dtype=torch.uint8)
mine[: zc * Hb * Wb] = torch.from_numpy(xy.reshape(-1).copy())
parts = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)
full_xy = np.concatenate([parts[r].numpy()[: sharding.slab_range(r, world, Db)[1] * Hb * Wb].reshape(-1, Hb, Wb) for r in range(world)])
assert full_xy.shape == (Db, Hb, Wb)
y0, yc = sharding.row_range(rank, world, Hb)
rows = line_pass(full_xy[:, y0:y0 + yc, :], 0)                      # z pass on this rank's block rows
r_s = sharding.slab_size(Hb, world)
mine = torch.full((Db * r_s * Wb,), 255, dtype=torch.uint8)
mine[: Db * yc * Wb] = torch.from_numpy(rows.reshape(-1).copy())
parts = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)
got = np.concatenate([parts[r].numpy()[: Db * sharding.row_range(r, world, Hb)[1] * Wb].reshape(Db, -1, Wb) for r in range(world)], axis=1)
assert np.array_equal(got, want), "sharded x/y-on-slabs + z-on-rows build differs from the whole-map transform"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''



TIMES_WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["VKV_ROOT"])
from vkvolume_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# rank 0 is slow on even steps, rank 1 on odd ones: per-rank totals 6.0 and 7.0, per-step maxima sum to 9.0
mine = torch.tensor([2.0, 1.0, 2.0, 1.0] if rank == 0 else [1.0, 2.5, 1.0, 2.5], dtype=torch.float64)
tot, lock = sharding.reduce_step_times(mine, independent=True)
assert tot == 7.0 and lock == 9.0, (tot, lock)
tot, lock = sharding.reduce_step_times(mine, independent=False)
assert tot == 9.0 and lock is None, (tot, lock)
assert mine.tolist() == ([2.0, 1.0, 2.0, 1.0] if rank == 0 else [1.0, 2.5, 1.0, 2.5])        # input untouched
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''

def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("worker", ["zslab", "frames", "sharded_k3", "step_times"])
def test_exchange_world_size_2_gloo(tmp_path, worker):
    script = tmp_path / "worker.py"
    script.write_text({"zslab": WORKER, "frames": FRAMES_WORKER, "sharded_k3": K3_WORKER, "step_times": TIMES_WORKER}[worker])
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VKV_ROOT=str(ROOT),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out}"
        assert f"rank {rank} ok" in out
