"""Multi-GPU parity on real devices (pytest -m gpu; skipped when the box has a single GPU).

Two (and four) ranks under torchrun + NCCL, one per GPU, exactly as bench.py --gpus N runs them:
* image-tile sharded ray casting with the gather fused into the kernel's epilogue (peer stores into rank 0's
  framebuffer through a CUDA-IPC mapping) must reproduce the single-GPU frame byte for byte;
* the z-slab sharded TF-change rebuild (slab occupancy + count, all-gather of the slab rows, u64 all-reduce, local
  distance transform) must reproduce the single-GPU occupancy-derived distance maps and the voxel count bit for bit.
"""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["VKV_ROOT"])
from vkvolume_b200 import capi, scene, sharding
from vkvolume_b200.capi import RenderOptions, VolumeOptions

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["VKV_GROUP_ALWAYS"] = "1"        # the library rebuilds volumes this small on every replica; the sharded path is what is under test
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream().cuda_stream
ctx = capi.Context(lr)
W, H, D = 208, 160, 118            # 52 x 40 x 30 blocks (ragged last slice): slabs of 15 block slices (2 ranks) or 8, 8, 8, 6 (4 ranks)
FW, FH = 640, 360
failures = []
for skip, tf in ((capi.SKIP_DISTANCE, dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2)),
                 (capi.SKIP_ANISOTROPIC_DISTANCE, dict(intensity_min=0.3, intensity_max=0.9, gradient_min=0.0, gradient_max=0.0)),
                 (capi.SKIP_BLOCK, dict(intensity_min=0.2, intensity_max=1.0, gradient_min=0.05, gradient_max=0.25))):
    vol = capi.Volume(ctx, W, H, D, block_size=4)
    capi.synth_volume(ctx, 0, 0x5EED0001, W, H, D, vol.device_voxels(), stream)
    vol.upload_device(vol.device_voxels(), stream)
    opt = VolumeOptions(**tf)
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(tfu, stream)
    Wb, Hb, Db = vol.map_extent
    n_maps = 8 if skip == capi.SKIP_ANISOTROPIC_DISTANCE else 1
    # single-GPU answer (every rank computes it: the replicas are identical)
    n_single = vol.update_transfer_function(opt, skip, count=True, stream=stream)
    single = [vol.download_distance_map(i).copy() for i in range(n_maps)]
    # sharded rebuild
    vol.set_number_of_distance_maps(n_maps)
    map_idx = n_maps - 1
    vol.update_transfer_function_texture(opt, stream)
    z0, zc = sharding.slab_range(rank, world, Db)
    count_t = torch.zeros(1, dtype=torch.int64, device=dev)
    full = torch.as_tensor(type("P", (), {"__cuda_array_interface__": {"shape": (Wb * Hb * Db,), "typestr": "|u1",
                           "data": (vol.device_distance_map(map_idx), False), "version": 3}})(), device=dev)
    full.fill_(77)                  # poison: every cell must be rewritten by a slab
    vol.compute_occupancy_slab(tfu, skip, z0, zc, count_dev=count_t.data_ptr(), stream=stream)
    sharding.all_gather_occupancy(full, rank, world, (Wb, Hb, Db))
    sharding.all_reduce_count(count_t)
    vol.compute_distance_from_occupancy(skip, stream)
    torch.cuda.synchronize()
    if int(count_t.item()) != n_single:
        failures.append(f"skip {skip}: sharded count {int(count_t.item())} != {n_single}")
    for i in range(n_maps):
        if not np.array_equal(vol.download_distance_map(i), single[i]):
            failures.append(f"skip {skip}: sharded map {i} differs from the single-GPU map")
    # the same rebuild through the library's own group: exchanges by peer copies over NVLink, barriers in peer memory, no NCCL
    # on the data path (vkv_update_transfer_function_sharded); twice in a row, with and without the count
    handles = [None] * world
    dist.all_gather_object(handles, vol.group_export())
    vol.group_open(rank, world, handles)
    for rep, want_count in enumerate((True, False, True)):
        torch.as_tensor(type("P", (), {"__cuda_array_interface__": {"shape": (Wb * Hb * Db,), "typestr": "|u1",
                        "data": (vol.device_distance_map(0), False), "version": 3}})(), device=dev).fill_(77)
        torch.cuda.synchronize()
        dist.barrier()
        n_grp = vol.update_transfer_function_sharded(opt, skip, count=want_count, stream=stream)
        torch.cuda.synchronize()
        if want_count and n_grp != n_single:
            failures.append(f"skip {skip}: group count {n_grp} != {n_single} (rep {rep})")
        for i in range(n_maps):
            if not np.array_equal(vol.download_distance_map(i), single[i]):
                failures.append(f"skip {skip}: group-sharded map {i} differs from the single-GPU map (rep {rep})")
    vol.group_close()
    dist.barrier()
    # tile-sharded frame with peer stores into rank 0
    it = scene.image_transform((0.004,) * 3, (W, H, D))
    cu, ru = vol.make_uniforms(scene.look_at_camera((95, 60, 130), aspect=FW / FH), it, 5.0)
    ropt = RenderOptions(skipping_type=skip, clip_distance=5.0)
    ref = torch.zeros((FH, FW, 4), dtype=torch.uint8, device=dev)
    vol.render(cu, ru, tfu, ropt, FW, FH, ref.data_ptr(), stream=stream)
    fb = torch.full((FH, FW, 4), 9, dtype=torch.uint8, device=dev) if rank == 0 else None
    handle = [None]
    if rank == 0:
        hbuf = (C.c_uint8 * capi.IPC_HANDLE_BYTES)()
        capi.check(capi.lib().vkv_ipc_export(C.c_void_p(fb.data_ptr()), hbuf))
        handle = [bytes(hbuf)]
    dist.broadcast_object_list(handle, src=0)
    peer = None
    if rank == 0:
        target = fb.data_ptr()
    else:
        hbuf = (C.c_uint8 * capi.IPC_HANDLE_BYTES).from_buffer_copy(handle[0])
        p = C.c_void_p()
        capi.check(capi.lib().vkv_ipc_open(hbuf, C.byref(p)))
        peer = target = p.value
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    ref_counts = torch.zeros(4, dtype=torch.int64, device=dev)
    vol.render(cu, ru, tfu, ropt, FW, FH, ref.data_ptr(), 0, ref_counts.data_ptr(), stream)
    vol.render_tiles(cu, ru, tfu, ropt, FW, FH, 64, 32, rank, world, target, 0, counts.data_ptr(), stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.all_reduce(counts)
    if rank == 0:
        if not torch.equal(fb, ref):
            failures.append(f"skip {skip}: tile-sharded frame differs in {(fb != ref).any(dim=2).sum().item()} pixels")
        if counts.tolist() != ref_counts.tolist():
            failures.append(f"skip {skip}: sample counters {counts.tolist()} != {ref_counts.tolist()}")
    dist.barrier()
    if peer:
        capi.check(capi.lib().vkv_ipc_close(C.c_void_p(peer)))
    vol.close()
flag = torch.tensor([len(failures)], device=dev)
dist.all_reduce(flag)
if failures:
    print(f"[rank {rank}] FAIL: " + "; ".join(failures), flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and flag.item() == 0:
    print("MULTIGPU_OK", flush=True)
sys.exit(1 if flag.item() else 0)
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_tiles_and_slabs_match_single_gpu(tmp_path, world):
    """world = 4: 30 block slices in slabs of 8, 8, 8, 6 (ragged last slab); 40 block rows in shares of 10."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, VKV_ROOT=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(worker)]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "MULTIGPU_OK" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]
