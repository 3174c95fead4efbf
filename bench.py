#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the VkVolume hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c2|c1|c3|c4|c5]

Metric (BASELINE.json): ray-cast Msamples/s (plus ms/frame and the TF-change ESS rebuild ms as
extra keys) on config 2 — a synthetic stag-beetle-shaped 832x832x494 volume, 1920x1080 frame,
ESS = Chebyshev distance map, ERT on — at N = 1.  A "step" is one ray-cast frame (the camera
orbits the volume, one view per step).  `value` times the frame with everything resident in
HBM; `e2e` goes through vkv_render_to_host (uniforms from host structs, RGBA8 frame and
counters copied back to pinned host memory inside the timed region).

N > 1 (launched under torchrun, one rank per GPU, NCCL): every rank holds a full replica and
stores its pixels straight into rank 0's HBM through a CUDA-IPC peer mapping (the gather is fused
into the ray caster's epilogue, no collective on the data path).  Two decompositions
(--parallelism): `frames` — a step is N consecutive views of the orbit, one per rank, each landing
in its slot of a frame ring on rank 0 (independent units, weak scaling; the default for the
1080p-class workloads, whose 0.13 ms frame is bound by its longest rays and cannot be cut N ways);
`tiles` — one frame cut into 64x32 tiles dealt round-robin to the ranks (strong scaling; the
default for the 8K workloads c5 / c5s, BASELINE.json config 5).  The TF-change rebuild shards the
O(N) occupancy pass by z-slabs and all-gathers the slab rows of the occupancy map over NCCL.

--impl reference times the reference algorithm on the host CPU cores (the oracle port: the
reference has no CPU path and cannot be built here) on the same config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (W, H, D, synth kind, seed, voxel size, axis-angle, TF options, frame, skip mode, clip, eye scale)
    "c1": dict(dim=(256, 256, 256), kind=0, seed=0x5EED0001, voxel=(0.004, 0.004, 0.004), axis_angle=(1, 0, 0, 0),
               tf=dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2), frame=(512, 512), skip=2, clip=5.0,
               name="blobs256: 256^3 u8 blobs, 512x512, ESS distance, ERT on"),
    "c2": dict(dim=(832, 832, 494), kind=1, seed=0x5EED0002, voxel=(0.001, 0.001, 0.001), axis_angle=(1, 0, 0, 90),
               tf=dict(intensity_min=0.086, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0), frame=(1920, 1080), skip=2, clip=5.0,
               name="beetle832: synthetic stag-beetle-shaped 832x832x494 u8, 1920x1080, ESS distance map, ERT on"),
    "c3": dict(dim=(1024, 1024, 795), kind=2, seed=0x5EED0003, voxel=(0.0003, 0.0003, 0.0007), axis_angle=(1, 0, 0, 90),
               tf=dict(intensity_min=0.2, intensity_max=0.8, gradient_min=0.0, gradient_max=0.0), frame=(1920, 1080), skip=3, clip=2.0,
               inside=True, name="aniso1024: 1024x1024x795 anisotropic voxels, camera inside + clip plane, anisotropic distance maps"),
    "c4": dict(dim=(1024, 1024, 1024), kind=0, seed=0x5EED0004, voxel=(0.001, 0.001, 0.001), axis_angle=(1, 0, 0, 0),
               tf=dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.05, gradient_max=0.25), frame=(1920, 1080), skip=2, clip=5.0,
               name="tf_sweep1024: 1024^3 blobs, TF sweep"),
    "c5": dict(dim=(4096, 4096, 2048), kind=3, seed=0x5EED0005, voxel=(0.00025, 0.00025, 0.00025), axis_angle=(1, 0, 0, 0),
               tf=dict(intensity_min=0.15, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0), frame=(7680, 4320), skip=2, clip=5.0,
               name="big4096: 4096x4096x2048 u8 (34 GB), 7680x4320, ESS distance, ERT on"),
    "c5s": dict(dim=(2048, 2048, 1024), kind=3, seed=0x5EED0005, voxel=(0.0005, 0.0005, 0.0005), axis_angle=(1, 0, 0, 0),
                tf=dict(intensity_min=0.15, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0), frame=(7680, 4320), skip=2, clip=5.0,
                name="big2048: 2048x2048x1024 u8 (4.3 GB) stand-in for config 5, 7680x4320"),
}
METRIC = "raycast_msamples_per_s"
UNIT = "Msamples/s"
TILE_W, TILE_H = 64, 32


def orbit_eye(step: int, n: int, wl) -> tuple:
    """Camera positions on a tilted circle around the volume (world units; node scale 100)."""
    W, H, D = wl["dim"]
    size = 100.0 * max(v * e for v, e in zip(wl["voxel"], wl["dim"]))
    ang = 2.0 * math.pi * (step % n) / n + 0.35
    if wl.get("inside"):
        r = 0.12 * size
        return (r * math.cos(ang), 0.05 * size * math.sin(2 * ang), r * math.sin(ang))
    r = 1.25 * size
    return (r * math.cos(ang), 0.45 * size + 0.15 * size * math.sin(2 * ang), r * math.sin(ang))


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.stop_flag = period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


class _DevPtr:
    """Wraps a raw device pointer as a torch tensor via __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ======================================================================================================
# native arm
# ======================================================================================================
def run_native(args):
    import torch
    import torch.distributed as dist

    from vkvolume_b200 import capi, scene
    from vkvolume_b200.capi import RenderOptions, SampleCounts, VolumeOptions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    wl = WORKLOADS[args.workload]
    W, H, D = wl["dim"]
    FW, FH = wl["frame"]
    K, Wm = args.steps, args.warmup
    par = args.parallelism
    if par == "auto":
        par = "tiles" if FW * FH >= 16_000_000 else "frames"
    frames_mode = world > 1 and par == "frames"
    slots = world if frames_mode else 1        # frames held in rank 0's HBM per step

    ctx = capi.Context(local_rank)
    vol = capi.Volume(ctx, W, H, D, block_size=4)
    capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), stream)
    vol.upload_device(vol.device_voxels(), stream)
    opt = VolumeOptions(**wl["tf"])
    tfu = capi.transfer_function_uniform(opt)
    skip = wl["skip"]
    N_vox, M_blk = vol.n_voxels, vol.n_blocks
    use_g = bool(tfu.use_gradient)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=5, flush=None):
        ts = []
        for _ in range(reps):
            if flush is not None:
                flush()
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return ts

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)        # 256 MiB write > 126 MB L2

    # ---- one-off: gradient map (K1) ------------------------------------------------------------------
    t_grad = timed(lambda: vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)), stream), reps=3, flush=flush_l2)
    vol.compute_gradient_map(tfu, stream)        # the map the TF actually asks for (all 255 when gradients are off)

    # ---- TF-change rebuild (K2a [+K2b] + K3) ------------------------------------------------------------
    def rebuild_single(count=False, o=opt):
        return vol.update_transfer_function(o, skip, count=count, stream=stream)

    from vkvolume_b200 import sharding
    slab = sharding.slab_size(vol.map_extent[2], world)
    if world > 1:
        Wb, Hb, Db = vol.map_extent
        map_idx = 7 if skip == 3 else 0
        vol.set_number_of_distance_maps(8 if skip == 3 else 1)
        gather_buf = torch.empty(world * slab * Wb * Hb, dtype=torch.uint8, device=dev)
        count_t = torch.zeros(1, dtype=torch.int64, device=dev)

    def rebuild_sharded(o=opt):
        """z-slab occupancy on every rank -> all-gather of the slab rows -> local distance transform."""
        Wb, Hb, Db = vol.map_extent
        vol.update_transfer_function_texture(o, stream)
        u = capi.transfer_function_uniform(o)
        z0, zc = sharding.slab_range(rank, world, Db)
        count_t.zero_()
        vol.compute_occupancy_slab(u, skip, z0, zc, count_dev=count_t.data_ptr(), stream=stream)
        full = torch.as_tensor(_DevPtr(vol.device_distance_map(map_idx), Wb * Hb * Db), device=dev)
        sharding.all_gather_occupancy(full, rank, world, (Wb, Hb, Db), gather_buf)
        sharding.all_reduce_count(count_t)
        vol.compute_distance_from_occupancy(skip, stream)

    rebuild = rebuild_sharded if world > 1 else rebuild_single
    rebuild()
    torch.cuda.synchronize()
    # sweep: 100 TF changes in config 4's pattern (imin = base + 0.004 k), fewer when asked to be quick
    n_changes = args.tf_changes
    rebuild_ms = []
    for k in range(n_changes):
        o = VolumeOptions(**{**wl["tf"], "intensity_min": wl["tf"]["intensity_min"] + 0.0004 * (k % 25)})
        flush_l2()
        a, b = ev(), ev()
        a.record()
        rebuild(o=o) if world > 1 else rebuild_single(False, o)
        b.record()
        b.synchronize()
        rebuild_ms.append(a.elapsed_time(b))
    rebuild()        # back to the workload's TF
    count_ms = timed(lambda: vol.compute_occupied_voxel_count(tfu, stream), reps=3, flush=flush_l2) if world == 1 else [float("nan")]
    occupied = vol.compute_occupied_voxel_count(tfu, stream)
    # stage breakdown of one rebuild (single-GPU only)
    stage_ms = {}
    if world == 1:
        stage_ms["tf_texture_and_masks"] = float(np.median(timed(lambda: vol.update_transfer_function_texture(opt, stream), 5)))
        stage_ms["occupancy"] = float(np.median(timed(lambda: vol.compute_occupancy_slab(tfu, skip, 0, vol.map_extent[2], stream=stream), 5, flush_l2)))
        stage_ms["distance"] = float(np.median(timed(lambda: vol.compute_distance_from_occupancy(skip, stream), 5)))
        vol.compute_occupancy_slab(tfu, skip, 0, vol.map_extent[2], stream=stream)
        vol.compute_distance_from_occupancy(skip, stream)

    # ---- frame buffer + peer mapping -------------------------------------------------------------------
    it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
    ropt = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1)
    fb_ptr, peer_ptr = None, None
    frame_bytes = FW * FH * 4
    if rank == 0:
        fb_all = torch.zeros((slots, FH, FW, 4), dtype=torch.uint8, device=dev)
        fb = fb_all[0]
        fb_ptr = fb_all.data_ptr()
    if world > 1:
        import ctypes as C
        handle = [None]
        if rank == 0:
            hbuf = (C.c_uint8 * capi.IPC_HANDLE_BYTES)()
            capi.check(capi.lib().vkv_ipc_export(C.c_void_p(fb_ptr), hbuf))
            handle = [bytes(hbuf)]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            hbuf = (C.c_uint8 * capi.IPC_HANDLE_BYTES).from_buffer_copy(handle[0])
            p = C.c_void_p()
            capi.check(capi.lib().vkv_ipc_open(hbuf, C.byref(p)))
            peer_ptr = p.value
            fb_ptr = peer_ptr + (sharding.frame_slot_offset(rank, FW, FH) if frames_mode else 0)        # this rank's slot of rank 0's frame ring
    counts_t = torch.zeros(4, dtype=torch.int64, device=dev)

    cams = [scene.look_at_camera(orbit_eye(k, 72, wl), aspect=FW / FH) for k in range(72)]        # synthetic inputs: the camera path

    def uniforms(step):
        return vol.make_uniforms(cams[step % 72], it, wl["clip"])        # host maths of VolumeRenderSubpass::draw (C ABI call)

    def render_step(step, counts_ptr):
        if frames_mode:        # N consecutive views per step, one per rank, each into its slot on rank 0
            cu, ru = uniforms(sharding.view_of_rank(step, rank, world))
            vol.render(cu, ru, tfu, ropt, FW, FH, fb_ptr, 0, counts_ptr, stream)
            return
        cu, ru = uniforms(step)
        if world == 1:
            vol.render(cu, ru, tfu, ropt, FW, FH, fb_ptr, 0, counts_ptr, stream)
        else:
            vol.render_tiles(cu, ru, tfu, ropt, FW, FH, TILE_W, TILE_H, rank, world, fb_ptr, 0, counts_ptr, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region: W warm-up + K steps ------------------------------------------------------------------
    for s in range(Wm):
        render_step(s, 0)
    barrier()
    # untimed self-check of the fused gather (N > 1): the frame the ranks assembled in rank 0's HBM through their peer
    # stores must equal rank 0's own full-frame render of the same view, byte for byte
    tiles_match = None
    if world > 1:
        fb_all.zero_() if rank == 0 else None
        barrier()
        render_step(0, 0)
        barrier()
        if rank == 0:
            full = torch.zeros_like(fb)
            tiles_match = True
            for r in range(slots):        # frames mode: slot r holds view r, rendered by rank r
                cu0, ru0 = uniforms(r)
                vol.render(cu0, ru0, tfu, ropt, FW, FH, full.data_ptr(), 0, 0, stream)
                torch.cuda.synchronize()
                tiles_match = tiles_match and bool(torch.equal(full, fb_all[r]))
            del full
        barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = capi.kernel_launch_count()
    starts, stops = [ev() for _ in range(K)], [ev() for _ in range(K)]
    counts_t.zero_()
    barrier()
    wall0 = time.perf_counter()
    for s in range(K):
        flush_l2()        # not timed: evict the previous frame's working set
        starts[s].record()
        render_step(s, counts_t.data_ptr())
        stops[s].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = capi.kernel_launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in zip(starts, stops)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)        # a step is done when the slowest rank is
        dist.all_reduce(counts_t)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())        # whole job
    total_ms = float(step_ms.sum().item())
    n_vol, n_dist, n_empty, n_cov = [int(x) for x in counts_t.tolist()]
    samples = n_vol + n_dist
    value = samples / (total_ms * 1e-3) / 1e6
    ms_per_step = total_ms / K

    # ---- per-mode frame times (config 2: ESS none vs occupancy vs distance) — N = 1 only ---------------------
    modes = {}
    if world == 1 and not args.quick:
        for mode, mname in ((0, "none"), (1, "block_occupancy"), (2, "distance"), (3, "anisotropic_distance")):
            vol.update_transfer_function(opt, mode, stream=stream)
            mopt = RenderOptions(skipping_type=mode, clip_distance=wl["clip"], early_ray_termination=1)
            counts_t.zero_()
            ts = []
            nviews = 12 if mode == 0 else 24
            for s in range(nviews):
                cu, ru = uniforms(s * 3)
                flush_l2()
                a, b = ev(), ev()
                a.record()
                vol.render(cu, ru, tfu, mopt, FW, FH, fb_ptr, 0, counts_t.data_ptr(), stream)
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            c = counts_t.tolist()
            modes[mname] = {"ms_per_frame": float(np.mean(ts)), "msamples_per_s": (c[0] + c[1]) / (sum(ts) * 1e-3) / 1e6,
                            "samples_per_frame": (c[0] + c[1]) / nviews}
        vol.update_transfer_function(opt, skip, stream=stream)

    # ---- e2e through the host-buffer C-ABI call (rank 0's view at N = 1; tiles + gather at N > 1) ---------------
    e2e = None
    if world == 1:
        import ctypes as C
        h2d = C.sizeof(capi.CameraUniform) + C.sizeof(capi.RayCastUniform) + C.sizeof(capi.TransferFunctionUniform) + C.sizeof(capi.RenderOptions)
        d2h = FW * FH * 4 + C.sizeof(SampleCounts)
        ke = max(10, K // 4)
        # (a) one synchronous call per frame: vkv_render_to_host
        host_fb = torch.empty((FH, FW, 4), dtype=torch.uint8).pin_memory()
        host_np = host_fb.numpy()
        for s in range(3):
            cu, ru = uniforms(s)
            vol.render_to_host(cu, ru, tfu, ropt, FW, FH, out=host_np, stream=stream)
        e_samples = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(ke):
            cu, ru = uniforms(s)        # host maths of VolumeRenderSubpass::draw is inside the timed region
            _, c = vol.render_to_host(cu, ru, tfu, ropt, FW, FH, out=host_np, stream=stream)
            e_samples += c.volume_samples + c.distance_samples
        t1 = time.perf_counter()
        e2e_sync = {"value": e_samples / (t1 - t0) / 1e6, "unit": UNIT, "ms_per_frame": (t1 - t0) * 1e3 / ke, "steps": ke,
                    "note": "vkv_render_to_host: one blocking call per frame (kernel, then RGBA8 frame + counters D2H, then stream sync)"}
        # (b) the pipelined call a frame sequence uses: vkv_render_to_host_async, frame k leaves over PCIe while frame k + 1 is cast
        ring = torch.empty((3, FH, FW, 4), dtype=torch.uint8).pin_memory()
        cnt = torch.zeros((ke, 4), dtype=torch.int64).pin_memory()
        for s in range(3):
            cu, ru = uniforms(s)
            vol.render_to_host_async(cu, ru, tfu, ropt, FW, FH, ring[s % 3].data_ptr(), cnt[s].data_ptr(), stream)
        vol.render_to_host_wait(stream)
        cnt.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(ke):
            cu, ru = uniforms(s)        # host maths of VolumeRenderSubpass::draw is inside the timed region
            vol.render_to_host_async(cu, ru, tfu, ropt, FW, FH, ring[s % 3].data_ptr(), cnt[s].data_ptr(), stream)
        vol.render_to_host_wait(stream)        # every frame and every counter block is in pinned host memory
        t1 = time.perf_counter()
        a_samples = int(cnt[:, 0].sum().item() + cnt[:, 1].sum().item())
        assert a_samples == e_samples, "pipelined and blocking e2e paths disagree on the samples taken"
        e2e = {"value": a_samples / (t1 - t0) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_frame": (t1 - t0) * 1e3 / ke, "steps": ke, "blocking_call": e2e_sync,
               "note": "vkv_render_to_host_async per frame + one vkv_render_to_host_wait: uniforms from host structs (kernel parameters), every RGBA8 frame + "
                       "counters copied D2H to pinned memory inside the timed region, the copy of frame k overlapping the casting of frame k+1; no L2 flush"}
    elif frames_mode:
        # frames: every rank delivers its own views to page-locked host memory over its own PCIe link (vkv_render_to_host_async);
        # nothing has to pass through rank 0 on the way to the host
        ke = max(10, K // 4)
        ring = torch.empty((3, FH, FW, 4), dtype=torch.uint8).pin_memory()
        cnt = torch.zeros((ke, 4), dtype=torch.int64).pin_memory()
        for s in range(3):
            cu, ru = uniforms(sharding.view_of_rank(s, rank, world))
            vol.render_to_host_async(cu, ru, tfu, ropt, FW, FH, ring[s % 3].data_ptr(), cnt[s].data_ptr(), stream)
        vol.render_to_host_wait(stream)
        cnt.zero_()
        barrier()
        t0 = time.perf_counter()
        for s in range(ke):
            cu, ru = uniforms(sharding.view_of_rank(s, rank, world))
            vol.render_to_host_async(cu, ru, tfu, ropt, FW, FH, ring[s % 3].data_ptr(), cnt[s].data_ptr(), stream)
        vol.render_to_host_wait(stream)
        t1 = time.perf_counter()
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        ns = torch.tensor([int(cnt[:, 0].sum().item() + cnt[:, 1].sum().item())], dtype=torch.int64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ns)
        e2e = {"value": ns.item() / tt.item() / 1e6, "unit": UNIT, "h2d_bytes_per_step": 460 * world, "d2h_bytes_per_step": world * (FW * FH * 4 + 32),
               "ms_per_frame": tt.item() * 1e3 / ke / world, "ms_per_step": tt.item() * 1e3 / ke, "steps": ke,
               "note": "every rank: vkv_render_to_host_async per view + one wait; each frame + counters copied D2H to that rank's pinned host memory over its own PCIe link"}
    else:
        # tiles: the frame lands in rank 0's HBM through the peer stores; rank 0 then copies it to pinned host memory
        if rank == 0:
            host_fb = torch.empty((slots, FH, FW, 4), dtype=torch.uint8).pin_memory()
        ke = max(10, K // 4)
        counts_t.zero_()
        barrier()
        t0 = time.perf_counter()
        for s in range(ke):
            render_step(s, counts_t.data_ptr())
            barrier()
            if rank == 0:
                host_fb.copy_(fb_all, non_blocking=False)
        t1 = time.perf_counter()
        dist.all_reduce(counts_t)
        c = counts_t.tolist()
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": (c[0] + c[1]) / tt.item() / 1e6, "unit": UNIT, "h2d_bytes_per_step": 480 * world, "d2h_bytes_per_step": FW * FH * 4,
               "ms_per_frame": tt.item() * 1e3 / ke, "steps": ke,
               "note": "tiles rendered on all ranks with peer stores into rank 0, barrier, rank 0 copies the frame to pinned host memory"}

    # ---- roofline --------------------------------------------------------------------------------------------------
    hbm_peak, peak_src = measured_peaks()
    gamma = 1 if use_g else 0
    texel_bytes_per_step = (n_vol * (8 * (1 + gamma) + 4) + n_dist) / K / max(world, 1)        # per launch: one frame (frames) or one rank's tiles
    tex_peak = None
    if rank == 0:
        try:
            tex_peak = {"l2_resident_256": capi.bench_tex3d(ctx, 256, 256, True), "hbm_resident_768": capi.bench_tex3d(ctx, 768, 256, True),
                        "random_256": capi.bench_tex3d(ctx, 256, 128, False)}
        except Exception as e:        # the microbenchmark must not take the bench line down
            tex_peak = {"error": str(e)}
    roofline = None
    if rank == 0:
        peak_fetch = tex_peak.get("l2_resident_256") if isinstance(tex_peak, dict) else None
        achieved = texel_bytes_per_step / (ms_per_step * 1e-3) / 1e9
        peak = peak_fetch * 8 / 1e9 if peak_fetch else None
        traffic = None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists() and args.workload == "c2" and world == 1:
            traffic = json.loads(tj.read_text()).get("raycast_kernel")        # dram bytes per launch from the committed ncu --set full capture
        roofline = {"kernel": "raycast_kernel", "bound": "texture", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if peak else None, "traffic": traffic,
                    "peak_source": "vkv_bench_tex3d measured live: coherent trilinear u8 tex3D fetches/s on an L2-resident 256^3 array x 8 texel bytes per fetch",
                    "algorithmic_bytes_per_launch": texel_bytes_per_step,
                    "definition": "n_vol*(8*(1+gamma)+4)+n_dist texel bytes per frame (SURVEY 8(d))"}
    hbm_rooflines = {}
    if world == 1:
        def rl(bytes_, ms):
            ach = bytes_ / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "algorithmic_bytes_per_launch": bytes_, "ms": ms, "peak_source": peak_src}
        hbm_rooflines["gradient_flat_kernel"] = rl(2 * N_vox, float(np.median(t_grad)))
        hbm_rooflines["gradient_flat_kernel"]["note"] = "2 B/voxel algorithmic; the kernel also writes the map into the texture array (+1 B/voxel)"
        hbm_rooflines["occupancy_tma_kernel"] = rl((2 if use_g else 1) * N_vox + M_blk, stage_ms["occupancy"])
        hbm_rooflines["occupancy_tma_kernel+count"] = rl((2 if use_g else 1) * N_vox, float(np.median(count_ms)))
        hbm_rooflines["occupancy_tma_kernel+count"]["note"] = "vkv_compute_occupied_voxel_count: memset + kernel + 8-byte D2H + stream sync inside the timed region"
        hbm_rooflines["distance_map_passes"] = rl((28 if skip == 3 else 6) * M_blk, stage_ms["distance"])
        hbm_rooflines["distance_map_passes"]["note"] = f"{M_blk} blocks: the map is L2-resident, the HBM fraction is not the meaningful figure (DESIGN.md)"

    # ---- CPU baseline (oracle port on the host cores; bounded sample) ------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_from_device(vol, wl, opt, tfu, ropt, it, FW, FH, skip, uniforms)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if (world > 1 and not frames_mode) else "weak", "vs_baseline": None, "dtype": "u8 voxels / f32 march",
            "data": "synthetic", "impl": "native",
            "config": {"workload": wl["name"], "volume": [W, H, D], "frame": [FW, FH], "block_size": 4, "ess": ["none", "block", "distance", "anisotropic"][skip],
                       "ert": True, "tf": wl["tf"], "views": "72-view orbit, one view per step",
                       "l2": "256 MiB flush write between timed steps (untimed); volume 342 MB > 126 MB L2" if args.workload == "c2" else "256 MiB flush write between timed steps (untimed)",
                       "parallelism": (f"frames: a step is {world} consecutive orbit views, one per rank, each stored into its slot of rank 0's frame ring through peer stores"
                                       if frames_mode else f"image tiles {TILE_W}x{TILE_H} round-robin over {world} ranks, peer stores into rank 0") if world > 1 else "single GPU",
                       "frames_per_step": slots},
            "ms_per_frame": ms_per_step, "samples_per_frame": samples / K / slots, "volume_samples_per_frame": n_vol / K / slots,
            "distance_samples_per_frame": n_dist / K / slots, "covered_pixels_per_frame": n_cov / K / slots,
            "mpixels_per_s": slots * FW * FH / (ms_per_step * 1e-3) / 1e6,
            "ess_rebuild_ms": {"median": float(np.median(rebuild_ms)), "mean": float(np.mean(rebuild_ms)), "p95": float(np.percentile(rebuild_ms, 95)),
                               "changes": n_changes, "stages": stage_ms, "sharded_z_slabs": world > 1},
            "occupied_voxels": occupied, "occupied_percent": 100.0 * occupied / N_vox,
            "modes": modes, "tiles_match_single_gpu_frame": tiles_match, "e2e": e2e, "gpu_launches": int(launches), "wall_s_timed_region": wall,
            "clocks": sampler.summary(), "roofline": roofline, "rooflines_hbm": hbm_rooflines, "tex3d_fetch_per_s": tex_peak,
            "cpu_baseline": cpu_baseline,
        }
        emit(line)
    if world > 1:
        if peer_ptr:
            capi.check(capi.lib().vkv_ipc_close(__import__("ctypes").c_void_p(peer_ptr)))
        dist.barrier()
        dist.destroy_process_group()


# ======================================================================================================
# CPU legs (the only places that may touch oracle/)
# ======================================================================================================
def _oracle():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_api as orc
    return orc


def cpu_render_sample(orc, V, G, tf, maps, dim_b, tfu, ropt, FW, FH, uniforms_fn, budget_s=12.0):
    """Times the oracle's ray caster on whole views of the orbit (or a centred band of rows when one
    view alone exceeds the budget) until about budget_s of CPU work has been done."""
    cu, ru = uniforms_fn(0)
    t0 = time.perf_counter()
    orc.render(V, G, tf, maps, dim_b, cu, ru, tfu, ropt, FW, FH, y_first=FH // 2 - 4, y_count=8)
    per8 = time.perf_counter() - t0
    rows = int(min(FH, max(8, 8 * budget_s / max(per8, 1e-4))))
    y0 = max(0, FH // 2 - rows // 2)
    n, dt, views = 0, 0.0, 0
    while dt < budget_s and views < 720:
        cu, ru = uniforms_fn(views)
        t0 = time.perf_counter()
        _, c, _, _ = orc.render(V, G, tf, maps, dim_b, cu, ru, tfu, ropt, FW, FH, y_first=y0, y_count=rows)
        dt += time.perf_counter() - t0
        n += c.volume_samples + c.distance_samples
        views += 1
    return n / dt / 1e6, dt, rows, n, views


def cpu_baseline_from_device(vol, wl, opt, tfu, ropt, it, FW, FH, skip, uniforms_fn):
    orc = _oracle()
    V = vol.download_voxels()
    G = vol.download_gradient() if tfu.use_gradient else None
    tf = vol.download_transfer_function()
    maps = None
    if skip == 3:
        maps = np.stack([vol.download_distance_map(i) for i in range(8)])
    elif skip != 0:
        maps = vol.download_distance_map(0)
    v, dt, rows, n, views = cpu_render_sample(orc, V, G, tf, maps, vol.map_extent, tfu, ropt, FW, FH, uniforms_fn)
    return {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"oracle ray caster (OpenMP, {orc.num_threads()} threads) on {rows} of {FH} rows of {views} orbit views of the same workload: {n} samples in {dt:.2f} s"}


def run_reference(args):
    """The reference algorithm on the host CPU: synthetic volume -> TF texture -> occupancy -> distance map -> ray-cast band."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = _oracle()
    from vkvolume_b200 import scene
    from vkvolume_b200.capi import RenderOptions, VolumeOptions
    wl = WORKLOADS[args.workload]
    W, H, D = wl["dim"]
    FW, FH = wl["frame"]
    K, Wm = args.steps, args.warmup
    t_setup = time.perf_counter()
    V = orc.synth_volume(wl["kind"], wl["seed"], W, H, D)
    opt = VolumeOptions(**wl["tf"])
    tfu = orc.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    G = orc.gradient_map(V, True) if tfu.use_gradient else None
    t0 = time.perf_counter()
    O = orc.occupancy_map(V, G, tf, 4, bool(tfu.use_gradient))
    t_occ = time.perf_counter() - t0
    skip = wl["skip"]
    t0 = time.perf_counter()
    maps = {0: None, 1: O}.get(skip)
    if skip == 2:
        maps = orc.distance_map(O)
    elif skip == 3:
        maps = orc.distance_map_anisotropic(O)
    t_dist = time.perf_counter() - t0
    dim_b, _ = orc.map_extent((W, H, D), 4)
    it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
    ropt = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1)
    t_setup = time.perf_counter() - t_setup

    def uniforms(step):
        cam = scene.look_at_camera(orbit_eye(step, 72, wl), aspect=FW / FH)
        return orc.make_uniforms((W, H, D), dim_b, cam, it, wl["clip"])

    # a step = a bounded band of rows of the frame of view `step` (about 1.5 s of CPU each)
    cu, ru = uniforms(0)
    t0 = time.perf_counter()
    orc.render(V, G, tf, maps, dim_b, cu, ru, tfu, ropt, FW, FH, y_first=FH // 2 - 4, y_count=8)
    per8 = time.perf_counter() - t0
    rows = int(min(FH, max(8, 8 * 1.5 / max(per8, 1e-3))))
    y0 = max(0, FH // 2 - rows // 2)
    samples, elapsed = 0, 0.0
    for s in range(Wm + K):
        cu, ru = uniforms(s)
        t0 = time.perf_counter()
        _, c, _, _ = orc.render(V, G, tf, maps, dim_b, cu, ru, tfu, ropt, FW, FH, y_first=y0, y_count=rows)
        dt = time.perf_counter() - t0
        if s >= Wm:
            samples += c.volume_samples + c.distance_samples
            elapsed += dt
    value = samples / elapsed / 1e6
    sample = f"oracle port, OpenMP {orc.num_threads()} threads: rows [{y0},{y0 + rows}) of each {FW}x{FH} view, {K} views"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K, "warmup": Wm,
            "ms_per_step": elapsed * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 voxels / f32 march",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": wl["name"], "volume": [W, H, D], "frame": [FW, FH], "block_size": 4,
                       "ess": ["none", "block", "distance", "anisotropic"][skip], "ert": True, "tf": wl["tf"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ess_rebuild_ms": {"occupancy": t_occ * 1e3, "distance": t_dist * 1e3}, "setup_s": t_setup,
            "note": "the reference (Vulkan/GLSL) has no CPU path and cannot be built in this image; this is the oracle port of its shaders on the host cores"}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries the one JSON line and nothing else: everything any library writes to fd 1 from here on (NCCL prints its
    version banner there at communicator creation) goes to stderr; emit() writes the line to the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--parallelism", default="auto", choices=["auto", "tiles", "frames"], help="N > 1: one view per rank (weak) or one frame cut into tiles (strong)")
    ap.add_argument("--tf-changes", type=int, default=100)
    ap.add_argument("--quick", action="store_true", help="skip the per-mode table")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 6 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.steps = 360 if args.steps is None else args.steps
        args.warmup = 10 if args.warmup is None else max(3, args.warmup)
        run_native(args)


if __name__ == "__main__":
    main()
