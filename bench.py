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
    # (kind 3, the noise-modulated blobs: the plain blobs of kind 0 are so smooth at 1024^3 that the sweep's gradient window shows nothing)
    "c4": dict(dim=(1024, 1024, 1024), kind=3, seed=0x5EED0004, voxel=(0.001, 0.001, 0.001), axis_angle=(1, 0, 0, 0),
               tf=dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.05, gradient_max=0.25), frame=(1920, 1080), skip=2, clip=5.0,
               name="tf_sweep1024: 1024^3 noise-modulated blobs, TF sweep"),
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


def sweep_options(wl, k: int):
    """TF setting k of BASELINE.json config 4's sweep: imin = 0.05 + 0.004 k, the gradient window alternating
    between off (0, 0) and (0.05, 0.25)."""
    from vkvolume_b200.capi import VolumeOptions
    gmin, gmax = ((0.0, 0.0), (0.05, 0.25))[k & 1]
    return VolumeOptions(intensity_min=0.05 + 0.004 * k, intensity_max=1.0, gradient_min=gmin, gradient_max=gmax)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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


def tiles_8k_series(ctx, torch, dist, capi, scene, sharding, rank, world, dev, stream, flush_l2, barrier, ev, name="c5s", frames=10):
    """N > 1, every workload whose headline is not already the tile decomposition: BASELINE.json config 5's sharding on record —
    one 7680x4320 frame cut into 64x32 tiles dealt round-robin to the ranks (full replica per rank, peer stores into rank 0's
    frame), beside rank 0 casting the whole frame alone, and a byte-for-byte comparison of the two frames.  `c5s` (2048x2048x1024,
    same generator and TF as config 5) keeps the set-up to a second; `c5` itself runs through --workload c5."""
    import ctypes as C
    from vkvolume_b200.capi import RenderOptions, VolumeOptions
    wl = WORKLOADS[name]
    W, H, D = wl["dim"]
    FW, FH = wl["frame"]
    vol = capi.Volume(ctx, W, H, D, block_size=4)
    capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels(), stream)
    vol.upload_device(vol.device_voxels(), stream)
    opt = VolumeOptions(**wl["tf"])
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(tfu, stream)
    vol.update_transfer_function(opt, wl["skip"], stream=stream)
    it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
    ropt = RenderOptions(skipping_type=wl["skip"], clip_distance=wl["clip"], early_ray_termination=1)
    fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device=dev) if rank == 0 else None
    own = torch.zeros((FH, FW, 4), dtype=torch.uint8, device=dev)        # private frame: the single-GPU series
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
    views = [vol.make_uniforms(scene.look_at_camera(orbit_eye(k, 72, wl), aspect=FW / FH), it, wl["clip"]) for k in range(frames + 3)]

    def series(fn):
        ts = []
        for k in range(frames + 3):
            flush_l2()
            a, b = ev(), ev()
            a.record()
            fn(k)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        t = torch.tensor(ts[3:], dtype=torch.float64, device=dev)
        return t

    barrier()
    t_n = series(lambda k: vol.render_tiles(views[k][0], views[k][1], tfu, ropt, FW, FH, TILE_W, TILE_H, rank, world, target, 0, 0, stream))
    dist.all_reduce(t_n, op=dist.ReduceOp.MAX)        # a frame is done when the slowest rank's tiles are
    barrier()
    t_1 = series(lambda k: vol.render(views[k][0], views[k][1], tfu, ropt, FW, FH, own.data_ptr(), 0, 0, stream))
    dist.all_reduce(t_1, op=dist.ReduceOp.MAX)        # the slowest of N independent single-GPU runs (they run concurrently on one box)
    # frame check on one view
    if rank == 0:
        fb.zero_()
    barrier()
    vol.render_tiles(views[5][0], views[5][1], tfu, ropt, FW, FH, TILE_W, TILE_H, rank, world, target, 0, 0, stream)
    barrier()
    match = None
    if rank == 0:
        vol.render(views[5][0], views[5][1], tfu, ropt, FW, FH, own.data_ptr(), 0, 0, stream)
        torch.cuda.synchronize()
        match = bool(torch.equal(own, fb))
    barrier()
    if peer:
        capi.check(capi.lib().vkv_ipc_close(C.c_void_p(peer)))
    barrier()
    vol.close()
    ms_n, ms_1 = float(t_n.median().item()), float(t_1.median().item())
    return {"workload": wl["name"], "frame": [FW, FH], "tiles": [TILE_W, TILE_H], "frames_timed": frames, "ms_per_frame_N": ms_n, "ms_per_frame_1": ms_1,
            "speedup": ms_1 / ms_n, "frame_matches": match, "n_gpus": world,
            "timing": "CUDA events per frame on every rank, max over ranks per frame, median over frames; 256 MiB L2 flush before each frame (untimed)"}


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
    # (computed with use_gradient = true whatever the workload's TF says — SURVEY A.8.1: harnesses pass true — so that the TF sweep
    # below can switch the gradient window on; kernels ignore G while tfu.use_gradient is 0)
    t_grad = timed(lambda: vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)), stream), reps=3, flush=flush_l2)

    # ---- TF-change rebuild (K2a [+K2b] + K3) ------------------------------------------------------------
    def rebuild_single(count=False, o=opt):
        return vol.update_transfer_function(o, skip, count=count, stream=stream)

    from vkvolume_b200 import sharding
    slab = sharding.slab_size(vol.map_extent[2], world)
    if world > 1:
        Wb, Hb, Db = vol.map_extent
        vol.set_number_of_distance_maps(8 if skip == 3 else 1)
        # the library's own group: peer mappings of every rank's map / xy-intermediate / signal block (CUDA IPC)
        handles = [None] * world
        dist.all_gather_object(handles, vol.group_export())
        vol.group_open(rank, world, handles)
    last_sharded_count = [None]

    def rebuild_sharded(o=opt, count=False):
        """vkv_update_transfer_function_sharded: z-slab occupancy (+ count), x/y distance passes on the slab, exchange of the slabs over
        NVLink peer memory, z pass on the rank's block rows, exchange of the rows; barriers in peer memory, no NCCL on the data path."""
        last_sharded_count[0] = vol.update_transfer_function_sharded(o, skip, count=count, stream=stream)

    rebuild = rebuild_sharded if world > 1 else rebuild_single
    rebuild()
    torch.cuda.synchronize()
    # untimed self-check (N > 1): the z-slab rebuild must reproduce this rank's own single-GPU maps and the voxel count, bit for bit
    sharded_rebuild_matches = None
    if world > 1:
        Wb, Hb, Db = vol.map_extent
        n_maps = 8 if skip == 3 else 1
        rebuild_sharded(count=True)
        torch.cuda.synchronize()
        got = [torch.as_tensor(_DevPtr(vol.device_distance_map(i), Wb * Hb * Db), device=dev).clone() for i in range(n_maps)]
        got_count = int(last_sharded_count[0])
        want_count = rebuild_single(True)
        ok = got_count == want_count
        for i in range(n_maps):
            ok = ok and bool(torch.equal(got[i], torch.as_tensor(_DevPtr(vol.device_distance_map(i), Wb * Hb * Db), device=dev)))
        okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        sharded_rebuild_matches = bool(okt.item())
        del got
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
        occ = lambda: vol.compute_occupancy_slab(tfu, skip, 0, vol.map_extent[2], stream=stream)
        stage_ms["tf_texture_and_masks"] = float(np.median(timed(lambda: vol.update_transfer_function_texture(opt, stream), 5)))
        stage_ms["occupancy"] = float(np.median(timed(occ, 5, flush_l2)))
        # the same kernel timed as a train of 8 launches between one pair of events when its input cannot stay in the L2 (every launch
        # streams V [and G] from HBM again): takes the ~10 us of event + call overhead per launch out of a 60-90 us kernel
        bytes_in = (2 if use_g else 1) * N_vox
        if bytes_in > 2 * 126e6:
            occ()
            a, b = ev(), ev()
            a.record()
            for _ in range(8):
                occ()
            b.record()
            b.synchronize()
            stage_ms["occupancy_back_to_back"] = a.elapsed_time(b) / 8
        # the distance transform consumes (overwrites) the occupancy map: rebuild it, untimed, before every timed repetition
        stage_ms["distance"] = float(np.median(timed(lambda: vol.compute_distance_from_occupancy(skip, stream), 5, flush=lambda: (occ(), flush_l2()))))
        occ()
        vol.compute_distance_from_occupancy(skip, stream)

    # ---- BASELINE.json config 4's sweep, as specified: every change = TF texture + occupancy + voxel count
    # (host-synchronous, as VolumeRender::update_transfer_function logs it) + distance map, then one frame ----------------------
    tf_sweep = None
    if world == 1 and args.tf_changes > 0:
        n_sw = args.tf_changes
        sweep_fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device=dev)
        it_sw = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
        ropt_sw = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1)
        cam_sw = scene.look_at_camera(orbit_eye(0, 72, wl), aspect=FW / FH)
        cu_sw, ru_sw = vol.make_uniforms(cam_sw, it_sw, wl["clip"])
        reb, ren, occ_pct = [], [], []
        for k in range(n_sw):
            o = sweep_options(wl, k)
            u = capi.transfer_function_uniform(o)
            flush_l2()
            a, b, c = ev(), ev(), ev()
            a.record()
            n_occ = vol.update_transfer_function(o, skip, count=True, stream=stream)
            b.record()
            vol.render(cu_sw, ru_sw, u, ropt_sw, FW, FH, sweep_fb.data_ptr(), 0, 0, stream)
            c.record()
            c.synchronize()
            reb.append(a.elapsed_time(b))
            ren.append(b.elapsed_time(c))
            occ_pct.append(100.0 * n_occ / N_vox)
        tf_sweep = {"changes": n_sw, "pattern": "imin = 0.05 + 0.004 k, (gmin, gmax) alternating (0, 0) / (0.05, 0.25); each change: TF texture + occupancy + "
                                                "voxel count (host-synchronous read-back) + distance map, then one frame of the headline view",
                    "rebuild_ms": {"median": float(np.median(reb)), "mean": float(np.mean(reb)), "p95": float(np.percentile(reb, 95)),
                                   "median_gradient_off": float(np.median(reb[0::2])), "median_gradient_on": float(np.median(reb[1::2])) if n_sw > 1 else None},
                    "render_ms": {"median": float(np.median(ren)), "mean": float(np.mean(ren))},
                    "occupied_percent_first_last": [occ_pct[0], occ_pct[-1]], "l2": "256 MiB flush write before every change (untimed)"}
        del sweep_fb
        rebuild()        # back to the workload's TF

    # ---- frame buffer + peer mapping -------------------------------------------------------------------
    it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
    ropt = RenderOptions(skipping_type=skip, clip_distance=wl["clip"], early_ray_termination=1)
    fb_ptr, peer_ptr = None, None
    frame_bytes = FW * FH * 4
    # Frames mode (one independent view per rank per step) keeps every rank's frame in that rank's own HBM, exactly as at N = 1: the
    # decomposition has no exchange step and none is invented.  --gather-frames restores round 1's variant (every rank stores its
    # pixels into its slot of a frame ring in rank 0's HBM through peer stores).  Tiles mode always gathers into rank 0 (north star).
    local_frames = frames_mode and not args.gather_frames
    if rank == 0 or local_frames:
        fb_all = torch.zeros((1 if local_frames else slots, FH, FW, 4), dtype=torch.uint8, device=dev)
        fb = fb_all[0]
        fb_ptr = fb_all.data_ptr()
    if world > 1 and not local_frames:
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
        fb_all.zero_() if (rank == 0 or local_frames) else None
        barrier()
        render_step(0, 0)
        barrier()
        got = fb_all if rank == 0 else None
        if local_frames:        # collect every rank's frame of step 0 on rank 0 (NCCL, untimed) for the comparison
            parts = [torch.empty_like(fb_all) for _ in range(world)] if rank == 0 else None
            dist.gather(fb_all, parts, dst=0)
            got = torch.cat(parts, dim=0) if rank == 0 else None
        if rank == 0:
            full = torch.zeros_like(fb)
            tiles_match = True
            for r in range(slots):        # frames mode: slot r holds view r, rendered by rank r
                cu0, ru0 = uniforms(r)
                vol.render(cu0, ru0, tfu, ropt, FW, FH, full.data_ptr(), 0, 0, stream)
                torch.cuda.synchronize()
                tiles_match = tiles_match and bool(torch.equal(full, got[r]))
            del full, got
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
    lockstep_ms = None
    if world > 1:
        # Independent frames (one view per rank per step, no exchange between ranks, no barrier inside the timed region): the job is
        # done when the slowest rank has rendered its K frames -> the contract's "time K steps, max over ranks".  The stricter figure —
        # every step waiting for its slowest view, as if a barrier followed each frame — is reported beside it.  Tiles of one frame:
        # a step is done when the slowest rank is.
        tot_ms, lock_tot = sharding.reduce_step_times(step_ms, frames_mode)
        lockstep_ms = lock_tot / K if lock_tot is not None else None
        step_ms = torch.full_like(step_ms, tot_ms / K)
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
                            "samples_per_frame": (c[0] + c[1]) / nviews, "volume_samples_per_frame": c[0] / nviews}
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

    # ---- driver-visible tile-sharded 8K series (N > 1; not when the headline already is that decomposition) -----------------
    tiles_8k = None
    if world > 1 and (frames_mode or FW * FH < 16_000_000) and not args.no_tiles_8k:
        tiles_8k = tiles_8k_series(ctx, torch, dist, capi, scene, sharding, rank, world, dev, stream, flush_l2, barrier, ev)

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
        # Units: the denominator is the FILTERED-FETCH rate of the texture pipe (vkv_bench_tex3d, measured live), so the numerator
        # counts only what goes through that pipe: n_vol * (1 + gamma) filtered fetches of 8 texels each.  The TF-table read (one
        # 16-byte read-only load per volume sample) and the skip-map bytes (one byte load per distance sample) use the LSU path and
        # are reported beside it, as is SURVEY 8(d)'s mixed byte figure for continuity with round 1.
        peak_fetch = tex_peak.get("l2_resident_256") if isinstance(tex_peak, dict) else None
        launches_per_step = max(world, 1)
        fetches_per_launch = n_vol * (1 + gamma) / K / launches_per_step
        t_launch = ms_per_step * 1e-3
        fetch_rate = fetches_per_launch / t_launch
        achieved = fetch_rate * 8 / 1e9
        peak = peak_fetch * 8 / 1e9 if peak_fetch else None
        survey_bytes = texel_bytes_per_step
        traffic = None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists() and args.workload == "c2" and world == 1:
            traffic = json.loads(tj.read_text()).get("raycast_kernel")        # dram bytes per launch from the committed ncu --set full capture
        roofline = {"kernel": "raycast_kernel", "bound": "texture", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if peak else None, "traffic": traffic,
                    "peak_source": "vkv_bench_tex3d measured live: coherent trilinear u8 tex3D fetches/s on an L2-resident 256^3 array x 8 texel bytes per fetch",
                    "algorithmic_bytes_per_launch": fetches_per_launch * 8,
                    "definition": "filtered texel bytes = n_vol*(1+gamma) trilinear fetches x 8 texels x 1 B per frame, over the frame time; same fraction as fetches/s over the fetch peak",
                    "filtered_fetches_per_launch": fetches_per_launch, "achieved_fetches_per_s": fetch_rate, "peak_fetches_per_s": peak_fetch,
                    "frac_fetch": (fetch_rate / peak_fetch) if peak_fetch else None,
                    "lsu_side": {"tf_table_bytes_per_launch": n_vol * 16 / K / launches_per_step, "skip_map_bytes_per_launch": n_dist / K / launches_per_step},
                    "survey_8d_mixed_bytes": {"bytes_per_launch": survey_bytes, "achieved": survey_bytes / t_launch / 1e9,
                                              "frac_of_texture_byte_peak": (survey_bytes / t_launch / 1e9 / peak) if peak else None,
                                              "note": "n_vol*(8*(1+gamma)+4)+n_dist: round 1's figure; mixes TF and skip-map bytes into a texture-pipe denominator"},
                    "regime": "latency-bound with ESS + ERT on (a few dependent fetches per ray); modes.none is the throughput-bound configuration"}
        if modes.get("none"):
            mn = modes["none"]
            fr = mn["volume_samples_per_frame"] * (1 + gamma) / (mn["ms_per_frame"] * 1e-3)
            roofline["modes_none"] = {"achieved_fetches_per_s": fr, "frac_fetch": (fr / peak_fetch) if peak_fetch else None}
    hbm_rooflines = {}
    if world == 1:
        def rl(bytes_, ms):
            ach = bytes_ / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "algorithmic_bytes_per_launch": bytes_, "ms": ms, "peak_source": peak_src}
        hbm_rooflines["gradient_walk_kernel"] = rl(2 * N_vox, float(np.median(t_grad)))
        hbm_rooflines["gradient_walk_kernel"]["note"] = "2 B/voxel algorithmic; the kernel also writes the map into the texture array (+1 B/voxel)"
        hbm_rooflines["occupancy_tma_kernel"] = rl((2 if use_g else 1) * N_vox + M_blk, stage_ms["occupancy"])
        hbm_rooflines["occupancy_tma_kernel+count"] = rl((2 if use_g else 1) * N_vox, float(np.median(count_ms)))
        hbm_rooflines["occupancy_tma_kernel+count"]["note"] = "vkv_compute_occupied_voxel_count: memset + kernel + 8-byte D2H + stream sync inside the timed region"
        hbm_rooflines["distance_map_passes"] = rl((28 if skip == 3 else 6) * M_blk, stage_ms["distance"])
        hbm_rooflines["distance_map_passes"]["note"] = f"{M_blk} blocks: the map is L2-resident, the HBM fraction is not the meaningful figure (DESIGN.md)"

    # ---- CPU baseline (oracle port on the host cores; bounded sample) ------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def gpu_frame(view):
            cu, ru = uniforms(view)
            vol.render(cu, ru, tfu, ropt, FW, FH, fb_ptr, 0, 0, stream)
            torch.cuda.synchronize()
            return fb.cpu().numpy()
        cpu_baseline = cpu_baseline_from_device(vol, wl, opt, tfu, ropt, it, FW, FH, skip, uniforms, gpu_frame)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if (world > 1 and not frames_mode) else "weak", "vs_baseline": None, "dtype": "u8 voxels / f32 march",
            "data": "synthetic", "impl": "native",
            "config": {"workload": wl["name"], "volume": [W, H, D], "frame": [FW, FH], "block_size": 4, "ess": ["none", "block", "distance", "anisotropic"][skip],
                       "ert": True, "tf": wl["tf"], "views": "72-view orbit, one view per step",
                       "l2": "256 MiB flush write between timed steps (untimed); volume 342 MB > 126 MB L2" if args.workload == "c2" else "256 MiB flush write between timed steps (untimed)",
                       "parallelism": ((f"frames: a step is {world} consecutive orbit views, one per rank, each into the rank's own HBM as at N = 1 (independent frames: no exchange step)"
                                        if local_frames else
                                        f"frames: a step is {world} consecutive orbit views, one per rank, each stored into its slot of rank 0's frame ring through peer stores")
                                       if frames_mode else f"image tiles {TILE_W}x{TILE_H} round-robin over {world} ranks, peer stores into rank 0") if world > 1 else "single GPU",
                       "frames_per_step": slots},
            "timing": ("CUDA events around every step on every rank (256 MiB L2 flush before it, untimed); " +
                       ("per-rank sum over the K steps, max over ranks: the frames are independent, the job ends when the slowest rank has rendered its K views"
                        if (world > 1 and frames_mode) else "per-step max over ranks, summed over the K steps" if world > 1 else "summed over the K steps")),
            "ms_per_step_lockstep": lockstep_ms,        # frames mode, N > 1: every step waiting for its slowest view (a barrier after each frame)
            "ms_per_frame": ms_per_step, "samples_per_frame": samples / K / slots, "volume_samples_per_frame": n_vol / K / slots,
            "distance_samples_per_frame": n_dist / K / slots, "covered_pixels_per_frame": n_cov / K / slots,
            "mpixels_per_s": slots * FW * FH / (ms_per_step * 1e-3) / 1e6,
            "ess_rebuild_ms": {"median": float(np.median(rebuild_ms)), "mean": float(np.mean(rebuild_ms)), "p95": float(np.percentile(rebuild_ms, 95)),
                               "changes": n_changes, "stages": stage_ms,
                               # the library shards volumes of 512 Mi voxels and more; smaller ones are rebuilt on every replica, with no communication
                               "sharded_z_slabs": world > 1 and int(np.prod(wl["dim"])) >= (512 << 20)},
            "occupied_voxels": occupied, "occupied_percent": 100.0 * occupied / N_vox,
            "modes": modes, "tiles_match_single_gpu_frame": tiles_match, "sharded_rebuild_matches": sharded_rebuild_matches,
            "tiles_8k": tiles_8k, "tf_sweep": tf_sweep, "e2e": e2e, "gpu_launches": int(launches), "wall_s_timed_region": wall,
            "clocks": sampler.summary(), "roofline": roofline, "rooflines_hbm": hbm_rooflines, "tex3d_fetch_per_s": tex_peak,
            "cpu_baseline": cpu_baseline,
        }
        emit(line)
    if world > 1:
        if peer_ptr:
            capi.check(capi.lib().vkv_ipc_close(__import__("ctypes").c_void_p(peer_ptr)))
        torch.cuda.synchronize()
        dist.barrier()
        vol.group_close()
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
    n, dt, views, kept = 0, 0.0, 0, []
    while dt < budget_s and views < 720:
        cu, ru = uniforms_fn(views)
        t0 = time.perf_counter()
        img, c, _, _ = orc.render(V, G, tf, maps, dim_b, cu, ru, tfu, ropt, FW, FH, y_first=y0, y_count=rows)
        dt += time.perf_counter() - t0
        n += c.volume_samples + c.distance_samples
        if len(kept) < 8 and views % 9 == 0:
            kept.append((views, img[y0:y0 + rows].copy()))
        views += 1
    return n / dt / 1e6, dt, rows, n, views, y0, kept


def cpu_baseline_from_device(vol, wl, opt, tfu, ropt, it, FW, FH, skip, uniforms_fn, gpu_frame_fn=None):
    orc = _oracle()
    orc.set_num_threads(host_cores())        # all host cores, whatever OMP_NUM_THREADS the launcher exported
    V = vol.download_voxels()
    G = vol.download_gradient() if tfu.use_gradient else None
    tf = vol.download_transfer_function()
    maps = None
    if skip == 3:
        maps = np.stack([vol.download_distance_map(i) for i in range(8)])
    elif skip != 0:
        maps = vol.download_distance_map(0)
    v, dt, rows, n, views, y0, frames = cpu_render_sample(orc, V, G, tf, maps, vol.map_extent, tfu, ropt, FW, FH, uniforms_fn)
    out = {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
           "sample": f"oracle ray caster (OpenMP, {orc.num_threads()} threads) on {rows} of {FH} rows of {views} orbit views of the same workload: {n} samples in {dt:.2f} s"}
    # the frames the oracle just rendered are compared (untimed) with the GPU's frames of the same views — the north-star bar
    if gpu_frame_fn is not None and frames:
        fr, ps, fa = [], [], []
        for view, ref in frames:
            img = gpu_frame_fn(view)[y0:y0 + rows]
            d = np.abs(img[..., :3].astype(np.int16) - ref[..., :3].astype(np.int16)).max(axis=2)
            fr.append(float((d <= 1).mean()))
            fa.append(float((np.abs(img[..., 3].astype(np.int16) - ref[..., 3].astype(np.int16)) <= 1).mean()))
            mse = float(np.mean((img[..., :3].astype(np.float64) - ref[..., :3].astype(np.float64)) ** 2))
            ps.append(99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse))
        out["frame_parity"] = {"views_compared": len(frames), "rows": [y0, y0 + rows], "frac_within_1": min(fr), "psnr": min(ps), "alpha_frac_within_1": min(fa),
                               "bar": "<= 1/255 per channel on >= 99.9 % of pixels, PSNR >= 50 dB (minimum over the views compared; hardware filter vs the oracle)",
                               "passes": bool(min(fr) >= 0.999 and min(ps) >= 50.0)}
    return out


def run_reference(args):
    """The reference algorithm on the host CPU: synthetic volume -> TF texture -> occupancy -> distance map -> ray-cast band."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = _oracle()
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm must use the whole host at every N, or the
    # ratios the driver computes at N > 1 are against one core.  Set explicitly, then verified below.
    cores = host_cores()
    orc.set_num_threads(cores)
    if orc.num_threads() != cores:
        raise RuntimeError(f"the oracle runs on {orc.num_threads()} threads, the host has {cores} cores: refusing to emit a line")
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
    ap.add_argument("--no-tiles-8k", action="store_true", help="N > 1: skip the extra tile-sharded 8K series")
    ap.add_argument("--gather-frames", action="store_true", help="N > 1, frames mode: store every rank's frame into a ring in rank 0's HBM through peer stores (round 1's variant)")
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
