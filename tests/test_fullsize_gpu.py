"""GPU parity at BASELINE.json's FULL sizes (pytest -m gpu): the CUDA path through the C ABI against the CPU oracle.

What the small-case suite (test_parity_gpu.py) cannot see is checked here at the sizes bench.py measures:

  * configs 1, 2, 3: occupancy map, voxel count, the isotropic and the 8 octant distance maps bit-exact over the whole map;
    frames at the configs' own frame sizes with the HARDWARE filter (the production path), rendered as the second frame of
    a sequence so that the tile-history scheduler is live, compared with the oracle's frame on RGB *and* the stored alpha byte
    a(1-a) at the north-star bar (<= 1/255 on >= 99.9 % of the pixels, PSNR >= 50 dB), every skip mode;
  * config 4: the distance map bit-exact at 1024^3 (occupancy / count / gradient are in test_parity_gpu.py);
  * config 5 (4096x4096x2048, 34 GB): gradient, occupancy and voxel count (> 2^32 voxels shown) against the oracle, the
    1024x1024x512 distance map bit-exact, and a band of rows of the 7680x4320 frame.

shaders/volume_render.frag:117-336, shaders/distance_map.comp:44-109, shaders/distance_map_anisotropic.comp:31-92,
shaders/occupancy_map.comp:45-73, shaders/occupied_voxel_count.comp:25-55 — through their restatement in oracle/.
The workloads (volume generator, TF, camera orbit) are bench.py's own, imported from it.
"""
import math
import os

import numpy as np
import pytest

import bench
import oracle_api as orc
from vkvolume_b200 import capi, scene
from vkvolume_b200.capi import (FILTER_HARDWARE, RenderOptions, VolumeOptions, SKIP_ANISOTROPIC_DISTANCE, SKIP_BLOCK,
                                SKIP_DISTANCE, SKIP_NONE)

pytestmark = pytest.mark.gpu


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def frame_bar_rgba(img, ref):
    """north-star frame tolerance on RGB and, separately, on the stored alpha byte a(1-a)."""
    d = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=2)
    da = np.abs(img[..., 3].astype(int) - ref[..., 3].astype(int))
    return float((d <= 1).mean()), psnr(img[..., :3], ref[..., :3]), float((da <= 1).mean()), psnr(img[..., 3], ref[..., 3])


class _DevPtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _dev_view(ptr, shape):
    import torch
    n = int(np.prod(shape))
    return torch.as_tensor(_DevPtr(ptr, n), device="cuda").view(*shape)


class Workload:
    """One bench.py workload resident on the GPU, plus the oracle's view of it on the host."""

    def __init__(self, ctx, name, download=True):
        import torch
        self.wl = wl = bench.WORKLOADS[name]
        self.W, self.H, self.D = W, H, D = wl["dim"]
        self.FW, self.FH = wl["frame"]
        self.vol = vol = capi.Volume(ctx, W, H, D, block_size=4)
        capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels())
        vol.upload_device(vol.device_voxels())
        self.opt = VolumeOptions(**wl["tf"])
        self.tfu = capi.transfer_function_uniform(self.opt)
        self.use_g = bool(self.tfu.use_gradient)
        vol.compute_gradient_map(self.tfu)
        self.it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
        self.tf = orc.transfer_function_texture(self.opt)
        self.V = self.G = None
        if download:
            self.V = vol.download_voxels()
            self.G = vol.download_gradient() if self.use_g else None
        self.fb = torch.zeros((self.FH, self.FW, 4), dtype=torch.uint8, device="cuda")
        self.counts = torch.zeros(4, dtype=torch.int64, device="cuda")
        self._maps = {}

    def uniforms(self, view):
        cam = scene.look_at_camera(bench.orbit_eye(view, 72, self.wl), aspect=self.FW / self.FH)
        return self.vol.make_uniforms(cam, self.it, self.wl["clip"])

    def oracle_maps(self, skip):
        if skip not in self._maps:
            O = self.oracle_occupancy()
            self._maps[skip] = {SKIP_NONE: None, SKIP_BLOCK: O, SKIP_DISTANCE: None, SKIP_ANISOTROPIC_DISTANCE: None}[skip]
            if skip == SKIP_DISTANCE:
                self._maps[skip] = orc.distance_map(O)
            elif skip == SKIP_ANISOTROPIC_DISTANCE:
                self._maps[skip] = orc.distance_map_anisotropic(O)
        return self._maps[skip]

    def oracle_occupancy(self):
        if "O" not in self._maps:
            self._maps["O"] = orc.occupancy_map(self.V, self.G, self.tf, 4, self.use_g)
        return self._maps["O"]

    def render_sequence(self, views, ropt, want_history=False):
        """Renders `views` back to back into the device frame (as a frame sequence would) and returns the LAST frame, its
        counters and the number of kernels that last frame launched."""
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        n0 = 0
        for k, v in enumerate(views):
            cu, ru = self.uniforms(v)
            if k == len(views) - 1:
                self.counts.zero_()
                torch.cuda.synchronize()
                n0 = capi.kernel_launch_count()
            self.vol.render(cu, ru, self.tfu, ropt, self.FW, self.FH, self.fb.data_ptr(), 0, self.counts.data_ptr(), stream)
        torch.cuda.synchronize()
        return self.fb.cpu().numpy(), [int(x) for x in self.counts.tolist()], capi.kernel_launch_count() - n0

    def close(self):
        self.vol.close()
        self.fb = self.counts = None
        self.V = self.G = None
        self._maps = {}


def _check_maps(w):
    """occupancy, count, isotropic and anisotropic maps of the workload's own TF, bit-exact over the whole map."""
    vol = w.vol
    n = vol.update_transfer_function(w.opt, SKIP_BLOCK, count=True)
    assert n == orc.occupied_voxel_count(w.V, w.G, w.tfu)
    O = w.oracle_occupancy()
    assert np.array_equal(vol.download_distance_map(0), O)
    assert 0 < (O == 0).mean() < 1
    vol.update_transfer_function(w.opt, SKIP_DISTANCE)
    assert np.array_equal(vol.download_distance_map(0), w.oracle_maps(SKIP_DISTANCE))
    vol.update_transfer_function(w.opt, SKIP_ANISOTROPIC_DISTANCE)
    want = w.oracle_maps(SKIP_ANISOTROPIC_DISTANCE)
    for i in range(8):
        assert np.array_equal(vol.download_distance_map(i), want[i]), f"octant map {i}"


def _check_frames(w, skip, views, expect_history):
    vol = w.vol
    vol.update_transfer_function(w.opt, skip)
    ropt = RenderOptions(skipping_type=skip, clip_distance=w.wl["clip"], early_ray_termination=1, filter=FILTER_HARDWARE)
    maps = w.oracle_maps(skip)
    for v in views:
        # the previous orbit view first, then this one twice: the frame compared is the third of a sequence (tile history live)
        img, counts, launches = w.render_sequence([v - 1, v, v], ropt)
        if expect_history:
            assert launches >= 2, f"expected tile_order_kernel + raycast_kernel (+ raycast_long_kernel) on a frame with history, saw {launches} launches"
        cu, ru = w.uniforms(v)
        ref, rc, _, _ = orc.render(w.V, w.G, w.tf, maps, vol.map_extent, cu, ru, w.tfu, ropt, w.FW, w.FH)
        frac, p, frac_a, p_a = frame_bar_rgba(img, ref)
        assert frac >= 0.999 and p >= 50.0, (skip, v, frac, p)
        assert frac_a >= 0.999 and p_a >= 50.0, ("alpha", skip, v, frac_a, p_a)
        assert counts[3] == rc.covered_pixels and rc.covered_pixels > 0.05 * w.FW * w.FH
        assert np.array_equal(img[..., 3] == 255, ref[..., 3] == 255) or frac_a >= 0.9999
        tot, rtot = counts[0] + counts[1], rc.volume_samples + rc.distance_samples
        assert abs(tot - rtot) <= 1e-2 * rtot, (skip, v, tot, rtot)        # hardware filter: TF-threshold flips move a few samples


# ---- config 2: beetle 832x832x494, 1920x1080 --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2(ctx):
    w = Workload(ctx, "c2")
    yield w
    w.close()


def test_config2_maps_and_count_bit_exact(c2):
    _check_maps(c2)


@pytest.mark.parametrize("skip", [SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE, SKIP_BLOCK, SKIP_NONE])
def test_config2_frames_1080p_hardware_filter_with_tile_history(c2, skip):
    views = (0, 17, 40) if skip != SKIP_NONE else (0, 40)        # view 0 is the headline frame of bench.py
    _check_frames(c2, skip, views, expect_history=skip in (SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE))


def test_config2_tile_chunks_beyond_65535_tiles(c2):
    """More than 65 535 tiles in one launch list (gridDim.z limit): 16x8 tiles of a 4096x2304 frame, dealt to two 'ranks';
    the assembled frame equals the full-frame render byte for byte."""
    import torch
    w = c2
    w.vol.update_transfer_function(w.opt, SKIP_DISTANCE)
    FW, FH = 4096, 2304
    cam = scene.look_at_camera(bench.orbit_eye(5, 72, w.wl), aspect=FW / FH)
    cu, ru = w.vol.make_uniforms(cam, w.it, w.wl["clip"])
    ropt = RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=w.wl["clip"])
    stream = torch.cuda.current_stream().cuda_stream
    full = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
    parts = torch.zeros_like(full)
    c_full = torch.zeros(4, dtype=torch.int64, device="cuda")
    c_parts = torch.zeros_like(c_full)
    w.vol.render(cu, ru, w.tfu, ropt, FW, FH, full.data_ptr(), 0, c_full.data_ptr(), stream)
    assert (FW // 16) * (FH // 8) > 65535
    w.vol.render_tiles(cu, ru, w.tfu, ropt, FW, FH, 16, 8, 0, 1, parts.data_ptr(), 0, c_parts.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(full, parts) and torch.equal(c_full, c_parts)
    parts.zero_()
    for r in range(2):
        w.vol.render_tiles(cu, ru, w.tfu, ropt, FW, FH, 16, 8, r, 2, parts.data_ptr(), 0, 0, stream)
    torch.cuda.synchronize()
    assert torch.equal(full, parts)


# ---- config 1: blobs 256^3, 512x512 ------------------------------------------------------------------------------------
def test_config1_maps_gradient_and_frames(ctx):
    w = Workload(ctx, "c1")
    try:
        assert w.use_g and np.array_equal(w.G, orc.gradient_map(w.V, True))
        _check_maps(w)
        for skip in (SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE, SKIP_BLOCK, SKIP_NONE):
            _check_frames(w, skip, (0, 29), expect_history=False)
    finally:
        w.close()


# ---- config 3: 1024x1024x795 anisotropic voxels, camera inside, clip polygon, 8 octant maps -------------------------------
def test_config3_maps_and_frames_camera_inside(ctx):
    w = Workload(ctx, "c3")
    try:
        _check_maps(w)
        cu, ru = w.uniforms(0)
        assert all(0.0 < ru.cam_pos_tex[k] < 1.0 for k in range(3)), "config 3's camera is inside the box"
        # (from inside the volume most tiles hold long rays: the scheduler's device-side decision is "throughput-bound, keep the
        # centre-out order", so no tile_order_kernel launch is asserted here)
        _check_frames(w, SKIP_ANISOTROPIC_DISTANCE, (0, 23, 50), expect_history=False)
        _check_frames(w, SKIP_DISTANCE, (0,), expect_history=False)
        _check_frames(w, SKIP_BLOCK, (23,), expect_history=False)
    finally:
        w.close()


# ---- config 4: 1024^3, the TF sweep's distance maps -----------------------------------------------------------------------
def test_config4_distance_maps_bit_exact_over_the_sweep(ctx):
    """Three settings of config 4's sweep (imin = 0.05 + 0.004 k, gradient window alternating): the GPU's occupancy map is fed
    to the oracle's distance transforms and the GPU's maps must equal them over all 256^3 blocks (K2 itself is compared with
    the oracle at this size in test_parity_gpu.py::test_full_size_properties_config4)."""
    wl = bench.WORKLOADS["c4"]
    W, H, D = wl["dim"]
    vol = capi.Volume(ctx, W, H, D)
    try:
        capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels())
        vol.upload_device(vol.device_voxels())
        vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
        for k in (0, 33, 98):
            opt = bench.sweep_options(wl, k)
            vol.update_transfer_function(opt, SKIP_BLOCK)
            O = vol.download_distance_map(0)
            assert 0 < (O == 0).mean() < 1
            vol.update_transfer_function(opt, SKIP_DISTANCE)
            assert np.array_equal(vol.download_distance_map(0), orc.distance_map(O)), k
            if k == 33:
                vol.update_transfer_function(opt, SKIP_ANISOTROPIC_DISTANCE)
                want = orc.distance_map_anisotropic(O)
                for i in range(8):
                    assert np.array_equal(vol.download_distance_map(i), want[i]), f"octant map {i}"
    finally:
        vol.close()


# ---- config 5: 4096x4096x2048 (34 GB), 7680x4320 -----------------------------------------------------------------------------
def _host_gb_available():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def test_config5_34GB_volume(ctx):
    """Everything config 5 asks of one GPU, against the oracle:
    (1) K1 on z-crops (first slices, an interior crop across an L2-sized stride, the last slices);
    (2) a gradient TF that shows > 2^32 voxels: K2a rows and K2b slab counts against the oracle on three z-slabs, the whole-volume
        count equal to the sum of the 32 slab counts (u64 accumulation past 2^32);
    (3) the workload's own TF: occupancy -> the 1024x1024x512 distance map bit-exact against the oracle's transform of the same
        occupancy map;
    (4) a band of rows of the 7680x4320 frame (hardware filter) against the oracle — needs the 34 GB volume on the host, so it
        runs when the box has the memory and says so when it does not."""
    import torch
    free_b, total_b = torch.cuda.mem_get_info()
    if total_b < 150 * 2 ** 30:
        pytest.skip("config 5 needs a 180 GB device")
    wl = bench.WORKLOADS["c5"]
    W, H, D = wl["dim"]
    FW, FH = wl["frame"]
    vol = capi.Volume(ctx, W, H, D, block_size=4)
    try:
        capi.synth_volume(ctx, wl["kind"], wl["seed"], W, H, D, vol.device_voxels())
        vol.upload_device(vol.device_voxels())
        gopt = VolumeOptions(intensity_min=0.004, intensity_max=1.0, gradient_min=0.0, gradient_max=0.25)
        gtfu = capi.transfer_function_uniform(gopt)
        vol.compute_gradient_map(gtfu)
        Vd = _dev_view(vol.device_voxels(), (D, H, W))
        Gd = _dev_view(vol.device_gradient(), (D, H, W))
        Wb, Hb, Db = vol.map_extent
        assert (Wb, Hb, Db) == (1024, 1024, 512)

        # (1) gradient map on z-crops (1-voxel halo from the neighbouring slices; the volume's first and last slices clamp)
        for z0, z1 in ((0, 10), (1021, 1031), (D - 9, D)):
            lo, hi = max(z0 - 1, 0), min(z1 + 1, D)
            Vc = Vd[lo:hi].cpu().numpy()
            want = orc.gradient_map(Vc, True)[z0 - lo:z0 - lo + (z1 - z0)]
            assert np.array_equal(Gd[z0:z1].cpu().numpy(), want), (z0, z1)

        # (2) > 2^32 visible voxels; slabs against the oracle
        gtf = orc.transfer_function_texture(gopt)
        total = vol.update_transfer_function(gopt, SKIP_BLOCK, count=True)
        assert total > 2 ** 32, total
        O_all = vol.download_distance_map(0)
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        slab_sum = 0
        for s in range(32):
            cnt.zero_()
            vol.compute_occupancy_slab(gtfu, SKIP_BLOCK, 16 * s, 16, count_dev=cnt.data_ptr())
            torch.cuda.synchronize()
            c = int(cnt.item())
            slab_sum += c
            if s in (0, 13, 31):
                Vc, Gc = Vd[64 * s:64 * s + 64].cpu().numpy(), Gd[64 * s:64 * s + 64].cpu().numpy()
                assert c == orc.occupied_voxel_count(Vc, Gc, gtfu), s
                assert np.array_equal(O_all[16 * s:16 * s + 16], orc.occupancy_map(Vc, Gc, gtf, 4, True)), s
                assert np.array_equal(vol.download_distance_map(0)[16 * s:16 * s + 16], O_all[16 * s:16 * s + 16])
        assert slab_sum == total

        # (3) the workload's TF (no gradient): occupancy, then the distance transform of 537 M blocks
        opt = VolumeOptions(**wl["tf"])
        tfu = capi.transfer_function_uniform(opt)
        vol.compute_gradient_map(tfu)        # gradients off at load time: all 255 (quirk A.8.1), as bench.py does
        vol.update_transfer_function(opt, SKIP_BLOCK)
        O = vol.download_distance_map(0)
        occ = float((O == 0).mean())
        assert 0.001 < occ < 0.5, occ
        tf = orc.transfer_function_texture(opt)
        for s in (5, 20):
            Vc = Vd[64 * s:64 * s + 64].cpu().numpy()
            assert np.array_equal(O[16 * s:16 * s + 16], orc.occupancy_map(Vc, None, tf, 4, False)), s
        vol.update_transfer_function(opt, SKIP_DISTANCE)
        Dm = vol.download_distance_map(0)
        want = orc.distance_map(O)
        assert np.array_equal(Dm, want)
        del want

        # (4) 8K frame band
        ropt = RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=wl["clip"], early_ray_termination=1, filter=FILTER_HARDWARE)
        it = scene.image_transform(wl["voxel"], wl["dim"], wl["axis_angle"])
        fb = torch.zeros((FH, FW, 4), dtype=torch.uint8, device="cuda")
        cu = ru = None
        for v in (71, 0, 0):
            cu, ru = vol.make_uniforms(scene.look_at_camera(bench.orbit_eye(v, 72, wl), aspect=FW / FH), it, wl["clip"])
            vol.render(cu, ru, tfu, ropt, FW, FH, fb.data_ptr(), 0, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        need_gb = W * H * D / 2 ** 30 * 1.25
        if _host_gb_available() < need_gb + 16 or os.environ.get("VKV_TEST_NO_34GB_HOST"):
            pytest.skip(f"(1)-(3) passed; the 8K frame band needs {need_gb:.0f} GB of host memory for the oracle's copy of the volume")
        Vh = np.empty((D, H, W), np.uint8)
        for z in range(0, D, 128):
            Vh[z:z + 128] = Vd[z:z + 128].cpu().numpy()
        y0, rows = FH // 2 - 24, 48
        ref, rc, _, _ = orc.render(Vh, None, tf, Dm, vol.map_extent, cu, ru, tfu, ropt, FW, FH, y_first=y0, y_count=rows)
        img = fb[y0:y0 + rows].cpu().numpy()
        frac, p, frac_a, p_a = frame_bar_rgba(img, ref[y0:y0 + rows])
        assert rc.covered_pixels > 0.2 * rows * FW
        assert frac >= 0.999 and p >= 50.0, (frac, p)
        assert frac_a >= 0.999 and p_a >= 50.0, (frac_a, p_a)
    finally:
        vol.close()
