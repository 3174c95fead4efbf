"""GPU parity tests (pytest -m gpu): every CUDA stage, called through the C ABI (libvkv.so via
vkvolume_b200.capi), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north star): occupancy maps, voxel counts and distance maps bit-exact;
gradient within 1e-5 relative (we additionally require byte-exact); frames within 1/255 per
channel on >= 99.9 % of pixels and PSNR >= 50 dB.
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_api as orc
from vkvolume_b200 import capi, scene
from vkvolume_b200.capi import (FILTER_EXACT, FILTER_HARDWARE, RenderOptions, SampleCounts, VolumeOptions,
                                SKIP_ANISOTROPIC_DISTANCE, SKIP_BLOCK, SKIP_DISTANCE, SKIP_NONE,
                                TEST_NUM_TEXTURE_SAMPLES, TEST_RAY_ENTRY, TEST_RAY_EXIT)

pytestmark = pytest.mark.gpu

TF_SETS = [
    dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2),
    dict(intensity_min=0.086, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0),
    dict(intensity_min=0.4, intensity_max=0.8, gradient_min=0.06, gradient_max=0.12),
    dict(intensity_min=0.0, intensity_max=1.0, gradient_min=0.1, gradient_max=0.3),
]


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def frame_bar(img, ref):
    """north-star frame tolerance on RGB (the reference's only export path forces alpha to 255)."""
    d = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=2)
    return float((d <= 1).mean()), psnr(img[..., :3], ref[..., :3])


# ---- transfer function --------------------------------------------------------------------------
@pytest.mark.parametrize("o", TF_SETS)
def test_tf_texture_bit_exact(ctx, o):
    vol = capi.Volume(ctx, 16, 16, 16)
    opt = VolumeOptions(**o)
    vol.update_transfer_function_texture(opt)
    assert np.array_equal(vol.download_transfer_function(), orc.transfer_function_texture(opt))
    u_gpu, u_cpu = capi.transfer_function_uniform(opt), orc.transfer_function_uniform(opt)
    assert bytes(u_gpu) == bytes(u_cpu)
    vol.close()


# ---- loader ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("type_name,dtype", [("uint8_t", np.uint8), ("int8_t", np.int8), ("uint16_t", np.uint16), ("int16_t", np.int16)])
@pytest.mark.parametrize("endian", ["little", "big"])
def test_loader_normalise_bit_exact(ctx, tmp_path, type_name, dtype, endian):
    W, H, D = 21, 10, 7
    rng = np.random.default_rng(11)
    info = np.iinfo(dtype)
    v = rng.integers(info.min, info.max + 1, size=W * H * D).astype(dtype)
    raw = v.astype(v.dtype.newbyteorder(">" if endian == "big" else "<"))
    lo, hi = (400.0, 2538.0) if dtype in (np.uint16, np.int16) else (10.0, 200.0)
    want = orc.normalise(raw.view(np.uint8), v.size, type_name, endian, lo, hi)
    # device path fused with the upload
    vol = capi.Volume(ctx, W, H, D)
    vol.upload_raw(raw.view(np.uint8), type_name, endian, lo, hi)
    assert np.array_equal(vol.download_voxels().ravel(), want)
    vol.close()
    # file path: header + raw file, LoadVolume::load_header / load_data
    fn = tmp_path / "vol.raw"
    raw.tofile(fn)
    (tmp_path / "vol.raw.header").write_text(
        f"{W} {H} {D} # extents\n0.004 0.004 0.008 # voxel size\n{lo} {hi} # normalisation\n{type_name} {endian} # type\n0 1 0 30 # rotation\n")
    h = capi.load_header(str(fn) + ".header")
    assert tuple(h.extent) == (W, H, D) and h.type.decode() == type_name
    ho = orc.parse_header((tmp_path / "vol.raw.header").read_text())
    assert np.allclose(list(h.image_transform), list(ho.image_transform), rtol=1e-6, atol=1e-7)
    assert np.array_equal(capi.load_data(str(fn), h).ravel(), want)


def test_loader_errors(ctx, tmp_path):
    with pytest.raises(capi.VkvError, match="Failed to open header file"):
        capi.load_header(str(tmp_path / "nope.header"))
    fn = tmp_path / "short.raw"
    np.zeros(10, np.uint8).tofile(fn)
    (tmp_path / "short.raw.header").write_text("4 4 4\n1 1 1\n0 255\nuint8_t little\n1 0 0 0\n")
    h = capi.load_header(str(fn) + ".header")
    with pytest.raises(capi.VkvError, match="File size does not match"):
        capi.load_data(str(fn), h)
    (tmp_path / "bad.raw.header").write_text("4 4 4\n1 1 1\n0 255\nfloat little\n1 0 0 0\n")
    hb = capi.load_header(str(tmp_path / "bad.raw.header"))
    with pytest.raises(capi.VkvError, match="unsupported image data type"):
        capi.load_data(str(fn), hb)


# ---- K1 gradient ------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(17, 17, 17), (9, 12, 32), (20, 33, 48), (5, 6, 16), (3, 4, 130), (1, 1, 16), (40, 40, 64)])
def test_gradient_byte_exact(ctx, shape):
    D, H, W = shape
    V = np.random.default_rng(sum(shape)).integers(0, 256, size=shape, dtype=np.uint8)
    if shape == (40, 40, 64):
        V = scene.blobs_volume(shape, seed=4)
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    opt = VolumeOptions(gradient_min=0.0, gradient_max=0.2)
    vol.compute_gradient_map(capi.transfer_function_uniform(opt))
    G = vol.download_gradient()
    assert np.array_equal(G, orc.gradient_map(V))
    # use_gradient false at gradient time -> all 255 (quirk A.8.1)
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.0)))
    assert (vol.download_gradient() == 255).all()
    vol.close()


@pytest.mark.parametrize("kind", ["noise", "smooth_ties", "low_noise"])
def test_gradient_integer_path_byte_exact(ctx, kind):
    """The integer (dp4a) gradient kernel against the oracle on inputs that stress its tie handling: full-range noise,
    a smooth ramp whose gradients sit exactly on rounding boundaries, and +-2 noise (many S == (4n+2)^2 ties)."""
    D, H, W = 37, 45, 160        # W % 16 == 0 -> the vectorised integer kernel
    rng = np.random.default_rng(123)
    if kind == "noise":
        V = rng.integers(0, 256, size=(D, H, W), dtype=np.uint8)
    elif kind == "smooth_ties":
        z, y, x = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
        V = ((2 * x + 4 * y + 6 * z) % 256).astype(np.uint8)
    else:
        V = (128 + rng.integers(-2, 3, size=(D, H, W))).astype(np.uint8)
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    Gref = orc.gradient_map(V)
    assert np.array_equal(vol.download_gradient(), Gref)
    assert np.array_equal(vol.download_gradient_texture(), Gref)        # the copy the ray caster samples
    vol.close()


@pytest.mark.parametrize("shape,kind", [((130, 200, 256), "blobs"), ((70, 90, 1024), "low_noise"), ((300, 64, 48), "noise")])
def test_gradient_persistent_walk_byte_exact(ctx, shape, kind):
    """More 16-voxel chunks than resident threads: every thread of the flat kernel walks several chunks, its tie queue
    carries over between them, and ties are patched into both copies of the map (linear + texture array)."""
    D, H, W = shape
    rng = np.random.default_rng(D + H + W)
    if kind == "blobs":
        V = scene.blobs_volume(shape, seed=11)
    elif kind == "noise":
        V = rng.integers(0, 256, size=shape, dtype=np.uint8)
    else:
        V = (100 + rng.integers(-2, 3, size=shape)).astype(np.uint8)
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    Gref = orc.gradient_map(V)
    assert np.array_equal(vol.download_gradient(), Gref)
    assert np.array_equal(vol.download_gradient_texture(), Gref)
    vol.close()


@pytest.mark.parametrize("knob", ["VKV_GRAD_FLAT", "VKV_GRAD_V1", "VKV_GRAD_FP32"])
def test_gradient_fallback_kernels_byte_exact(ctx, knob, monkeypatch):
    """The column walk is the default since round 2; the flat walk, the row-task kernel and the pure-fp32 kernel remain as fallbacks
    (indices that do not fit, grad_magnitude_modifier != 1) and must keep producing the oracle's bytes."""
    monkeypatch.setenv(knob, "1")
    D, H, W = 41, 77, 208
    V = scene.blobs_volume((D, H, W), seed=5)
    V[::3] = (V[::3].astype(np.int32) + np.random.default_rng(9).integers(-2, 3, size=V[::3].shape)).clip(0, 255).astype(np.uint8)
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    Gref = orc.gradient_map(V)
    assert np.array_equal(vol.download_gradient(), Gref)
    assert np.array_equal(vol.download_gradient_texture(), Gref)
    vol.close()


@pytest.mark.parametrize("shape", [(3, 2, 16), (2, 1, 32), (5, 100, 16), (4, 7, 528), (66, 3, 48)])
def test_gradient_column_walk_edge_shapes(ctx, shape):
    """Column walk corner cases: one or two rows (both parities clamp), a single chunk per row (every lane is first and last),
    rows longer than a warp's 256-byte run (the outer lanes read the neighbouring word from memory), more planes than rows."""
    D, H, W = shape
    V = np.random.default_rng(D * H + W).integers(0, 256, size=shape, dtype=np.uint8)
    V[:, :, ::2] = 128        # many exact ties
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    Gref = orc.gradient_map(V)
    assert np.array_equal(vol.download_gradient(), Gref)
    assert np.array_equal(vol.download_gradient_texture(), Gref)
    vol.close()


# ---- K2a / K2b -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,bs", [((16, 16, 32), 4), ((9, 10, 13), 4), ((24, 20, 48), 4), ((12, 12, 32), 2), ((16, 24, 64), 8),
                                      ((7, 9, 10), 3), ((15, 11, 48), 5), ((10, 10, 10), 1), ((6, 6, 6), 8), ((33, 17, 80), 4),
                                      ((8, 8, 1040), 3)])
@pytest.mark.parametrize("o", TF_SETS)
def test_occupancy_and_count_bit_exact(ctx, shape, bs, o):
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=sum(shape) + bs, n_blobs=5)
    opt = VolumeOptions(**o)
    tfu = capi.transfer_function_uniform(opt)
    G = orc.gradient_map(V, True)        # always feed the oracle's G
    vol = capi.Volume(ctx, W, H, D, block_size=bs)
    vol.upload(V)
    vol.upload_gradient(G)
    vol.update_transfer_function_texture(opt)
    tf = orc.transfer_function_texture(opt)
    want_O = orc.occupancy_map(V, G, tf, bs, bool(tfu.use_gradient))
    assert vol.map_extent == want_O.shape[::-1]
    vol.compute_distance_map(tfu, SKIP_BLOCK)        # occupancy only
    assert np.array_equal(vol.download_distance_map(0), want_O)
    want_n = orc.occupied_voxel_count(V, G, tfu)
    assert vol.compute_occupied_voxel_count(tfu) == want_n
    # fused TF-change path gives the same map and count
    n = vol.update_transfer_function(opt, SKIP_BLOCK, count=True)
    assert n == want_n and np.array_equal(vol.download_distance_map(0), want_O)
    vol.close()


def test_occupancy_arbitrary_texture(ctx):
    """A host-supplied TF texture (not a grey ramp): occupancy follows the texture's alpha plane."""
    shape = (16, 20, 32)
    D, H, W = shape
    V = np.random.default_rng(5).integers(0, 256, size=shape, dtype=np.uint8)
    G = np.random.default_rng(6).integers(0, 256, size=shape, dtype=np.uint8)
    tf = np.zeros((256, 256, 4), np.uint8)
    rng = np.random.default_rng(7)
    tf[..., 3] = np.where(rng.random((256, 256)) < 0.002, rng.integers(1, 256, (256, 256)), 0)
    tf[..., :3] = rng.integers(0, 256, (256, 256, 3))
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    vol.upload_gradient(G)
    vol.set_transfer_function_texture(tf)
    tfu = capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.5))
    vol.compute_distance_map(tfu, SKIP_BLOCK)
    assert np.array_equal(vol.download_distance_map(0), orc.occupancy_map(V, G, tf, 4, True))
    vol.close()


# ---- K3 distance maps ---------------------------------------------------------------------------------------
def _volume_with_occupancy(ctx, O):
    """Builds a volume whose occupancy map equals O (1 voxel = 1 block) so K3 can be driven in isolation."""
    Db, Hb, Wb = O.shape
    vol = capi.Volume(ctx, Wb, Hb, Db, block_size=1)
    vol.upload(np.where(O == 0, 255, 0).astype(np.uint8))
    opt = VolumeOptions(intensity_min=0.5, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
    vol.update_transfer_function_texture(opt)
    return vol, capi.transfer_function_uniform(opt)


@pytest.mark.parametrize("shape,p", [((8, 9, 10), 0.05), ((16, 12, 20), 0.01), ((5, 24, 7), 0.1), ((1, 1, 30), 0.1), ((3, 1, 1), 0.5),
                                     ((20, 40, 70), 0.002), ((33, 65, 129), 0.0005), ((4, 4, 300), 0.003), ((300, 3, 5), 0.003),
                                     ((2, 1100, 3), 0.002), ((6, 6, 6), 0.0),
                                     # long lines through the fast paths (Wb % 4 == 0): z walk with saturation beyond 255 and its value
                                     # table wrapping (lines > 256 cells), chamfer y sweep, 64-word x rows, sparse sources
                                     ((300, 8, 16), 0.002), ((40, 300, 32), 0.001), ((70, 36, 128), 0.0005), ((520, 4, 8), 0.001),
                                     ((4, 8, 2048), 0.0005), ((2, 4, 2052), 0.0008), ((130, 50, 64), 0.00005), ((33, 17, 1024), 0.0002),
                                     # rows wider than one 256-cell strip of the y sweep: several edge exchanges (every 16 rows), a ragged
                                     # last strip, a last strip that owns a handful of cells; 8 / 16 cells per lane in the x pass
                                     ((6, 70, 544), 0.001), ((3, 40, 260), 0.004), ((2, 100, 1000), 0.0002), ((5, 33, 456), 0.002),
                                     ((700, 6, 36), 0.001)])
def test_distance_maps_bit_exact(ctx, shape, p):
    rng = np.random.default_rng(abs(hash(shape)) % 1000)
    O = np.where(rng.random(shape) < p, 0, 255).astype(np.uint8)
    vol, tfu = _volume_with_occupancy(ctx, O)
    vol.compute_distance_map(tfu, SKIP_BLOCK)
    assert np.array_equal(vol.download_distance_map(0), O)
    vol.compute_distance_map(tfu, SKIP_DISTANCE)
    assert np.array_equal(vol.download_distance_map(0), orc.distance_map(O))
    vol.compute_distance_map(tfu, SKIP_ANISOTROPIC_DISTANCE)
    assert vol.map_extent == O.shape[::-1]
    want = orc.distance_map_anisotropic(O)
    for i in range(8):
        assert np.array_equal(vol.download_distance_map(i), want[i]), f"octant map {i}"
    vol.close()


@pytest.mark.parametrize("shape,cells", [((600, 8, 16), [(0, 0, 0)]), ((600, 8, 16), [(599, 7, 15), (300, 0, 8)]),
                                         ((8, 600, 16), [(7, 0, 3)]), ((4, 8, 1024), [(0, 0, 1023)]), ((290, 12, 36), [(289, 0, 0), (0, 11, 35)]),
                                         ((64, 64, 64), []), ((512, 4, 32), [(255, 1, 16)])])
def test_distance_maps_saturation_and_far_sources(ctx, shape, cells):
    """Single far sources: every line saturates at 255 somewhere (the cap of distance_map.comp:77,96) and the segment heads
    of the z walk / the chunk boundaries of the y sweep sit in empty space."""
    O = np.full(shape, 255, np.uint8)
    for c in cells:
        O[c] = 0
    vol, tfu = _volume_with_occupancy(ctx, O)
    vol.compute_distance_map(tfu, SKIP_DISTANCE)
    assert np.array_equal(vol.download_distance_map(0), orc.distance_map(O))
    vol.compute_distance_map(tfu, SKIP_ANISOTROPIC_DISTANCE)
    want = orc.distance_map_anisotropic(O)
    for i in range(8):
        assert np.array_equal(vol.download_distance_map(i), want[i]), f"octant map {i}"
    vol.close()


# ---- K4 ray caster ----------------------------------------------------------------------------------------------
def _render_case(ctx, shape, opt, eye, clip, skip, width, height, voxel_size=(0.004, 0.004, 0.004), axis_angle=(1, 0, 0, 0),
                 ert=1, test=0, bs=4, filt=FILTER_EXACT, seed=1):
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=seed)
    tfu = capi.transfer_function_uniform(opt)
    G = orc.gradient_map(V, bool(tfu.use_gradient))
    tf = orc.transfer_function_texture(opt)
    O = orc.occupancy_map(V, G, tf, bs, bool(tfu.use_gradient))
    maps = None
    if skip == SKIP_BLOCK:
        maps = O
    elif skip == SKIP_DISTANCE:
        maps = orc.distance_map(O)
    elif skip == SKIP_ANISOTROPIC_DISTANCE:
        maps = orc.distance_map_anisotropic(O)
    vol = capi.Volume(ctx, W, H, D, block_size=bs)
    vol.upload(V)
    vol.upload_gradient(G)
    vol.update_transfer_function_texture(opt)
    vol.compute_distance_map(tfu, skip)
    it = scene.image_transform(voxel_size, (W, H, D), axis_angle)
    cam = scene.look_at_camera(eye, aspect=width / height)
    cu, ru = vol.make_uniforms(cam, it, clip)
    ropt = RenderOptions(skipping_type=skip, clip_distance=clip, early_ray_termination=ert, test=test, filter=filt)
    img, counts = vol.render_to_host(cu, ru, tfu, ropt, width, height)
    ref, rcounts, rf, _ = orc.render(V, G, tf, maps, vol.map_extent, cu, ru, tfu, ropt, width, height, want_float=True)
    vol.close()
    return img, counts, ref, rcounts, rf


@pytest.mark.parametrize("skip", [SKIP_NONE, SKIP_BLOCK, SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE])
@pytest.mark.parametrize("o", TF_SETS[:3])
def test_render_exact_filter_matches_oracle(ctx, skip, o):
    img, counts, ref, rcounts, _ = _render_case(ctx, (48, 64, 80), VolumeOptions(**o), (34, 22, 50), 5.0, skip, 160, 128)
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)
    assert counts.covered_pixels == rcounts.covered_pixels
    # sample counters agree to within the handful of pixels whose step count flips by rounding
    tot, rtot = counts.volume_samples + counts.distance_samples, rcounts.volume_samples + rcounts.distance_samples
    assert abs(tot - rtot) <= 1e-3 * rtot + 8
    assert np.array_equal(img[..., 3] == 255, ref[..., 3] == 255)        # coverage mask via the clear alpha


@pytest.mark.parametrize("skip", [SKIP_NONE, SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE])
def test_render_hardware_filter_within_frame_bar(ctx, skip):
    o = TF_SETS[0]
    img, counts, ref, rcounts, _ = _render_case(ctx, (48, 64, 80), VolumeOptions(**o), (34, 22, 50), 5.0, skip, 160, 128,
                                                filt=FILTER_HARDWARE)
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)


def test_render_camera_inside_with_clip_plane_anisotropic_voxels(ctx):
    """Config-3 shaped case: anisotropic voxels, rotated volume, camera inside the box, clip polygon on screen."""
    opt = VolumeOptions(**TF_SETS[0])
    img, counts, ref, rcounts, rf = _render_case(ctx, (40, 64, 64), opt, (3, 2, 6), 4.0, SKIP_ANISOTROPIC_DISTANCE, 128, 96,
                                                 voxel_size=(0.003, 0.003, 0.007), axis_angle=(1, 0, 0, 90))
    assert rcounts.covered_pixels == 128 * 96        # inside the box: every pixel enters through the clip polygon
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)
    assert counts.covered_pixels == rcounts.covered_pixels


# ---- depth output, depth attachment, compositing over an existing target -------------------------------
def _scene_for_compositing(ctx, shape, opt, seed, skip, bs=4):
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=seed)
    tfu = capi.transfer_function_uniform(opt)
    G = orc.gradient_map(V, bool(tfu.use_gradient))
    tf = orc.transfer_function_texture(opt)
    O = orc.occupancy_map(V, G, tf, bs, bool(tfu.use_gradient))
    maps = {SKIP_NONE: None, SKIP_BLOCK: O, SKIP_DISTANCE: orc.distance_map(O), SKIP_ANISOTROPIC_DISTANCE: orc.distance_map_anisotropic(O)}[skip]
    vol = capi.Volume(ctx, W, H, D, block_size=bs)
    vol.upload(V)
    vol.upload_gradient(G)
    vol.update_transfer_function_texture(opt)
    vol.compute_distance_map(tfu, skip)
    return dict(V=V, G=G, tf=tf, tfu=tfu, maps=maps, vol=vol)


def _depth_close(d, ref):
    return np.allclose(d, ref, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("skip", [SKIP_NONE, SKIP_DISTANCE])
def test_render_depth_output_matches_oracle(ctx, skip):
    """gl_FragDepth at the first contributing sample (volume_render.frag:314-321), 0 where nothing was hit."""
    opt = VolumeOptions(**TF_SETS[0])
    sc = _scene_for_compositing(ctx, (48, 64, 80), opt, 1, skip)
    vol, width, height = sc["vol"], 160, 128
    it = scene.image_transform((0.004,) * 3, (80, 64, 48))
    cu, ru = vol.make_uniforms(scene.look_at_camera((34, 22, 50), aspect=width / height), it, 5.0)
    ropt = RenderOptions(skipping_type=skip, clip_distance=5.0, filter=FILTER_EXACT)
    img, depth, counts = vol.render_over_host(cu, ru, sc["tfu"], ropt, width, height, depth=np.full((height, width), 7.0, np.float32))
    ref, rcounts, _, rdepth = orc.render(sc["V"], sc["G"], sc["tf"], sc["maps"], vol.map_extent, cu, ru, sc["tfu"], ropt, width, height, want_depth=True)
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)
    assert (rdepth > 0).sum() > 1000 and np.array_equal(depth > 0, rdepth > 0)
    assert _depth_close(depth, rdepth)
    vol.close()


def _depth_wall(sc, cu, ru, tfu, width, height, clip):
    """A synthetic depth attachment: stripes of nothing (0 = far) | geometry inside the volume | geometry in front of it."""
    ropt = RenderOptions(skipping_type=SKIP_NONE, clip_distance=clip, test=TEST_RAY_ENTRY)
    _, _, rf, _ = orc.render(sc["V"], sc["G"], sc["tf"], None, sc["vol"].map_extent, cu, ru, tfu, ropt, width, height, want_float=True)
    cov = rf[..., 3] >= 0
    m = lambda a: np.array(list(a), np.float32).astype(np.float64).reshape(4, 4).T
    pvm = (m(cu.proj) @ m(cu.view)) @ m(cu.model)
    pm = np.concatenate([rf[..., :3].astype(np.float64) - 0.5, np.ones((height, width, 1))], axis=2)
    pos = pm @ pvm.T
    front = np.where(cov, pos[..., 2] / pos[..., 3], 0.0)
    yy, xx = np.mgrid[0:height, 0:width]
    sel = ((xx // 8) + (yy // 8)) % 4
    depth = (front * np.choose(sel, [0.0, 1 / 1.004, 1 / 1.02, 1.1])).astype(np.float32)
    depth[~cov] = np.where(sel[~cov] == 3, np.float32(0.25), np.float32(0.0))
    return depth, cov


@pytest.mark.parametrize("skip", [SKIP_NONE, SKIP_BLOCK, SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE])
@pytest.mark.parametrize("filt", [FILTER_EXACT, FILTER_HARDWARE])
def test_render_depth_attachment_over_background(ctx, skip, filt):
    """DEPTH_ATTACHMENT variant (volume_render.frag:122-136,151-165) over a coloured background with a depth image: fragments
    behind the scene are discarded, rays stop at the scene's depth, the rest blends over the background (sRGB target)."""
    opt = VolumeOptions(**TF_SETS[0])
    sc = _scene_for_compositing(ctx, (48, 64, 80), opt, 1, skip)
    vol, width, height, clip = sc["vol"], 160, 128, 5.0
    it = scene.image_transform((0.004,) * 3, (80, 64, 48))
    cu, ru = vol.make_uniforms(scene.look_at_camera((34, 22, 50), aspect=width / height), it, clip)
    depth0, cov = _depth_wall(sc, cu, ru, sc["tfu"], width, height, clip)
    rng = np.random.default_rng(3)
    bg = rng.integers(0, 256, size=(height, width, 4), dtype=np.uint8)
    ropt = RenderOptions(skipping_type=skip, clip_distance=clip, filter=filt, depth_attachment=1, load_framebuffer=1)
    img, depth, counts = vol.render_over_host(cu, ru, sc["tfu"], ropt, width, height, rgba=bg.copy(), depth=depth0.copy())
    ref, rcounts, rf, rdepth = orc.render(sc["V"], sc["G"], sc["tf"], sc["maps"], vol.map_extent, cu, ru, sc["tfu"], ropt, width, height,
                                          want_float=True, rgba_init=bg, depth_init=depth0)
    discarded = rf[..., 3] == -2.0
    assert discarded.sum() > 500 and (cov & ~discarded).sum() > 2000
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)
    keep = ~cov | discarded        # pixels the volume does not write keep the background and its depth, byte for byte
    assert np.array_equal(img[keep], bg[keep]) and np.array_equal(depth[keep], depth0[keep])
    assert _depth_close(depth, rdepth)
    assert counts.covered_pixels == rcounts.covered_pixels
    if filt == FILTER_EXACT:
        tot, rtot = counts.volume_samples + counts.distance_samples, rcounts.volume_samples + rcounts.distance_samples
        assert abs(tot - rtot) <= 1e-3 * rtot + 8
    vol.close()


def test_render_depth_attachment_argument_checks(ctx):
    opt = VolumeOptions(**TF_SETS[1])
    sc = _scene_for_compositing(ctx, (16, 16, 16), opt, 2, SKIP_NONE)
    vol = sc["vol"]
    cu, ru = vol.make_uniforms(scene.look_at_camera((10, 8, 14), aspect=1.0), scene.image_transform((0.004,) * 3, (16, 16, 16)), 1.0)
    with pytest.raises(capi.VkvError):        # no depth buffer
        vol.render_over_host(cu, ru, sc["tfu"], RenderOptions(skipping_type=SKIP_NONE, depth_attachment=1, load_framebuffer=1), 32, 32)
    with pytest.raises(capi.VkvError):        # depth attachment without loading the target
        vol.render_over_host(cu, ru, sc["tfu"], RenderOptions(skipping_type=SKIP_NONE, depth_attachment=1), 32, 32, depth=np.zeros((32, 32), np.float32))
    vol.close()


@pytest.mark.parametrize("skip", [SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE])
def test_render_two_volumes_composited_in_scene_order(ctx, skip):
    """VolumeRenderSubpass::draw loops over the scene's volumes (volume_render_subpass.cpp:219-293): the second volume blends
    and depth-tests over what the first one left in the target."""
    width, height, clip = 160, 128, 5.0
    a = _scene_for_compositing(ctx, (48, 64, 80), VolumeOptions(**TF_SETS[0]), 1, skip)
    b = _scene_for_compositing(ctx, (40, 40, 56), VolumeOptions(**TF_SETS[1]), 5, skip)
    cam = scene.look_at_camera((34, 22, 50), aspect=width / height)
    cu_a, ru_a = a["vol"].make_uniforms(cam, scene.image_transform((0.004,) * 3, (80, 64, 48)), clip)
    cu_b, ru_b = b["vol"].make_uniforms(cam, scene.image_transform((0.006, 0.005, 0.004), (56, 40, 40), (0, 1, 0, 25)), clip)
    first = RenderOptions(skipping_type=skip, clip_distance=clip, filter=FILTER_EXACT)
    second = RenderOptions(skipping_type=skip, clip_distance=clip, filter=FILTER_EXACT, load_framebuffer=1)
    img, depth, _ = a["vol"].render_over_host(cu_a, ru_a, a["tfu"], first, width, height, depth=np.zeros((height, width), np.float32))
    only_a = img.copy()
    img, depth, _ = b["vol"].render_over_host(cu_b, ru_b, b["tfu"], second, width, height, rgba=img, depth=depth)
    ref, _, _, rdepth = orc.render(a["V"], a["G"], a["tf"], a["maps"], a["vol"].map_extent, cu_a, ru_a, a["tfu"], first, width, height, want_depth=True)
    ref, _, rf, rdepth = orc.render(b["V"], b["G"], b["tf"], b["maps"], b["vol"].map_extent, cu_b, ru_b, b["tfu"], second, width, height,
                                    want_float=True, rgba_init=ref, depth_init=rdepth)
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)
    assert _depth_close(depth, rdepth)
    assert (img != only_a).any(axis=2).sum() > 500        # the second volume really drew over the first
    a["vol"].close()
    b["vol"].close()


@pytest.mark.parametrize("test_mode", [TEST_RAY_ENTRY, TEST_RAY_EXIT, TEST_NUM_TEXTURE_SAMPLES])
def test_render_debug_views(ctx, test_mode):
    img, counts, ref, rcounts, _ = _render_case(ctx, (48, 64, 80), VolumeOptions(**TF_SETS[0]), (34, 22, 50), 5.0, SKIP_DISTANCE,
                                                128, 128, ert=0, test=test_mode)
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)


def test_render_sampling_factor_and_alpha_factor(ctx):
    opt = VolumeOptions(sampling_factor=2.0, voxel_alpha_factor=1.5, **TF_SETS[1])
    img, counts, ref, rcounts, _ = _render_case(ctx, (48, 64, 80), opt, (34, 22, 50), 5.0, SKIP_DISTANCE, 128, 128)
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)


# ---- on-the-fly gradient variant (`--gradient_test`, PRECOMPUTED_GRADIENT undefined) ----------------
@pytest.mark.parametrize("shape,bs", [((16, 16, 32), 4), ((9, 10, 13), 4), ((7, 9, 10), 3), ((33, 17, 300), 4), ((12, 12, 32), 2)])
@pytest.mark.parametrize("o", TF_SETS)
def test_on_the_fly_gradient_occupancy_and_count_bit_exact(ctx, shape, bs, o):
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=sum(shape) + bs, n_blobs=5)
    opt = VolumeOptions(use_precomputed_gradient=0, **o)
    tfu = capi.transfer_function_uniform(opt)
    vol = capi.Volume(ctx, W, H, D, block_size=bs, use_precomputed_gradient=False)        # no G, no gradient array: V only
    vol.upload(V)
    tf = orc.transfer_function_texture(opt)
    want_O = orc.occupancy_map(V, None, tf, bs, bool(tfu.use_gradient), precomputed=False)
    want_n = orc.occupied_voxel_count(V, None, tfu, precomputed=False)
    n = vol.update_transfer_function(opt, SKIP_BLOCK, count=True)
    assert np.array_equal(vol.download_distance_map(0), want_O)
    assert n == want_n
    assert vol.compute_occupied_voxel_count(tfu) == want_n
    with pytest.raises(capi.VkvError):
        vol.compute_gradient_map(tfu)        # there is no map to compute in this mode
    vol.close()


@pytest.mark.parametrize("skip", [SKIP_NONE, SKIP_DISTANCE])
@pytest.mark.parametrize("filt", [FILTER_EXACT, FILTER_HARDWARE])
def test_on_the_fly_gradient_render_matches_oracle(ctx, skip, filt):
    shape, width, height = (48, 64, 80), 160, 128
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=1)
    opt = VolumeOptions(use_precomputed_gradient=0, **TF_SETS[0])
    tfu = capi.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    O = orc.occupancy_map(V, None, tf, 4, True, precomputed=False)
    maps = orc.distance_map(O) if skip == SKIP_DISTANCE else None
    vol = capi.Volume(ctx, W, H, D, block_size=4, use_precomputed_gradient=False)
    vol.upload(V)
    vol.update_transfer_function(opt, skip)
    if skip == SKIP_DISTANCE:
        assert np.array_equal(vol.download_distance_map(0), maps)
    it = scene.image_transform((0.004,) * 3, (W, H, D))
    cu, ru = vol.make_uniforms(scene.look_at_camera((34, 22, 50), aspect=width / height), it, 5.0)
    ropt = RenderOptions(skipping_type=skip, clip_distance=5.0, filter=filt)
    img, counts = vol.render_to_host(cu, ru, tfu, ropt, width, height)
    ref, rcounts, _, _ = orc.render(V, None, tf, maps, vol.map_extent, cu, ru, tfu, ropt, width, height, precomputed=False)
    vol.close()
    frac, p = frame_bar(img, ref)
    assert frac >= 0.999 and p >= 50.0, (frac, p)
    assert counts.covered_pixels == rcounts.covered_pixels
    tot, rtot = counts.volume_samples + counts.distance_samples, rcounts.volume_samples + rcounts.distance_samples
    assert abs(tot - rtot) <= (1e-3 if filt == FILTER_EXACT else 2e-2) * rtot + 8


def test_render_tiles_equal_full_frame(ctx):
    """Image-tile sharding: rendering tile subsets (as ranks would) reassembles the full frame byte for byte."""
    import torch
    shape = (48, 64, 80)
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=1)
    opt = VolumeOptions(**TF_SETS[0])
    tfu = capi.transfer_function_uniform(opt)
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(V)
    vol.compute_gradient_map(tfu)
    vol.update_transfer_function(opt, SKIP_DISTANCE)
    it = scene.image_transform((0.004,) * 3, (W, H, D))
    width, height = 200, 120
    cu, ru = vol.make_uniforms(scene.look_at_camera((34, 22, 50), aspect=width / height), it, 5.0)
    ropt = RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0)
    full = torch.zeros((height, width, 4), dtype=torch.uint8, device="cuda")
    vol.render(cu, ru, tfu, ropt, width, height, full.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    parts = torch.zeros_like(full)
    for rank in range(3):
        vol.render_tiles(cu, ru, tfu, ropt, width, height, 32, 16, rank, 3, parts.data_ptr(),
                         stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(full, parts)
    vol.close()


# ---- end to end at BASELINE size: size-independent properties ---------------------------------------------------------------
def test_full_size_properties_config2(ctx):
    """832x832x494 synthetic beetle (config 2 size): properties that need no CPU pass over the full volume."""
    import torch
    W, H, D = 832, 832, 494
    vol = capi.Volume(ctx, W, H, D)
    capi.synth_volume(ctx, 1, 0x5EED0002, W, H, D, vol.device_voxels())
    vol.upload_device(vol.device_voxels())
    opt = VolumeOptions(intensity_min=0.086, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(tfu)
    n = vol.update_transfer_function(opt, SKIP_ANISOTROPIC_DISTANCE, count=True)
    maps = np.stack([vol.download_distance_map(i) for i in range(8)])
    vol.update_transfer_function(opt, SKIP_DISTANCE)
    iso = vol.download_distance_map(0)
    vol.update_transfer_function(opt, SKIP_BLOCK)
    O = vol.download_distance_map(0)
    # (1) isotropic map == min over the 8 octant maps; zero exactly on occupied blocks
    assert np.array_equal(iso, maps.min(axis=0))
    assert np.array_equal(iso == 0, O == 0)
    # (2) 1-Lipschitz in every axis (a Chebyshev distance field), saturating at 255
    for ax in range(3):
        d = np.abs(np.diff(iso.astype(np.int16), axis=ax))
        assert d.max() <= 1
    # (3) the count is a checksum of the occupancy: every counted voxel lies in an occupied block for this TF
    #     (no gradient, so analytic alpha > 0 <=> V >= 22 while texture alpha > 0 <=> V >= 23)
    Vh = vol.download_voxels()
    assert n == int((Vh.astype(np.float32) / np.float32(255) > np.float32(0.086)).sum())
    # effective block size is 4 in x,y; z: ceil(494/124) = 4 with a 2-voxel last block
    Ob = np.zeros(iso.shape, bool)
    vis = Vh >= 23
    for z in range(iso.shape[0]):
        Ob[z] = vis[z * 4:(z + 1) * 4].any(axis=0).reshape(H // 4, 4, W // 4, 4).any(axis=(1, 3))
    assert np.array_equal(O == 0, Ob)
    # (4) a crop of the full-size result equals the oracle run on the crop's own closed form neighbourhood
    sub = O[40:56, 60:84, 70:100]
    if (sub == 0).any():
        crop = iso[40:56, 60:84, 70:100]
        loc = orc.distance_map(sub)
        assert (crop <= loc).all()        # more occupied blocks outside the crop can only shorten distances
    vol.close()


def test_render_to_host_async_pipeline_equals_blocking_calls(ctx):
    """vkv_render_to_host_async (copy-out of frame k overlapping the casting of frame k+1, ring of device frames) returns the same
    frames and counters as one blocking vkv_render_to_host per view."""
    import torch
    opt = VolumeOptions(**TF_SETS[0])
    sc = _scene_for_compositing(ctx, (48, 64, 80), opt, 1, SKIP_DISTANCE)
    vol, width, height, n = sc["vol"], 192, 128, 7
    it = scene.image_transform((0.004,) * 3, (80, 64, 48))
    views = [vol.make_uniforms(scene.look_at_camera((34 * math.cos(0.3 * k), 22, 50 * math.sin(0.3 * k) + 20), aspect=width / height), it, 5.0) for k in range(n)]
    ropt = RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0)
    frames = torch.empty((n, height, width, 4), dtype=torch.uint8).pin_memory()
    counts = torch.zeros((n, 4), dtype=torch.int64).pin_memory()
    for k, (cu, ru) in enumerate(views):
        vol.render_to_host_async(cu, ru, sc["tfu"], ropt, width, height, frames[k].data_ptr(), counts[k].data_ptr())
    vol.render_to_host_wait()
    for k, (cu, ru) in enumerate(views):
        img, c = vol.render_to_host(cu, ru, sc["tfu"], ropt, width, height)
        assert np.array_equal(frames[k].numpy(), img), k
        assert counts[k].tolist() == [c.volume_samples, c.distance_samples, c.empty_samples, c.covered_pixels]
    pageable = np.empty((height, width, 4), np.uint8)
    with pytest.raises(capi.VkvError):        # pageable destinations are refused, not silently made synchronous
        vol.render_to_host_async(views[0][0], views[0][1], sc["tfu"], ropt, width, height, pageable.ctypes.data)
    vol.close()


@pytest.mark.parametrize("dims", [(1024, 512, 256), (1024, 1024, 160)])
def test_tma_occupancy_ring_on_wide_volumes_with_gradient(ctx, dims):
    """Regression (config 4's shape class): 1024-voxel-wide volumes of 128 MB and more with a gradient transfer function.  With a
    ring of five stages a consumer group could run ahead of an out-of-order bulk copy and consume a slot that was not its own
    (unspecified launch failure).  The TMA-staged kernel must agree with the register-staged one, map and count."""
    import os
    W, H, D = dims
    vol = capi.Volume(ctx, W, H, D)
    capi.synth_volume(ctx, 3, 0x5EED0005, W, H, D, vol.device_voxels())
    vol.upload_device(vol.device_voxels())
    opt = VolumeOptions(intensity_min=0.15, intensity_max=1.0, gradient_min=0.02, gradient_max=0.2)
    tfu = capi.transfer_function_uniform(opt)
    vol.compute_gradient_map(tfu)
    results = []
    for no_tma in ("0", "1", "0"):
        os.environ["VKV_OCC_NO_TMA"] = no_tma if no_tma == "1" else ""
        if no_tma != "1":
            os.environ.pop("VKV_OCC_NO_TMA")
        n = vol.update_transfer_function(opt, SKIP_BLOCK, count=True)
        results.append((n, vol.download_distance_map(0)))
    os.environ.pop("VKV_OCC_NO_TMA", None)
    assert results[0][0] == results[1][0] == results[2][0] and results[0][0] > 0
    assert np.array_equal(results[0][1], results[1][1]) and np.array_equal(results[0][1], results[2][1])
    assert 0 < (results[0][1] == 0).mean() < 1
    vol.close()


def test_full_size_properties_config4(ctx):
    """1024^3 with a gradient transfer function (config 4 size): occupancy map and voxel count bit-exact against the oracle run over
    the whole downloaded volume, and the distance-map invariants, at three settings of the TF sweep."""
    W = H = D = 1024
    vol = capi.Volume(ctx, W, H, D)
    capi.synth_volume(ctx, 3, 0x5EED0004, W, H, D, vol.device_voxels())
    vol.upload_device(vol.device_voxels())
    vol.compute_gradient_map(capi.transfer_function_uniform(VolumeOptions(gradient_min=0.0, gradient_max=0.2)))
    Vh, Gh = vol.download_voxels(), vol.download_gradient()
    assert np.array_equal(Gh, orc.gradient_map(Vh, True))        # K1 byte-exact over the whole 1024^3 volume
    assert np.array_equal(vol.download_gradient_texture(), Gh)   # and the copy the ray caster samples (written by the same kernel)
    counts = []
    for k in (0, 7, 24):        # three settings of the sweep (imin = 0.05 + 0.004 k, gradient window on)
        opt = VolumeOptions(intensity_min=0.05 + 0.004 * k, intensity_max=1.0, gradient_min=0.05, gradient_max=0.25)
        tfu = capi.transfer_function_uniform(opt)
        n = vol.update_transfer_function(opt, SKIP_DISTANCE, count=True)
        counts.append(n)
        iso = vol.download_distance_map(0)
        vol.update_transfer_function(opt, SKIP_BLOCK)
        O = vol.download_distance_map(0)
        if k == 7:        # the oracle over the full volume (OpenMP): occupancy map and count bit-exact at BASELINE size
            tf = orc.transfer_function_texture(opt)
            assert np.array_equal(O, orc.occupancy_map(Vh, Gh, tf, 4, True))
            assert n == orc.occupied_voxel_count(Vh, Gh, tfu)
        assert np.array_equal(iso == 0, O == 0)
        for ax in range(3):
            assert np.abs(np.diff(iso.astype(np.int16), axis=ax)).max() <= 1
    assert counts[0] > counts[1] > counts[2] > 0        # a rising intensity threshold shows fewer voxels
    vol.close()


def test_distance_from_occupancy_needs_a_fresh_occupancy_map(ctx):
    """K3 consumes the occupancy map in place (quirk A.8.3): running it twice in a row must fail loudly instead of transforming a
    distance map and calling the result valid."""
    O = np.where(np.random.default_rng(3).random((12, 16, 20)) < 0.02, 0, 255).astype(np.uint8)
    vol, tfu = _volume_with_occupancy(ctx, O)
    vol.compute_occupancy_slab(tfu, SKIP_DISTANCE, 0, vol.map_extent[2])
    vol.compute_distance_from_occupancy(SKIP_DISTANCE)
    want = orc.distance_map(O)
    assert np.array_equal(vol.download_distance_map(0), want)
    with pytest.raises(capi.VkvError, match="does not hold an occupancy map"):
        vol.compute_distance_from_occupancy(SKIP_DISTANCE)
    assert np.array_equal(vol.download_distance_map(0), want)        # untouched by the refused call
    vol.compute_occupancy_slab(tfu, SKIP_BLOCK, 0, vol.map_extent[2])
    vol.compute_distance_from_occupancy(SKIP_BLOCK)                   # block mode leaves the occupancy map in place: repeatable
    vol.compute_distance_from_occupancy(SKIP_BLOCK)
    assert np.array_equal(vol.download_distance_map(0), O)
    vol.close()


def test_two_streams_alternating_on_one_volume_are_ordered(ctx):
    """A volume's colour table, tile history and counters are shared state: renders issued alternately on two streams with
    different sampling / alpha factors (each forces a rebuild of the colour table) must give the frames the same calls give on one
    stream — the library orders a call on a new stream after the volume's previous stream (include/vkv.h, stream model)."""
    import torch
    sc = _scene_for_compositing(ctx, (48, 64, 80), VolumeOptions(**TF_SETS[0]), 1, SKIP_DISTANCE)
    vol, width, height = sc["vol"], 512, 384
    it = scene.image_transform((0.004,) * 3, (80, 64, 48))
    cu, ru = vol.make_uniforms(scene.look_at_camera((34, 22, 50), aspect=width / height), it, 5.0)
    ropt = RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0)
    tfus = [capi.transfer_function_uniform(VolumeOptions(sampling_factor=sf, voxel_alpha_factor=af, **TF_SETS[0])) for sf, af in ((1.0, 1.0), (2.0, 0.6))]
    want = []
    for u in tfus:
        fb = torch.zeros((height, width, 4), dtype=torch.uint8, device="cuda")
        vol.render(cu, ru, u, ropt, width, height, fb.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want.append(fb)
    assert not torch.equal(want[0], want[1])
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    frames = [torch.zeros((height, width, 4), dtype=torch.uint8, device="cuda") for _ in range(12)]
    torch.cuda.synchronize()
    for k, fb in enumerate(frames):
        vol.render(cu, ru, tfus[k & 1], ropt, width, height, fb.data_ptr(), stream=streams[k & 1].cuda_stream)
    torch.cuda.synchronize()
    for k, fb in enumerate(frames):
        assert torch.equal(fb, want[k & 1]), k
    vol.close()


@pytest.mark.parametrize("skip", [SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE])
def test_long_ray_pass_is_bit_identical(ctx, skip):
    """Rays still marching after VKV_RC_LONG_T loop trips are suspended and finished by raycast_long_kernel (one ray per warp, 64
    lattice steps evaluated at once, the shader's state machine replayed over them).  The replay visits the same steps in the same
    order with the same arithmetic: frames, depth and all four counters are identical whatever the hand-over point — never (0),
    the default (64), or almost at once (6: nearly every ray of the frame goes through the queue)."""
    import os
    import torch
    shape, width, height = (96, 128, 160), 512, 384
    D, H, W = shape
    opt = VolumeOptions(**TF_SETS[0])
    tfu = capi.transfer_function_uniform(opt)
    vol = capi.Volume(ctx, W, H, D)
    vol.upload(scene.blobs_volume(shape, seed=8, n_blobs=20))
    vol.compute_gradient_map(tfu)
    vol.update_transfer_function(opt, skip)
    it = scene.image_transform((0.004,) * 3, (W, H, D))
    ropt = RenderOptions(skipping_type=skip, clip_distance=5.0)
    stream = torch.cuda.current_stream().cuda_stream
    results = {}
    try:
        for T in ("0", "64", "6"):
            os.environ["VKV_RC_LONG_T"] = T
            os.environ["VKV_RC_LONG_ALWAYS"] = "1"
            frames = []
            for eye in ((70, 40, 100), (72, 44, 98), (-60, 80, 90)):
                cu, ru = vol.make_uniforms(scene.look_at_camera(eye, aspect=width / height), it, 5.0)
                fb = torch.zeros((height, width, 4), dtype=torch.uint8, device="cuda")
                depth = torch.zeros((height, width), dtype=torch.float32, device="cuda")
                counts = torch.zeros(4, dtype=torch.int64, device="cuda")
                vol.render(cu, ru, tfu, ropt, width, height, fb.data_ptr(), depth.data_ptr(), counts.data_ptr(), stream)
                torch.cuda.synchronize()
                frames.append((fb, depth, counts.tolist()))
            results[T] = frames
    finally:
        os.environ.pop("VKV_RC_LONG_T", None)
        os.environ.pop("VKV_RC_LONG_ALWAYS", None)
    for T in ("64", "6"):
        for (fa, da, ca), (fb_, db, cb) in zip(results["0"], results[T]):
            assert torch.equal(fa, fb_) and torch.equal(da, db) and ca == cb, T
    assert results["0"][0][2][0] > 100000 and results["0"][0][2][3] > 10000        # a real frame: samples taken, pixels covered
    vol.close()
