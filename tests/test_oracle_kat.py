"""CPU tests: known-answer and property tests that pin the oracle (tests/oracle_api.py ->
oracle/vkv_oracle.c) against the shader semantics of SURVEY.md Appendix A, independent numpy
restatements, brute-force closed forms and hand-computable cases (Appendix C matrix).
The reference ships no tests for this path; these are ours.
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_api as orc
from vkvolume_b200 import scene
from vkvolume_b200.capi import (RenderOptions, VolumeOptions, SKIP_ANISOTROPIC_DISTANCE, SKIP_BLOCK, SKIP_DISTANCE,
                                SKIP_NONE, TEST_NUM_TEXTURE_SAMPLES, TEST_RAY_ENTRY, TEST_RAY_EXIT)

f32 = np.float32

TF_SETS = [
    dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2),      # reference defaults
    dict(intensity_min=0.086, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0),    # beetle, no gradient
    dict(intensity_min=0.4, intensity_max=0.8, gradient_min=0.06, gradient_max=0.12),    # snake-like window
    dict(intensity_min=0.0, intensity_max=1.0, gradient_min=0.1, gradient_max=0.3),
]


def np_tf_texture(o):
    """Independent numpy restatement of Volume::update_transfer_function_texture (volume_component.cpp:242-261)."""
    i = np.arange(256, dtype=f32)[None, :]
    g = np.arange(256, dtype=f32)[:, None]
    i_inv = f32(1.0) / (f32(o["intensity_max"]) - f32(o["intensity_min"]))
    use_g = f32(o["gradient_max"]) != f32(o["gradient_min"])
    ai = np.clip((i / f32(255.0) - f32(o["intensity_min"])) * i_inv, f32(0), f32(1)).astype(f32)
    if use_g:
        g_inv = f32(1.0) / (f32(o["gradient_max"]) - f32(o["gradient_min"]))
        ag = np.clip((g / f32(255.0) - f32(o["gradient_min"])) * g_inv, f32(0), f32(1)).astype(f32)
    else:
        ag = np.ones_like(g)
    a = np.clip((ai * ag).astype(f32) * f32(255), f32(0), f32(255)).astype(np.uint8)        # truncation
    return np.repeat(a[:, :, None], 4, axis=2)


@pytest.mark.parametrize("o", TF_SETS)
def test_tf_texture_all_texels(o):
    tex = orc.transfer_function_texture(VolumeOptions(**o))
    assert np.array_equal(tex, np_tf_texture(o))
    u = orc.transfer_function_uniform(VolumeOptions(**o))
    assert bool(u.use_gradient) == (f32(o["gradient_max"]) != f32(o["gradient_min"]))
    assert u.grad_magnitude_modifier == 1.0
    assert f32(u.intensity_range_inv) == f32(1.0) / (f32(o["intensity_max"]) - f32(o["intensity_min"]))


# ---- loader ---------------------------------------------------------------------------------
def test_parse_header_with_comments():
    text = "832 832 494 # extents\n0.001 0.001 0.001 # voxel size\n400.0 2538.0 # normalisation range\n" \
           "uint16_t little # data type and endianness (big or little)\n1 0 0 90 # rotation axis and angle (degrees)\n"
    h = orc.parse_header(text)
    assert tuple(h.extent) == (832, 832, 494)
    assert h.type == b"uint16_t" and h.endianness == b"little"
    assert tuple(h.normalisation_range) == (400.0, 2538.0)
    M = np.array(list(h.image_transform), dtype=np.float64).reshape(4, 4).T        # column-major -> [r, c]
    # rotate 90 deg about x then scale by physical size: x -> x*0.832, y -> z*0.832, z -> -y ... (R*S)
    assert M[0, 0] == pytest.approx(0.832, rel=1e-6)
    assert M[2, 1] == pytest.approx(0.832, rel=1e-6) and abs(M[1, 1]) < 1e-6
    assert M[1, 2] == pytest.approx(-0.494, rel=1e-6)


@pytest.mark.parametrize("type_name,dtype", [("uint8_t", np.uint8), ("int8_t", np.int8), ("uint16_t", np.uint16), ("int16_t", np.int16)])
@pytest.mark.parametrize("endian", ["little", "big"])
def test_normalise_matches_formula(type_name, dtype, endian):
    rng = np.random.default_rng(3)
    info = np.iinfo(dtype)
    v = rng.integers(info.min, info.max + 1, size=5000).astype(dtype)
    raw = v.astype(v.dtype.newbyteorder(">" if endian == "big" else "<"))
    lo, hi = (400.0, 2538.0) if dtype in (np.uint16, np.int16) else (10.0, 200.0)
    got = orc.normalise(raw.view(np.uint8), v.size, type_name, endian, lo, hi)
    t = (v.astype(f32) - f32(lo)) / (f32(hi) - f32(lo))
    want = (f32(255) * np.clip(t, f32(0), f32(1))).astype(np.uint8)        # truncation, load_volume.cpp:165-169
    assert np.array_equal(got, want)


def test_normalise_rejects_unknown_type():
    with pytest.raises(ValueError):
        orc.normalise(np.zeros(4, np.uint8), 4, "float", "little", 0, 1)


# ---- K1 gradient -----------------------------------------------------------------------------
def np_gradient(V):
    """Independent numpy restatement of get_gradient_compute.glsl:12-20 (fp32, same summation order)."""
    D, H, W = V.shape
    v = V.astype(f32) / f32(255.0)
    zi, yi, xi = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")

    def tap(dx, dy, dz):
        return v[np.clip(zi + dz, 0, D - 1), np.clip(yi + dy, 0, H - 1), np.clip(xi + dx, 0, W - 1)]
    a, b, c, d = tap(1, -1, -1), tap(-1, -1, 1), tap(-1, 1, -1), tap(1, 1, 1)
    gx = f32(0.25) * (((a - b) - c) + d)
    gy = f32(0.25) * (((-a - b) + c) + d)
    gz = f32(0.25) * (((-a + b) - c) + d)
    return np.clip(np.sqrt(((gx * gx + gy * gy) + gz * gz).astype(f32)), f32(0), f32(1)).astype(f32)


def test_gradient_constant_volume_is_zero():
    V = np.full((5, 6, 7), 93, np.uint8)
    assert not orc.gradient_map(V).any()


def test_gradient_disabled_is_all_ones():
    V = scene.blobs_volume((6, 7, 9), seed=2)
    assert (orc.gradient_map(V, use_gradient=False) == 255).all()


def test_gradient_x_ramp_by_hand():
    # V = 10*x: interior taps give (a-b-c+d) = 2*(20/255) -> gx = 0.25*40/255, gy = gz = 0
    W = 9
    V = np.broadcast_to((10 * np.arange(W)).astype(np.uint8), (5, 5, W)).copy()
    G, Gf = orc.gradient_map(V, want_float=True)
    assert Gf[2, 2, 4] == pytest.approx(10.0 / 255.0, rel=1e-6)
    assert G[2, 2, 4] == round(255 * 10.0 / 255.0)
    # x border: clamped tap repeats the edge voxel -> half the slope
    assert Gf[2, 2, 0] == pytest.approx(5.0 / 255.0, rel=1e-6)


def test_gradient_random_vs_numpy():
    V = np.random.default_rng(0).integers(0, 256, size=(17, 17, 17), dtype=np.uint8)
    G, Gf = orc.gradient_map(V, want_float=True)
    ref = np_gradient(V)
    assert np.allclose(Gf, ref, rtol=1e-5, atol=1e-7)        # north-star tolerance on the float value
    t = ref * f32(255.0)
    tie = np.abs(t - np.floor(t) - 0.5) < 1e-3
    assert np.array_equal(G[~tie], np.rint(t).astype(np.uint8)[~tie])        # byte-exact outside the tie zone


# ---- K2a occupancy -----------------------------------------------------------------------------
def np_occupancy(V, G, tf, bs_req, use_gradient):
    D, H, W = V.shape
    (Wb, Hb, Db), (bx, by, bz) = orc.map_extent((W, H, D), bs_req)
    vis = tf[(G if use_gradient else np.full_like(V, 255)).astype(int), V.astype(int), 3] > 0
    O = np.full((Db, Hb, Wb), 255, np.uint8)
    for z in range(Db):
        for y in range(Hb):
            for x in range(Wb):
                if vis[z * bz:(z + 1) * bz, y * by:(y + 1) * by, x * bx:(x + 1) * bx].any():
                    O[z, y, x] = 0
    return O


@pytest.mark.parametrize("shape,bs", [((9, 10, 13), 4), ((8, 8, 16), 4), ((7, 9, 10), 3), ((12, 5, 6), 2), ((6, 6, 6), 8), ((10, 10, 10), 1)])
@pytest.mark.parametrize("o", TF_SETS[:3])
def test_occupancy_vs_numpy(shape, bs, o):
    V = scene.blobs_volume(shape, seed=sum(shape) + bs, n_blobs=3)
    G = orc.gradient_map(V)
    opt = VolumeOptions(**o)
    tf = orc.transfer_function_texture(opt)
    use_g = bool(orc.transfer_function_uniform(opt).use_gradient)
    assert np.array_equal(orc.occupancy_map(V, G, tf, bs, use_g), np_occupancy(V, G, tf, bs, use_g))


def test_effective_block_size_is_rederived():
    # dim 9, requested 4 -> 3 blocks -> effective block size 3
    dim_b, bs = orc.map_extent((9, 10, 494), 4)
    assert dim_b == (3, 3, 124) and bs == (3, 4, 4)


def test_occupancy_empty_full_and_corners():
    opt = VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
    tf = orc.transfer_function_texture(opt)
    assert (orc.occupancy_map(np.zeros((8, 8, 8), np.uint8), None, tf, 4, False) == 255).all()
    assert (orc.occupancy_map(np.full((8, 8, 8), 255, np.uint8), None, tf, 4, False) == 0).all()
    for corner in [(0, 0, 0), (3, 3, 3), (0, 3, 0), (3, 0, 3), (4, 4, 4), (7, 7, 7), (4, 7, 4), (7, 4, 7)]:
        V = np.zeros((8, 8, 8), np.uint8)
        V[corner] = 200
        O = orc.occupancy_map(V, None, tf, 4, False)
        want = np.full((2, 2, 2), 255, np.uint8)
        want[tuple(c // 4 for c in corner)] = 0
        assert np.array_equal(O, want)


# ---- K2b count -----------------------------------------------------------------------------------
@pytest.mark.parametrize("o", TF_SETS)
def test_count_vs_numpy_and_dispatch(o):
    V = scene.blobs_volume((11, 19, 21), seed=5, n_blobs=4)
    G = orc.gradient_map(V)
    opt = VolumeOptions(**o)
    u = orc.transfer_function_uniform(opt)
    v = V.astype(f32) / f32(255)
    g = G.astype(f32) / f32(255) if u.use_gradient else np.ones(V.shape, f32)
    with np.errstate(all="ignore"):
        aI = np.clip((v - f32(u.intensity_min)) * f32(u.intensity_range_inv), f32(0), f32(1))
        aG = np.clip((g - f32(u.gradient_min)) * f32(u.gradient_range_inv), f32(0), f32(1))
    want = int(((aI * aG) > 0).sum())
    assert orc.occupied_voxel_count(V, G, u) == want
    for sg in (32, 64, 8):        # the strided reduce of occupied_voxel_count_reduce.comp sums to the same total
        assert orc.occupied_voxel_count(V, G, u, dispatch_subgroup=sg) == want


def test_count_and_texture_tf_disagree_where_alpha_truncates():
    # analytic alpha in (0, 1/255) is visible to the counter but truncated to 0 in the texture
    opt = VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
    u = orc.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    V = np.full((4, 4, 4), 26, np.uint8)        # 26/255 = 0.10196 > 0.1 but alpha*255 < 1
    assert tf[255, 26, 3] == 0
    assert orc.occupied_voxel_count(V, None, u) == 64
    assert (orc.occupancy_map(V, None, tf, 4, False) == 255).all()


# ---- K3 distance maps ------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,p", [((8, 9, 10), 0.05), ((16, 12, 20), 0.01), ((5, 24, 7), 0.1), ((1, 1, 30), 0.1), ((3, 1, 1), 0.5)])
def test_distance_literal_equals_closed_form(shape, p):
    rng = np.random.default_rng(hash(shape) % 1000)
    O = np.where(rng.random(shape) < p, 0, 255).astype(np.uint8)
    assert np.array_equal(orc.distance_map(O), orc.distance_map_closed_form(O))
    D8 = orc.distance_map_anisotropic(O)
    for i in range(8):
        assert np.array_equal(D8[i], orc.distance_map_closed_form(O, i)), f"octant {i}"
    assert np.array_equal(D8.min(axis=0), orc.distance_map(O))


def test_distance_single_block_all_empty_and_saturation():
    O = np.full((9, 9, 9), 255, np.uint8)
    assert (orc.distance_map(O) == 255).all()        # nothing occupied -> 255 everywhere
    O[4, 4, 4] = 0
    D = orc.distance_map(O)
    z, y, x = np.meshgrid(*[np.arange(9)] * 3, indexing="ij")
    assert np.array_equal(D, np.maximum(np.maximum(abs(z - 4), abs(y - 4)), abs(x - 4)).astype(np.uint8))
    # a line longer than 255 saturates at 255
    L = np.full((1, 1, 300), 255, np.uint8)
    L[0, 0, 0] = 0
    D = orc.distance_map(L)[0, 0]
    assert np.array_equal(D, np.minimum(np.arange(300), 255).astype(np.uint8))
    D8 = orc.distance_map_anisotropic(L)
    assert (D8[0][0, 0, 1:] == 255).all() and D8[0][0, 0, 0] == 0        # +x octant sees nothing ahead
    assert np.array_equal(D8[4][0, 0], np.minimum(np.arange(300), 255).astype(np.uint8))        # -x octant


# ---- host maths ---------------------------------------------------------------------------------------
def test_uniforms_geometry():
    W, H, D = 80, 64, 48
    it = scene.image_transform((0.004, 0.004, 0.004), (W, H, D))
    cam = scene.look_at_camera((60, 40, 90), aspect=1.5)
    dim_b, _ = orc.map_extent((W, H, D), 4)
    cu, ru = orc.make_uniforms((W, H, D), dim_b, cam, it, 7.0)
    view = np.array(list(cu.view)).reshape(4, 4).T
    model = np.array(list(cu.model)).reshape(4, 4).T
    model_inv = np.array(list(cu.model_inv)).reshape(4, 4).T
    assert np.allclose(model @ model_inv, np.eye(4), atol=1e-4)
    eye = np.linalg.inv(view)[:3, 3]
    assert np.allclose(eye, (60, 40, 90), atol=1e-3)
    # camera in texture space = model^-1 * eye + 0.5
    assert np.allclose(list(ru.cam_pos_tex)[:3], (model_inv @ np.append(eye, 1))[:3] + 0.5, rtol=1e-5)
    # plane_tex evaluated at a texture-space point equals plane evaluated at its world position
    p_tex = np.array([0.3, 0.6, 0.2, 1.0])
    p_world = model @ np.append(p_tex[:3] - 0.5, 1)
    assert np.dot(list(ru.plane_tex), p_tex) == pytest.approx(np.dot(list(ru.plane), p_world), rel=1e-4, abs=1e-3)
    # the plane is clip_distance in front of the camera
    assert np.dot(list(ru.plane), np.append(eye, 1)) == pytest.approx(-7.0, abs=1e-3)
    # reverse-Z, Y-flipped projection: near plane -> depth 1, far plane -> depth 0
    proj = np.array(list(cu.proj)).reshape(4, 4).T
    near = proj @ np.array([0, 0, -cam.znear, 1.0])
    far = proj @ np.array([0, 0, -cam.zfar, 1.0])
    assert near[2] / near[3] == pytest.approx(1.0, abs=1e-5) and far[2] / far[3] == pytest.approx(0.0, abs=1e-5)
    assert proj[1, 1] < 0 and proj[0, 0] == pytest.approx(1.0 / (1.5 * math.tan(0.5)), rel=1e-6)
    assert list(ru.block_size)[:3] == [4.0, 4.0, 4.0]


# ---- K4 ray caster ----------------------------------------------------------------------------------------
def _scene(shape=(48, 64, 80), opt=None, clip=5.0, eye=(34, 22, 50), aspect=1.0):
    D, H, W = shape
    V = scene.blobs_volume(shape, seed=1)
    opt = opt or VolumeOptions(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2)
    tfu = orc.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    G = orc.gradient_map(V, bool(tfu.use_gradient))
    O = orc.occupancy_map(V, G, tf, 4, bool(tfu.use_gradient))
    dim_b, _ = orc.map_extent((W, H, D), 4)
    it = scene.image_transform((0.004, 0.004, 0.004), (W, H, D))
    cu, ru = orc.make_uniforms((W, H, D), dim_b, scene.look_at_camera(eye, aspect=aspect), it, clip)
    return dict(V=V, G=G, tf=tf, tfu=tfu, O=O, dim_b=dim_b, cu=cu, ru=ru)


def test_render_uniform_volume_closed_form_opacity():
    # constant volume, constant TF alpha a per sample: out.a = 1 - (1-a)^n, out.rgb = a_colour * out.a
    shape = (16, 16, 16)
    V = np.full(shape, 128, np.uint8)
    tf = np.zeros((256, 256, 4), np.uint8)
    tf[..., :] = 8        # alpha 8/255 everywhere
    opt = VolumeOptions(intensity_min=0.0, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0)
    tfu = orc.transfer_function_uniform(opt)
    it = scene.image_transform((0.01, 0.01, 0.01), (16, 16, 16))
    cu, ru = orc.make_uniforms((16, 16, 16), (4, 4, 4), scene.look_at_camera((0, 0, 60)), it, 1.0)
    ropt = RenderOptions(skipping_type=SKIP_NONE, clip_distance=1.0, early_ray_termination=0)
    img, counts, rf, _ = orc.render(V, None, tf, None, (4, 4, 4), cu, ru, tfu, ropt, 64, 64, want_float=True)
    cov = rf[..., 3] >= 0
    assert cov.sum() > 100
    # centre pixel: ray along -z through the whole cube: len = 1, n = ceil(16 * 1 * 1) samples (step = len/(n-1))
    a = f32(8) / f32(255)
    c = rf[32, 32]
    # the pixel centre is half a pixel off the axis, so len is a hair over 1 and n = ceil(16 * len) is 16 or 17
    n = round(math.log(1 - float(c[3])) / math.log(1 - float(a)))
    assert n in (16, 17)
    assert c[3] == pytest.approx(1 - (1 - float(a)) ** n, rel=1e-4)
    assert c[0] == pytest.approx(float(a) * c[3], rel=1e-4)
    # stored alpha is a*(1-a) (blend state quirk), RGB is sRGB encoded
    assert img[32, 32, 3] == int(c[3] * (1 - c[3]) * 255 + 0.5)
    lin = float(c[0])
    srgb = 12.92 * lin if lin <= 0.0031308 else 1.055 * lin ** (1 / 2.4) - 0.055
    assert img[32, 32, 0] == int(srgb * 255 + 0.5)
    # uncovered pixels keep the clear colour (0,0,0,255)
    assert (img[~cov] == (0, 0, 0, 255)).all()


def test_render_skip_modes_agree_and_save_samples():
    s = _scene()
    Dm = orc.distance_map(s["O"])
    D8 = orc.distance_map_anisotropic(s["O"])
    base = None
    samples = {}
    for skip, maps in [(SKIP_NONE, None), (SKIP_BLOCK, s["O"]), (SKIP_DISTANCE, Dm), (SKIP_ANISOTROPIC_DISTANCE, D8)]:
        ropt = RenderOptions(skipping_type=skip, clip_distance=5.0)
        img, c, rf, _ = orc.render(s["V"], s["G"], s["tf"], maps, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, 96, 96, want_float=True)
        samples[skip] = c.volume_samples + c.distance_samples
        if base is None:
            base = rf
            assert c.covered_pixels > 500
        else:
            mse = float(np.mean((rf[..., :3] - base[..., :3]) ** 2))
            assert mse < 1e-5        # ESS may change the picture a little, not much
    assert samples[SKIP_BLOCK] < samples[SKIP_NONE]
    assert samples[SKIP_DISTANCE] < samples[SKIP_BLOCK]
    assert samples[SKIP_ANISOTROPIC_DISTANCE] <= samples[SKIP_DISTANCE]


def test_render_ert_changes_alpha_by_less_than_threshold():
    opt = VolumeOptions(intensity_min=0.05, intensity_max=0.3, gradient_min=0.0, gradient_max=0.0, voxel_alpha_factor=2.0)
    s = _scene(opt=opt)
    outs = []
    for ert in (1, 0):
        ropt = RenderOptions(skipping_type=SKIP_NONE, clip_distance=5.0, early_ray_termination=ert)
        _, c, rf, _ = orc.render(s["V"], s["G"], s["tf"], None, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, 64, 64, want_float=True)
        outs.append((rf, c.volume_samples))
    assert outs[0][1] < outs[1][1]        # ERT saves samples
    assert np.abs(outs[0][0][..., 3] - outs[1][0][..., 3]).max() <= 0.01 + 1e-6


def test_render_entry_exit_views_and_clip_plane():
    s = _scene(clip=5.0)
    ropt = RenderOptions(skipping_type=SKIP_NONE, clip_distance=5.0, test=TEST_RAY_ENTRY)
    _, c, entry, _ = orc.render(s["V"], s["G"], s["tf"], None, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, 96, 96, want_float=True)
    ropt.test = TEST_RAY_EXIT
    _, _, exit_, _ = orc.render(s["V"], s["G"], s["tf"], None, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, 96, 96, want_float=True)
    cov = entry[..., 3] >= 0
    e, x = entry[cov][:, :3], exit_[cov][:, :3]
    eps = 1e-4
    assert (e > -eps).all() and (e < 1 + eps).all() and (x > -eps).all() and (x < 1 + eps).all()
    # camera outside, plane in front of the box: entries lie on a box face, exits too
    on_face = lambda p: (np.minimum(np.abs(p), np.abs(1 - p)).min(axis=1) < 1e-3)
    assert on_face(e).mean() > 0.99 and on_face(x).mean() > 0.99
    # camera pushed inside the box with a near clip plane: entries lie on the plane
    s2 = _scene(clip=3.0, eye=(0, 0, 10))
    ropt = RenderOptions(skipping_type=SKIP_NONE, clip_distance=3.0, test=TEST_RAY_ENTRY)
    _, c2, entry2, _ = orc.render(s2["V"], s2["G"], s2["tf"], None, s2["dim_b"], s2["cu"], s2["ru"], s2["tfu"], ropt, 64, 64, want_float=True)
    cov2 = entry2[..., 3] >= 0
    assert cov2.all()        # inside the box every pixel is covered
    pl = np.array(list(s2["ru"].plane_tex))
    d = entry2[cov2][:, :3] @ pl[:3] + pl[3]
    assert np.abs(d).max() < 1e-2 * np.linalg.norm(pl[:3])


def test_render_num_samples_mode():
    s = _scene()
    Dm = orc.distance_map(s["O"])
    ropt = RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0, early_ray_termination=0, test=TEST_NUM_TEXTURE_SAMPLES)
    img, c, rf, _ = orc.render(s["V"], s["G"], s["tf"], Dm, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, 64, 64, want_float=True)
    n_max = int(math.ceil(80 * math.sqrt(3.0)) * 1.0)
    cov = rf[..., 3] >= 0
    # the image integrates to the counters: sum(rgb * n_max) == n_vol + n_dist over pixels that marched
    total = float((rf[..., 0][cov & (rf[..., 3] == 1.0)] * n_max).sum())
    assert total == pytest.approx(c.volume_samples + c.distance_samples, rel=1e-4)


def test_worked_skip_example_from_survey():
    # SURVEY A.6 worked example: step = +0.1 block/iter, u.x = 3.25, dist = 2 -> 18 iterations; -0.1 -> 13; block mode -> 8
    def delta(sdt, u, u_i, dist, block):
        sdi = f32(1) / f32(sdt)
        r = max(min(f32(u_i) - f32(u), f32(0)), f32(-1))
        if block:
            v = ((f32(0) if sdi < 0 else f32(1)) + r) * sdi
        else:
            st = f32(0) if -sdi < 0 else f32(1)
            sg = f32(1) if sdi > 0 else f32(-1)
            v = (st + sg * f32(dist) + r) * sdi
        return max(1, int(math.ceil(float(v))))
    assert delta(0.1, 3.25, 3, 2, False) == 18
    assert delta(-0.1, 3.25, 3, 2, False) == 13
    assert delta(0.1, 3.25, 3, 2, True) == 8


# ---- compositing over an existing target, depth attachment (volume_render_subpass.cpp:176-190, volume_render.frag:122-165) --------
def test_load_over_explicit_clear_equals_clear_and_far_depth_changes_nothing():
    s = _scene()
    Dm = orc.distance_map(s["O"])
    args = (s["V"], s["G"], s["tf"], Dm, s["dim_b"], s["cu"], s["ru"], s["tfu"])
    plain, c0, _, d0 = orc.render(*args, RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0), 96, 96, want_depth=True)
    clear = np.zeros((96, 96, 4), np.uint8)
    clear[..., 3] = 255
    zero = np.zeros((96, 96), np.float32)
    over, c1, _, d1 = orc.render(*args, RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0, load_framebuffer=1), 96, 96,
                                 rgba_init=clear, depth_init=zero)
    assert np.array_equal(over, plain) and np.array_equal(d1, d0)
    assert (c1.volume_samples, c1.distance_samples, c1.covered_pixels) == (c0.volume_samples, c0.distance_samples, c0.covered_pixels)
    # a depth attachment that holds "far" everywhere (0 in reverse-Z) neither discards nor shortens anything
    far, c2, rf, d2 = orc.render(*args, RenderOptions(skipping_type=SKIP_DISTANCE, clip_distance=5.0, load_framebuffer=1, depth_attachment=1),
                                 96, 96, want_float=True, rgba_init=clear, depth_init=zero)
    assert not (rf[..., 3] == -2.0).any()
    assert np.array_equal(far, plain) and np.array_equal(d2, d0) and c2.volume_samples == c0.volume_samples


def test_depth_attachment_in_front_of_the_volume_discards_everything_and_blend_is_over():
    s = _scene()
    args = (s["V"], s["G"], s["tf"], None, s["dim_b"], s["cu"], s["ru"], s["tfu"])
    rng = np.random.default_rng(2)
    bg = rng.integers(0, 256, size=(64, 64, 4), dtype=np.uint8)
    near = np.ones((64, 64), np.float32)        # reverse-Z: 1 = the near plane, in front of every fragment
    img, c, rf, dp = orc.render(*args, RenderOptions(skipping_type=SKIP_NONE, clip_distance=5.0, load_framebuffer=1, depth_attachment=1), 64, 64,
                                want_float=True, rgba_init=bg, depth_init=near)
    cov = rf[..., 3] != -1.0
    assert cov.sum() > 200 and (rf[..., 3][cov] == -2.0).all()        # every covered fragment discarded
    assert np.array_equal(img, bg) and np.array_equal(dp, near) and c.volume_samples == 0
    # the blend itself: rgb = src.rgb + dst.rgb * (1 - src.a) on linear values of an sRGB target, a = src.a * (1 - src.a)
    img2, _, rf2, _ = orc.render(*args, RenderOptions(skipping_type=SKIP_NONE, clip_distance=5.0, load_framebuffer=1), 64, 64,
                                 want_float=True, rgba_init=bg, depth_init=np.zeros((64, 64), np.float32))
    def dec(b):
        c_ = b.astype(np.float64) / 255.0
        return np.where(c_ <= 0.04045, c_ / 12.92, ((c_ + 0.055) / 1.055) ** 2.4)
    def enc(l):
        l = np.clip(l, 0, 1)
        return np.where(l <= 0.0031308, 12.92 * l, 1.055 * l ** (1 / 2.4) - 0.055)
    sa = np.clip(rf2[..., 3:4].astype(np.float64), 0, 1)
    want = np.floor(np.clip(enc(np.clip(rf2[..., :3], 0, 1) + dec(bg[..., :3]) * (1 - sa)), 0, 1) * 255 + 0.5)
    got = img2[..., :3].astype(np.float64)
    assert np.abs(got[cov] - want[cov]).max() <= 1        # fp32 vs fp64 evaluation of the same formula
    assert np.array_equal(img2[~cov], bg[~cov])
    a_want = np.floor(sa[..., 0] * (1 - sa[..., 0]) * 255 + 0.5)
    assert np.abs(img2[..., 3].astype(np.float64)[cov] - a_want[cov]).max() <= 1
