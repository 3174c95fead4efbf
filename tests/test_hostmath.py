"""CPU-only: the product's host maths of VolumeRenderSubpass::draw (vkv_make_uniforms_for_extent / vkv_make_uniforms,
vkvolume_b200/csrc/hostmath.cu; src/volume_render_subpass.cpp:221-249) against

  * the glm goldens (tests/golden/reference_outputs.npz: the reference's own host maths evaluated with its vendored glm), and
  * the oracle's orc_make_uniforms (itself pinned against those goldens),

at 2e-5 relative — over cameras outside / inside the box, rotated and anisotropic image transforms, node transforms and
clip distances.  No device is needed: this is matrix arithmetic on the host (no kernel is launched).
"""
import math
from pathlib import Path

import numpy as np
import pytest

import oracle_api as orc
import ref_cases as cases
from vkvolume_b200 import capi, scene

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "reference_outputs.npz")
CAM_FIELDS = ("view", "proj", "view_proj_inv", "model", "model_inv")
RAY_FIELDS = ("plane", "plane_tex", "cam_pos_tex", "block_size")


def _close(a, b, what):
    a, b = np.array(list(a), np.float64), np.array(list(b), np.float64)
    assert np.allclose(a, b, rtol=2e-5, atol=2e-5 * max(np.abs(b).max(), 1e-30)), (what, a, b)


@pytest.mark.parametrize("tag", ["outside", "inside"])
def test_make_uniforms_matches_glm_goldens(tag):
    s = cases.render_scene("default", tag == "inside")
    D, H, W = cases.RENDER_SHAPE
    cu, ru = capi.make_uniforms_for_extent((W, H, D), s["dim_b"], s["cam"], s["it"], s["clip"])
    for k in CAM_FIELDS:
        _close(getattr(cu, k), GOLD[f"uniforms/{tag}/{k}"], k)
    for k in RAY_FIELDS:
        _close(getattr(ru, k), GOLD[f"uniforms/{tag}/{k}"], k)
    assert ru.front_index == int(GOLD[f"uniforms/{tag}/front_index"])


def _cameras():
    rng = np.random.default_rng(1234)
    out = []
    for k in range(24):
        dim = [(64, 48, 40), (832, 832, 494), (1024, 1024, 795), (9, 10, 13), (4096, 4096, 2048)][k % 5]
        bs = [4, 4, 4, 3, 4][k % 5]
        voxel = [(0.004,) * 3, (0.001,) * 3, (0.0003, 0.0003, 0.0007), (0.01, 0.02, 0.015), (0.00025,) * 3][k % 5]
        axis_angle = [(1, 0, 0, 0), (1, 0, 0, 90), (0, 1, 0, 30), (1, 1, 0, 45), (0.3, -0.2, 0.9, 200)][(k // 5) % 5]
        size = 100.0 * max(v * e for v, e in zip(voxel, dim))
        inside = k % 3 == 0
        r = (0.1 if inside else 1.3) * size
        ang = 0.35 + 0.9 * k
        eye = (r * math.cos(ang), 0.3 * r * math.sin(1.7 * ang) + (0 if inside else 0.2 * size), r * math.sin(ang))
        aspect = [16 / 9, 1.0, 4 / 3][k % 3]
        cam = scene.look_at_camera(eye, target=tuple(rng.uniform(-0.05, 0.05, 3) * size), aspect=aspect, yfov=[1.0, 0.6, 1.3][k % 3],
                                   znear=[1.0, 0.1][k % 2], zfar=[4000.0, 900.0][k % 2])
        if k % 4 == 1:        # a volume node that is translated, rotated and non-uniformly scaled (benchmark mode rescales the node)
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            cam.node_translation[:] = [float(x) for x in rng.uniform(-20, 20, 3)]
            cam.node_rotation[:] = [float(x) for x in q]
            cam.node_scale[:] = [float(x) for x in rng.uniform(40, 160, 3)]
        clip = [5.0, 50.0, 1.0, 0.02 * size][k % 4]
        out.append((dim, bs, voxel, axis_angle, cam, clip))
    return out


@pytest.mark.parametrize("case", range(24))
def test_make_uniforms_matches_oracle(case):
    dim, bs, voxel, axis_angle, cam, clip = _cameras()[case]
    dim_b, _ = orc.map_extent(dim, bs)
    it = scene.image_transform(voxel, dim, axis_angle)
    cu, ru = capi.make_uniforms_for_extent(dim, dim_b, cam, it, clip)
    ocu, oru = orc.make_uniforms(dim, dim_b, cam, it, clip)
    for k in CAM_FIELDS:
        _close(getattr(cu, k), getattr(ocu, k), k)
    for k in RAY_FIELDS:
        _close(getattr(ru, k), getattr(oru, k), k)
    assert ru.front_index == oru.front_index
    # block_size is the EFFECTIVE block size re-derived from the extents (volume_render_subpass.cpp:243-249), an exact integer
    assert [ru.block_size[a] for a in range(3)] == [float(-(-dim[a] // dim_b[a])) for a in range(3)]


def test_make_uniforms_argument_checks():
    cam = scene.look_at_camera((1, 2, 3))
    it = scene.image_transform((0.004,) * 3, (8, 8, 8))
    with pytest.raises(capi.VkvError):
        capi.make_uniforms_for_extent((8, 8, 8), (2, 0, 2), cam, it, 1.0)
