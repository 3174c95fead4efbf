"""Seeded cases shared by tests/golden/make_golden.py (runs the reference's own sources through oracle/_ref and
stores their outputs) and tests/test_oracle_vs_reference.py (checks the oracle against those outputs)."""
from __future__ import annotations

import numpy as np

import oracle_api as orc
from vkvolume_b200 import scene
from vkvolume_b200.capi import VolumeOptions

TF_SETS = {
    "default": dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2),
    "beetle_nograd": dict(intensity_min=0.086, intensity_max=1.0, gradient_min=0.0, gradient_max=0.0),
    "snake_window": dict(intensity_min=0.4, intensity_max=0.8, gradient_min=0.06, gradient_max=0.12),
}

VOLUME_CASES = {        # name: (shape [D,H,W], block size, seed)
    "v24": ((20, 24, 28), 4, 11),
    "v_odd": ((9, 10, 13), 4, 12),
    "v_bs3": ((12, 14, 16), 3, 13),
}

DIST_CASES = {          # name: (shape [Db,Hb,Wb], occupied probability, seed)
    "d_small": ((8, 9, 10), 0.05, 1),
    "d_sparse": ((12, 20, 33), 0.004, 2),
    "d_line": ((1, 1, 300), 0.004, 3),
    "d_empty": ((5, 6, 7), 0.0, 4),
    "d_tall": ((3, 270, 2), 0.003, 5),
}

RENDER_SHAPE = (40, 48, 64)


def volume(name):
    shape, bs, seed = VOLUME_CASES[name]
    return scene.blobs_volume(shape, seed=seed, n_blobs=5), bs


def occupancy_grid(name):
    shape, p, seed = DIST_CASES[name]
    rng = np.random.default_rng(seed)
    O = np.where(rng.random(shape) < p, 0, 255).astype(np.uint8)
    if name == "d_line":
        O[0, 0, 0] = 0
    return O


def render_scene(tf_name="default", inside=False):
    D, H, W = RENDER_SHAPE
    V = scene.blobs_volume(RENDER_SHAPE, seed=21, n_blobs=8)
    opt = VolumeOptions(**TF_SETS[tf_name])
    tfu = orc.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    G = orc.gradient_map(V, bool(tfu.use_gradient))
    O = orc.occupancy_map(V, G, tf, 4, bool(tfu.use_gradient))
    dim_b, _ = orc.map_extent((W, H, D), 4)
    it = scene.image_transform((0.004, 0.004, 0.007), (W, H, D), (1, 0, 0, 90))
    eye, clip = ((2.0, 1.5, 4.0), 3.0) if inside else ((30.0, 21.0, 44.0), 5.0)
    cam = scene.look_at_camera(eye, aspect=4 / 3)
    cu, ru = orc.make_uniforms((W, H, D), dim_b, cam, it, clip)
    return dict(V=V, G=G, tf=tf, tfu=tfu, O=O, Dm=orc.distance_map(O), D8=orc.distance_map_anisotropic(O), dim_b=dim_b, cu=cu, ru=ru,
                cam=cam, it=it, clip=clip, opt=opt)


def ray_entries(s, width=40, height=30):
    """ray_entry varyings of the covered pixels of a small frame (from the oracle's analytic entry, TEST_RAY_ENTRY view)."""
    from vkvolume_b200.capi import RenderOptions, TEST_RAY_ENTRY
    ropt = RenderOptions(skipping_type=0, clip_distance=s["clip"], test=TEST_RAY_ENTRY)
    _, _, rf, _ = orc.render(s["V"], s["G"], s["tf"], None, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, width, height, want_float=True)
    cov = rf[..., 3] >= 0
    return rf[cov][:, :3].copy(), cov


def entry_positions(s, entries):
    """The `position` varying of each ray entry: gl_Position = proj * view * model * (ray_entry - 0.5)
    (volume_render_clipped.vert:58-62), evaluated like the oracle does (fp64 product, rounded once)."""
    m = lambda a: np.array(list(a), np.float32).astype(np.float64).reshape(4, 4).T        # column-major uniforms
    pvm = (m(s["cu"].proj) @ m(s["cu"].view)) @ m(s["cu"].model)
    pm = np.concatenate([entries.astype(np.float64) - 0.5, np.ones((len(entries), 1))], axis=1)
    return (pm @ pvm.T).astype(np.float32)


def depth_attachment_pattern(s, width=40, height=30):
    """A synthetic depth attachment for the DEPTH_ATTACHMENT variants, as a full frame: per pixel one of
    nothing behind the volume (0 = far in reverse-Z) | geometry just inside the front face | geometry deeper inside |
    geometry in front of the volume (the fragment is discarded)."""
    entries, cov = ray_entries(s, width, height)
    pos = entry_positions(s, entries)
    front = np.zeros((height, width), np.float32)
    front[cov] = pos[:, 2] / pos[:, 3]
    yy, xx = np.mgrid[0:height, 0:width]
    sel = (xx + 2 * yy) % 4
    factor = np.choose(sel, [0.0, 1.0 / 1.004, 1.0 / 1.02, 1.1]).astype(np.float32)
    depth = (front * factor).astype(np.float32)
    depth[~cov] = np.float32(0.25) * (sel[~cov] == 3)        # something also behind pixels the volume does not cover
    return depth, entries, pos, cov
