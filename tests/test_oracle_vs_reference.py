"""Pins the CPU oracle (oracle/vkv_oracle.c) against the REFERENCE ITSELF.

tests/golden/reference_outputs.npz holds outputs of the reference's own sources — the GLSL shaders compiled as C++
through oracle/ref_shim and executed with the reference's dispatch shapes, src/load_volume.cpp compiled as is, and the
host maths of volume_render_subpass.cpp evaluated with the reference's vendored glm — on the seeded cases of
tests/ref_cases.py (generator: tests/golden/make_golden.py).  Every test here recomputes the same case with the oracle
and compares: bit-exact for the integer/byte stages, tight fp32 tolerances for the fragment shader and matrices.
When oracle/_ref/libvkv_ref.so is present (build container) the goldens are additionally re-derived live.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import oracle_api as orc
import ref_api as ref
import ref_cases as cases
from vkvolume_b200.capi import RenderOptions, VolumeOptions

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "reference_outputs.npz")


def gold(key):
    return GOLD[key]


@pytest.mark.parametrize("vname", list(cases.VOLUME_CASES))
@pytest.mark.parametrize("tname", list(cases.TF_SETS))
def test_gradient_occupancy_count_match_reference_shaders(vname, tname):
    V, bs = cases.volume(vname)
    opt = VolumeOptions(**cases.TF_SETS[tname])
    tfu = orc.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    # K1: gradient_map.comp + get_gradient_compute.glsl
    G = orc.gradient_map(V, bool(tfu.use_gradient))
    assert np.array_equal(G, gold(f"grad/{vname}/{tname}"))
    # K2a: occupancy_map.comp, precomputed and on-the-fly gradient variants
    assert np.array_equal(orc.occupancy_map(V, G, tf, bs, bool(tfu.use_gradient), precomputed=True), gold(f"occ/{vname}/{tname}"))
    assert np.array_equal(orc.occupancy_map(V, None, tf, bs, bool(tfu.use_gradient), precomputed=False), gold(f"occ_otf/{vname}/{tname}"))
    # K2b + K2c: occupied_voxel_count.comp + occupied_voxel_count_reduce.comp for three subgroup sizes
    want = orc.occupied_voxel_count(V, G, tfu)
    for sg in (8, 32, 64):
        c0, partial = gold(f"count/{vname}/{tname}/s{sg}")
        assert int(c0) == want and int(partial) == want
        assert orc.occupied_voxel_count(V, G, tfu, dispatch_subgroup=sg) == int(c0)
    assert orc.occupied_voxel_count(V, None, tfu, precomputed=False) == int(gold(f"count_otf/{vname}/{tname}")[0])


@pytest.mark.parametrize("dname", list(cases.DIST_CASES))
def test_distance_maps_match_reference_shaders(dname):
    O = cases.occupancy_grid(dname)
    assert np.array_equal(orc.distance_map(O), gold(f"dist/{dname}"))
    assert np.array_equal(orc.distance_map_anisotropic(O), gold(f"dist8/{dname}"))
    # and the closed form the CUDA kernels implement
    if O.size <= 4000:
        assert np.array_equal(orc.distance_map_closed_form(O), gold(f"dist/{dname}"))
        for i in range(8):
            assert np.array_equal(orc.distance_map_closed_form(O, i), gold(f"dist8/{dname}")[i])


@pytest.mark.parametrize("tag", ["outside", "inside"])
def test_uniforms_match_glm(tag):
    s = cases.render_scene("default", tag == "inside")
    cu, ru = s["cu"], s["ru"]
    for k in ("view", "proj", "view_proj_inv", "model", "model_inv"):
        g = gold(f"uniforms/{tag}/{k}")
        assert np.allclose(np.array(list(getattr(cu, k))), g, rtol=2e-5, atol=2e-5 * np.abs(g).max()), k
    for k in ("plane", "plane_tex", "cam_pos_tex", "block_size"):
        g = gold(f"uniforms/{tag}/{k}")
        assert np.allclose(np.array(list(getattr(ru, k))), g, rtol=2e-5, atol=2e-5 * np.abs(g).max()), k
    assert ru.front_index == int(gold(f"uniforms/{tag}/front_index"))


@pytest.mark.parametrize("tag", ["outside", "inside"])
def test_analytic_entry_matches_vertex_shaders(tag):
    """The oracle's per-pixel analytic ray entry lies on the geometry the two vertex shaders emit:
    cube front faces (volume_render_clipped.vert) or the box/plane polygon (volume_render_plane_intersection.vert)."""
    s = cases.render_scene("default", tag == "inside")
    vc, vp = gold(f"vert_clipped/{tag}"), gold(f"vert_plane/{tag}")
    # clipped.vert: ray_entry = position + 0.5 (cube corners), clip distance = dot(plane, world position)
    corners = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32)
    assert np.allclose(vc[:, 4:7], corners)
    model = np.array(list(s["cu"].model)).reshape(4, 4).T
    world = (model @ np.c_[corners - 0.5, np.ones(8)].T).T
    assert np.allclose(vc[:, 7], world @ np.array(list(s["ru"].plane)), rtol=1e-4, atol=1e-3)
    entries, cov = cases.ray_entries(s)
    pl = np.array(list(s["ru"].plane_tex), np.float64)
    on_plane = np.abs(entries @ pl[:3] + pl[3]) < 2e-3 * np.linalg.norm(pl[:3])
    on_face = np.minimum(np.abs(entries), np.abs(1 - entries)).min(axis=1) < 1e-4
    assert (on_plane | on_face).all()
    # polygon vertices (those that are not the shader's NaN/inf "no intersection" marker) lie on the plane and on box edges
    ok = np.isfinite(vp[:, 4:7]).all(axis=1)
    if tag == "inside":
        assert ok.sum() >= 3 and on_plane.all()        # camera inside: every pixel enters through the clip polygon
        poly = vp[ok, 4:7].astype(np.float64)
        assert np.abs(poly @ pl[:3] + pl[3]).max() < 2e-3 * np.linalg.norm(pl[:3])
        # entries lie inside the polygon's bounding box on the plane
        assert (entries.min(axis=0) >= poly.min(axis=0) - 1e-3).all() and (entries.max(axis=0) <= poly.max(axis=0) + 1e-3).all()
    else:
        assert on_face.all()


SKIP_MAPS = {0: None, 1: "O", 2: "Dm", 3: "D8"}


@pytest.mark.parametrize("tname", ["default", "beetle_nograd", "snake_window"])
def test_fragment_shader_variants_match_reference(tname):
    """volume_render.frag, 16 #define variants + entry/exit views, fragment by fragment."""
    s = cases.render_scene(tname, False)
    entries, cov = cases.ray_entries(s)
    assert np.array_equal(entries, gold(f"frag_entries/{tname}"))
    W, H = 40, 30
    for skip in (0, 1, 2, 3):
        maps = None if skip == 0 else s[SKIP_MAPS[skip]]
        for ert in (0, 1):
            for test in (0, 3):
                ropt = RenderOptions(skipping_type=skip, clip_distance=s["clip"], early_ray_termination=ert, test=test)
                _, _, rf, dp = orc.render(s["V"], s["G"], s["tf"], maps, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, W, H, want_float=True, want_depth=True)
                got, want = rf[cov], gold(f"frag/{tname}/s{skip}_e{ert}_t{test}")
                assert np.abs(got - want).max() <= 2e-6, (skip, ert, test, float(np.abs(got - want).max()))
                assert np.allclose(dp[cov], gold(f"frag_depth/{tname}/s{skip}_e{ert}_t{test}"), rtol=1e-4, atol=1e-6)
    for test in (1, 2):
        ropt = RenderOptions(skipping_type=2, clip_distance=s["clip"], early_ray_termination=1, test=test)
        _, _, rf, _ = orc.render(s["V"], s["G"], s["tf"], s["Dm"], s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, W, H, want_float=True)
        assert np.abs(rf[cov] - gold(f"frag/{tname}/s2_e1_t{test}")).max() <= 2e-6
    for skip in (0, 2):        # on-the-fly gradient variant
        ropt = RenderOptions(skipping_type=skip, clip_distance=s["clip"], early_ray_termination=1)
        maps = None if skip == 0 else s["Dm"]
        _, _, rf, _ = orc.render(s["V"], None, s["tf"], maps, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, W, H, precomputed=False, want_float=True)
        assert np.abs(rf[cov] - gold(f"frag_otf/{tname}/s{skip}")).max() <= 2e-6


@pytest.mark.parametrize("tname", ["default", "beetle_nograd", "snake_window"])
def test_depth_attachment_variants_match_reference(tname):
    """volume_render.frag with DEPTH_ATTACHMENT (:122-136,151-165): discard behind the scene's depth, rays shortened at it."""
    s = cases.render_scene(tname, False)
    depth, entries, pos, cov = cases.depth_attachment_pattern(s)
    assert np.array_equal(depth, gold(f"frag_d1/{tname}/depth_in")) and np.array_equal(pos, gold(f"frag_d1/{tname}/position"))
    W, H = 40, 30
    clear = np.zeros((H, W, 4), np.uint8)
    clear[..., 3] = 255
    n_short = 0
    for skip, test in ((0, 0), (1, 0), (2, 0), (3, 0), (2, 2)):
        maps = None if skip == 0 else s[SKIP_MAPS[skip]]
        ropt = RenderOptions(skipping_type=skip, clip_distance=s["clip"], early_ray_termination=1, test=test, depth_attachment=1, load_framebuffer=1)
        _, _, rf, dp = orc.render(s["V"], s["G"], s["tf"], maps, s["dim_b"], s["cu"], s["ru"], s["tfu"], ropt, W, H, want_float=True,
                                  rgba_init=clear, depth_init=depth)
        got, disc = rf[cov], gold(f"frag_d1_discard/{tname}/s{skip}_t{test}").astype(bool)
        assert np.array_equal(got[:, 3] == -2.0, disc) and 0 < disc.sum() < len(disc)
        assert np.abs(got[~disc] - gold(f"frag_d1/{tname}/s{skip}_t{test}")[~disc]).max() <= 2e-6
        # depth written where the fragment survives = the shader's gl_FragDepth (>= the attachment's value: the test passes)
        assert np.allclose(dp[cov][~disc], gold(f"frag_d1_depth/{tname}/s{skip}_t{test}")[~disc], rtol=1e-4, atol=1e-6)
        assert np.array_equal(dp[cov][disc], depth[cov][disc]) and np.array_equal(dp[~cov], depth[~cov])        # untouched elsewhere
        if test == 2:
            full = gold(f"frag/{tname}/s2_e1_t2")
            n_short = int((np.abs(full[~disc, :3] - got[~disc, :3]).max(axis=1) > 1e-6).sum())
    assert n_short > 20        # the pattern really shortens rays


@pytest.mark.parametrize("tname", ["uint8_t", "int8_t", "uint16_t", "int16_t"])
@pytest.mark.parametrize("endian", ["little", "big"])
def test_loader_matches_reference_load_volume_cpp(tname, endian):
    raw = gold(f"loader/{tname}_{endian}/raw")
    lo, hi = (400.0, 2538.0) if "16" in tname else (10.0, 200.0)
    got = orc.normalise(raw, 11 * 7 * 5, tname, endian, lo, hi)
    assert np.array_equal(got, gold(f"loader/{tname}_{endian}/u8"))
    h = orc.parse_header(bytes(gold(f"loader/{tname}_{endian}/header_text")).decode())
    assert np.allclose(list(h.image_transform), gold(f"loader/{tname}_{endian}/image_transform"), rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libvkv_ref.so not built (needs /root/reference)")
def test_goldens_are_reproducible_from_the_reference_sources():
    """Re-runs a slice of the golden generation live so a stale .npz cannot hide a drift."""
    V, bs = cases.volume("v24")
    opt = VolumeOptions(**cases.TF_SETS["default"])
    tfu = orc.transfer_function_uniform(opt)
    assert np.array_equal(ref.gradient_map(V, tfu), gold("grad/v24/default"))
    O = cases.occupancy_grid("d_sparse")
    assert np.array_equal(ref.distance_map(O), gold("dist/d_sparse"))
    assert np.array_equal(ref.distance_map_anisotropic(O), gold("dist8/d_sparse"))
