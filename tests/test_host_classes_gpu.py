"""The reference-named C++ host classes (vkvolume_b200/host/vkvolume.h: LoadVolume, Volume, ComputeGradientMap,
ComputeDistanceMap, VolumeRenderSubpass, VolumeRender) driven end to end by host_selftest, a C++ program that follows the
reference application's sequence (src/volume_render.cpp:163-245), and what they PRODUCED — voxels after the loader, gradient
map, TF texture, every distance map, the frame and the depth image — compared with the oracle, byte for byte / at the frame bar."""
import math
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_api as orc
from vkvolume_b200 import capi, scene
from vkvolume_b200.capi import FILTER_HARDWARE, RenderOptions, VolumeOptions

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
APP = ROOT / "vkvolume_b200" / "lib" / "host_selftest"


def _write_dataset(path, V16, lo, hi, voxel, axis_angle, endian="little"):
    D, H, W = V16.shape
    V16.astype(V16.dtype.newbyteorder(">" if endian == "big" else "<")).tofile(path)
    Path(str(path) + ".header").write_text(
        f"{W} {H} {D} # extents\n{voxel[0]} {voxel[1]} {voxel[2]} # voxel size\n{lo} {hi} # normalisation\nuint16_t {endian} # type\n"
        f"{axis_angle[0]} {axis_angle[1]} {axis_angle[2]} {axis_angle[3]} # rotation\n")


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def _dataset(shape, seed, lo, hi):
    V8 = scene.blobs_volume(shape, seed=seed).astype(np.float64)
    rng = np.random.default_rng(seed)
    return np.clip(lo + V8 / 255.0 * (hi - lo) + rng.integers(-3, 4, size=shape), 0, 65535).astype(np.uint16)


@pytest.mark.parametrize("skipmode", [2, 3, 1])
@pytest.mark.parametrize("two_volumes", [False, True])
def test_host_classes_outputs_match_the_oracle(tmp_path, skipmode, two_volumes):
    assert APP.exists(), "host_selftest is not built (python -m vkvolume_b200.build)"
    if two_volumes and skipmode == 1:
        pytest.skip("one compositing case per distance-map mode is enough")
    shape, lo, hi = (48, 64, 80), 400.0, 2538.0
    voxel, axis_angle = (0.004, 0.004, 0.005), (1, 0, 0, 90)
    width, height, clip, bs = 192, 128, 5.0, 4
    tfo = dict(intensity_min=0.1, intensity_max=1.0, gradient_min=0.0, gradient_max=0.2)
    raw = _dataset(shape, 3, lo, hi)
    _write_dataset(tmp_path / "a.raw", raw, lo, hi, voxel, axis_angle, "big")
    datasets = [tmp_path / "a.raw"]
    raws = [(raw, "big")]
    if two_volumes:
        raw2 = _dataset(shape, 9, lo, hi)
        _write_dataset(tmp_path / "b.raw", raw2, lo, hi, voxel, axis_angle, "little")
        datasets.append(tmp_path / "b.raw")
        raws.append((raw2, "little"))
    cam = scene.look_at_camera((38, 24, 52), aspect=width / height)
    out = tmp_path / "out"
    out.mkdir()
    args = [str(APP), str(datasets[0]), str(out), tfo["intensity_min"], tfo["intensity_max"], tfo["gradient_min"], tfo["gradient_max"], skipmode, bs,
            width, height, clip, *cam.translation, *cam.rotation] + [str(d) for d in datasets[1:]]
    p = subprocess.run([str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "Updated gradient map in" in p.stdout

    D, H, W = shape
    opt = VolumeOptions(**tfo)
    tfu = orc.transfer_function_uniform(opt)
    tf = orc.transfer_function_texture(opt)
    ropts = RenderOptions(skipping_type=skipmode, clip_distance=clip, filter=FILTER_HARDWARE)
    ref = rdepth = None
    for k, (r16, endian) in enumerate(raws):
        file_bytes = r16.astype(r16.dtype.newbyteorder(">" if endian == "big" else "<")).view(np.uint8)
        V = orc.normalise(file_bytes, r16.size, "uint16_t", endian, lo, hi).reshape(shape)
        G = orc.gradient_map(V, True)
        O = orc.occupancy_map(V, G, tf, bs, True)
        maps = {1: O, 2: orc.distance_map(O), 3: orc.distance_map_anisotropic(O)}[skipmode]
        dim_b, _ = orc.map_extent((W, H, D), bs)
        if k == 0:        # volume 0's resources are dumped
            info = (out / "info.txt").read_text().split("\n")
            assert [int(x) for x in info[0].split()] == [W, H, D] and [int(x) for x in info[1].split()] == list(dim_b)
            assert int(info[2]) == (8 if skipmode == 3 else 1)
            assert np.array_equal(np.fromfile(out / "voxels.u8", np.uint8).reshape(shape), V)
            assert np.array_equal(np.fromfile(out / "gradient.u8", np.uint8).reshape(shape), G)
            assert np.array_equal(np.fromfile(out / "tf.rgba", np.uint8).reshape(256, 256, 4), tf)
            for i in range(int(info[2])):
                want = maps[i] if skipmode == 3 else maps
                assert np.array_equal(np.fromfile(out / f"map{i}.u8", np.uint8).reshape(want.shape), want), i
            it_file = [float(x) for x in info[4].split()]
            assert np.allclose(it_file, list(orc.parse_header((tmp_path / "a.raw.header").read_text()).image_transform), rtol=1e-6, atol=1e-7)
        it = scene.image_transform(voxel, (W, H, D), axis_angle)
        if k == 1:
            cam.node_translation[:] = [8.0, -3.0, 0.0]
        cu, ru = orc.make_uniforms((W, H, D), dim_b, cam, it, clip)
        ro = RenderOptions(skipping_type=skipmode, clip_distance=clip, filter=FILTER_HARDWARE, load_framebuffer=1 if k else 0)
        ref, rc, _, rdepth = orc.render(V, G, tf, maps, dim_b, cu, ru, tfu, ro, width, height, want_depth=True,
                                        rgba_init=ref if k else None, depth_init=rdepth if k else None)
    img = np.fromfile(out / "frame.rgba", np.uint8).reshape(height, width, 4)
    depth = np.fromfile(out / "depth.f32", np.float32).reshape(height, width)
    d = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=2)
    assert (d <= 1).mean() >= 0.999 and _psnr(img[..., :3], ref[..., :3]) >= 50.0
    assert (np.abs(img[..., 3].astype(int) - ref[..., 3].astype(int)) <= 1).mean() >= 0.999
    same_hit = (depth > 0) == (rdepth > 0)
    assert same_hit.mean() >= 0.999
    assert np.isclose(depth[same_hit], rdepth[same_hit], rtol=1e-4, atol=1e-6).mean() >= 0.995        # hardware filter: the last contributing sample moves on a few rays
    assert (ref[..., :3].max(axis=2) > 0).mean() > 0.02        # something was drawn
