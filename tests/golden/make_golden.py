#!/usr/bin/env python
"""Generates tests/golden/reference_outputs.npz by running the REFERENCE'S OWN SOURCES (GLSL shaders through
oracle/ref_shim, load_volume.cpp as is, glm host maths) on the seeded cases of tests/ref_cases.py.

Run in the build container (needs /root/reference to have built oracle/_ref/libvkv_ref.so):
    python oracle/ref_shim/build_ref.py && python tests/golden/make_golden.py
The .npz is committed; tests/test_oracle_vs_reference.py checks the oracle against it everywhere (including the
GPU box, where /root/reference does not exist).
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import oracle_api as orc        # noqa: E402  (only to build inputs: TF textures / uniforms the shaders consume)
import ref_api as ref           # noqa: E402
import ref_cases as cases       # noqa: E402
from vkvolume_b200.capi import VolumeOptions        # noqa: E402


def main():
    out = {}
    # K1 / K2a / K2b on small volumes, per TF
    for vname in cases.VOLUME_CASES:
        V, bs = cases.volume(vname)
        D, H, W = V.shape
        dim_b, _ = orc.map_extent((W, H, D), bs)
        for tname, o in cases.TF_SETS.items():
            opt = VolumeOptions(**o)
            tfu = orc.transfer_function_uniform(opt)
            tf = orc.transfer_function_texture(opt)
            G = ref.gradient_map(V, tfu)
            out[f"grad/{vname}/{tname}"] = G
            out[f"occ/{vname}/{tname}"] = ref.occupancy_map(V, G, tf, dim_b, tfu, precomputed=True)
            out[f"occ_otf/{vname}/{tname}"] = ref.occupancy_map(V, None, tf, dim_b, tfu, precomputed=False)
            for sg in (8, 32, 64):
                c0, partial = ref.occupied_voxel_count(V, G, tfu, sg, precomputed=True)
                out[f"count/{vname}/{tname}/s{sg}"] = np.array([c0, partial], np.uint64)
            c0, _ = ref.occupied_voxel_count(V, None, tfu, 32, precomputed=False)
            out[f"count_otf/{vname}/{tname}"] = np.array([c0], np.uint64)
    # K3
    for dname in cases.DIST_CASES:
        O = cases.occupancy_grid(dname)
        out[f"dist/{dname}"] = ref.distance_map(O)
        out[f"dist8/{dname}"] = ref.distance_map_anisotropic(O)
    # host maths (glm) + vertex shaders + fragment shader variants
    for inside in (False, True):
        s = cases.render_scene("default", inside)
        D, H, W = cases.RENDER_SHAPE
        tag = "inside" if inside else "outside"
        u = ref.make_uniforms(s["cam"], s["it"], s["clip"], (W, H, D), s["dim_b"])
        for k, v in u.items():
            out[f"uniforms/{tag}/{k}"] = np.asarray(v)
        out[f"vert_clipped/{tag}"] = ref.vertices(s["cu"], s["ru"], "clipped")
        out[f"vert_plane/{tag}"] = ref.vertices(s["cu"], s["ru"], "plane")
    for tname in ("default", "beetle_nograd", "snake_window"):
        s = cases.render_scene(tname, False)
        entries, _ = cases.ray_entries(s)
        out[f"frag_entries/{tname}"] = entries
        maps = {0: None, 1: s["O"], 2: s["Dm"], 3: s["D8"]}
        for skip in (0, 1, 2, 3):
            for ert in (0, 1):
                for test in (0, 3):
                    col, dep = ref.fragments(s["V"], s["G"], s["tf"], maps[skip], s["dim_b"], s["cu"], s["ru"], s["tfu"], entries, skip, bool(ert), test)
                    out[f"frag/{tname}/s{skip}_e{ert}_t{test}"] = col
                    out[f"frag_depth/{tname}/s{skip}_e{ert}_t{test}"] = dep
        for test in (1, 2):
            col, _ = ref.fragments(s["V"], s["G"], s["tf"], s["Dm"], s["dim_b"], s["cu"], s["ru"], s["tfu"], entries, 2, True, test)
            out[f"frag/{tname}/s2_e1_t{test}"] = col
        # DEPTH_ATTACHMENT variants against a synthetic depth attachment (far / just inside / deeper inside / in front)
        depth, d_entries, d_pos, d_cov = cases.depth_attachment_pattern(s)
        out[f"frag_d1/{tname}/depth_in"] = depth
        out[f"frag_d1/{tname}/position"] = d_pos
        for skip, test in ((0, 0), (1, 0), (2, 0), (3, 0), (2, 2)):
            col, dep, disc = ref.fragments(s["V"], s["G"], s["tf"], maps[skip], s["dim_b"], s["cu"], s["ru"], s["tfu"], d_entries, skip, True, test,
                                           positions=d_pos, depth_in=depth[d_cov])
            out[f"frag_d1/{tname}/s{skip}_t{test}"] = col
            out[f"frag_d1_depth/{tname}/s{skip}_t{test}"] = dep
            out[f"frag_d1_discard/{tname}/s{skip}_t{test}"] = disc
        for skip in (0, 2):        # on-the-fly gradient variant (--gradient_test)
            col, _ = ref.fragments(s["V"], None, s["tf"], maps[skip], s["dim_b"], s["cu"], s["ru"], s["tfu"], entries, skip, True, 0, precomputed=False)
            out[f"frag_otf/{tname}/s{skip}"] = col
    # loader: the reference's load_volume.cpp on small raw files
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as td:
        for tname, dt in (("uint8_t", np.uint8), ("int8_t", np.int8), ("uint16_t", np.uint16), ("int16_t", np.int16)):
            for endian in ("little", "big"):
                W, H, D = 11, 7, 5
                info = np.iinfo(dt)
                v = rng.integers(info.min, info.max + 1, size=W * H * D).astype(dt)
                raw = v.astype(v.dtype.newbyteorder(">" if endian == "big" else "<"))
                fn = Path(td) / f"{tname}_{endian}.raw"
                raw.tofile(fn)
                lo, hi = (400.0, 2538.0) if dt in (np.uint16, np.int16) else (10.0, 200.0)
                hdr = f"{W} {H} {D} # extents\n0.004 0.004 0.008 # voxel size\n{lo} {hi} # range\n{tname} {endian} # type\n0 1 0 30 # rotation\n"
                (Path(td) / (fn.name + ".header")).write_text(hdr)
                h = ref.load_header(str(fn) + ".header")
                out[f"loader/{tname}_{endian}/raw"] = raw.view(np.uint8)
                out[f"loader/{tname}_{endian}/u8"] = ref.load_data(str(fn) + ".header", str(fn), W * H * D)
                out[f"loader/{tname}_{endian}/image_transform"] = np.array(h["image_transform"], np.float32)
                out[f"loader/{tname}_{endian}/header_text"] = np.frombuffer(hdr.encode(), np.uint8)
    np.savez_compressed(HERE / "reference_outputs.npz", **out)
    print(f"wrote {HERE / 'reference_outputs.npz'}: {len(out)} arrays, {sum(v.nbytes for v in out.values()) / 1e6:.2f} MB raw")


if __name__ == "__main__":
    main()
