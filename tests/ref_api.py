"""ctypes binding of oracle/_ref/libvkv_ref.so — the reference's own shader sources (and load_volume.cpp, and
glm host maths) executed on the CPU through oracle/ref_shim.  TEST INFRASTRUCTURE; present only where
/root/reference was available at build time (the .so travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_LIB = ROOT / "oracle" / "_ref" / "libvkv_ref.so"
_P = C.c_void_p


class RefArgs(C.Structure):
    """Mirror of oracle/ref_shim/ref_args.h."""
    _fields_ = [
        ("V", _P), ("G", _P), ("W", C.c_int32), ("H", C.c_int32), ("D", C.c_int32), ("tf_rgba", _P),
        ("sampling_factor", C.c_float), ("voxel_alpha_factor", C.c_float), ("grad_magnitude_modifier", C.c_float),
        ("use_gradient", C.c_int32), ("intensity_min", C.c_float), ("intensity_range_inv", C.c_float),
        ("gradient_min", C.c_float), ("gradient_range_inv", C.c_float),
        ("maps", _P * 8), ("swap", _P), ("Wb", C.c_int32), ("Hb", C.c_int32), ("Db", C.c_int32), ("block_size", C.c_int32 * 3),
        ("count", _P), ("count_elements", C.c_uint64), ("subgroup_size", C.c_uint32),
        ("view", C.c_float * 16), ("proj", C.c_float * 16), ("view_proj_inv", C.c_float * 16), ("model", C.c_float * 16),
        ("model_inv", C.c_float * 16), ("plane", C.c_float * 4), ("plane_tex", C.c_float * 4), ("cam_pos_tex", C.c_float * 4),
        ("block_size_f", C.c_float * 4), ("front_index", C.c_int32),
        ("n_frag", C.c_int32), ("frag_entry", _P), ("frag_out", _P), ("frag_depth", _P),
        ("vert_out", C.c_float * 64),
        ("frag_position", _P), ("frag_depth_in", _P), ("frag_discarded", _P),
    ]


_lib = None


def available() -> bool:
    return REF_LIB.exists()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(REF_LIB))
        _lib.ref_loader_error.restype = C.c_char_p
    return _lib


def _call(variant: str, a: RefArgs):
    fn = getattr(lib(), "ref_" + variant)
    fn.restype = None
    fn(C.byref(a))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_P)


def _set_tfu(a: RefArgs, tfu):
    for k in ("sampling_factor", "voxel_alpha_factor", "grad_magnitude_modifier", "intensity_min", "intensity_range_inv",
              "gradient_min", "gradient_range_inv"):
        setattr(a, k, getattr(tfu, k))
    a.use_gradient = int(tfu.use_gradient)


def _set_volume(a: RefArgs, V, G=None):
    D, H, W = V.shape
    a.V, a.W, a.H, a.D = _ptr(V), W, H, D
    if G is not None:
        a.G = _ptr(G)


def gradient_map(V: np.ndarray, tfu) -> np.ndarray:
    V = np.ascontiguousarray(V, np.uint8)
    G = np.zeros_like(V)
    a = RefArgs()
    _set_volume(a, V, G)
    _set_tfu(a, tfu)
    _call("gradient_map", a)
    return G


def occupancy_map(V, G, tf, dim_b_whd, tfu, precomputed=True) -> np.ndarray:
    V = np.ascontiguousarray(V, np.uint8)
    G = np.ascontiguousarray(G if G is not None else np.zeros_like(V), np.uint8)
    tf = np.ascontiguousarray(tf, np.uint8)
    Wb, Hb, Db = dim_b_whd
    O = np.full((Db, Hb, Wb), 77, np.uint8)
    a = RefArgs()
    _set_volume(a, V, G)
    _set_tfu(a, tfu)
    a.tf_rgba = _ptr(tf)
    a.maps[0] = _ptr(O).value
    a.Wb, a.Hb, a.Db = Wb, Hb, Db
    _call(f"occupancy_map_p{int(precomputed)}", a)
    return O


def occupied_voxel_count(V, G, tfu, subgroup_size=32, precomputed=True) -> int:
    """Both dispatches: per-subgroup partial sums, then the strided reduce loop; returns count[0] like get_result()."""
    V = np.ascontiguousarray(V, np.uint8)
    G = np.ascontiguousarray(G if G is not None else np.zeros_like(V), np.uint8)
    D, H, W = V.shape
    n = ((W + 7) // 8) * ((H + 7) // 8) * ((D + 7) // 8) * (512 // subgroup_size)        # initialise_buffer (:67-78)
    count = np.zeros(n, np.uint64)
    a = RefArgs()
    _set_volume(a, V, G)
    _set_tfu(a, tfu)
    a.count, a.count_elements, a.subgroup_size = _ptr(count), n, subgroup_size
    _call(f"occupied_voxel_count_p{int(precomputed)}", a)
    partial_total = int(count.sum())
    _call(f"occupied_voxel_count_reduce_s{subgroup_size}", a)
    return int(count[0]), partial_total


def distance_map(O: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(O, np.uint8).copy()
    swap = np.zeros_like(m)
    Db, Hb, Wb = m.shape
    a = RefArgs()
    a.maps[0], a.swap = _ptr(m).value, _ptr(swap)
    a.Wb, a.Hb, a.Db = Wb, Hb, Db
    _call("distance_map", a)
    return m


def distance_map_anisotropic(O: np.ndarray) -> np.ndarray:
    Db, Hb, Wb = O.shape
    maps = np.zeros((8, Db, Hb, Wb), np.uint8)
    maps[7] = O
    swap = np.zeros((Db, Hb, Wb), np.uint8)
    a = RefArgs()
    for i in range(8):
        a.maps[i] = maps[i].ctypes.data_as(_P).value
    a.swap = _ptr(swap)
    a.Wb, a.Hb, a.Db = Wb, Hb, Db
    _call("distance_map_anisotropic", a)
    return maps


def _set_camera(a: RefArgs, cu, ru):
    for k in ("view", "proj", "view_proj_inv", "model", "model_inv"):
        setattr(a, k, getattr(cu, k))
    a.plane, a.plane_tex, a.cam_pos_tex, a.block_size_f, a.front_index = ru.plane, ru.plane_tex, ru.cam_pos_tex, ru.block_size, ru.front_index


def fragments(V, G, tf, maps, dim_b_whd, cu, ru, tfu, entries: np.ndarray, skip: int, ert: bool, test: int = 0, precomputed=True,
              positions: np.ndarray | None = None, depth_in: np.ndarray | None = None):
    """Runs volume_render.frag (the selected #define variant) for each ray_entry in `entries` [n,3].
    With `positions` [n,4] and `depth_in` [n] the DEPTH_ATTACHMENT variant runs and a third array (1 = discarded) is returned."""
    V = np.ascontiguousarray(V, np.uint8)
    G = np.ascontiguousarray(G if G is not None else np.zeros_like(V), np.uint8)
    tf = np.ascontiguousarray(tf, np.uint8)
    entries = np.ascontiguousarray(entries, np.float32)
    n = entries.shape[0]
    out = np.zeros((n, 4), np.float32)
    depth = np.zeros(n, np.float32)
    a = RefArgs()
    _set_volume(a, V, G)
    _set_tfu(a, tfu)
    _set_camera(a, cu, ru)
    a.tf_rgba = _ptr(tf)
    Wb, Hb, Db = dim_b_whd
    a.Wb, a.Hb, a.Db = Wb, Hb, Db
    keep = []
    if maps is not None:
        maps = np.ascontiguousarray(maps, np.uint8)
        if maps.ndim == 4:
            for i in range(8):
                a.maps[i] = maps[i].ctypes.data_as(_P).value
        else:
            a.maps[0] = _ptr(maps).value
        keep.append(maps)
    a.n_frag, a.frag_entry, a.frag_out, a.frag_depth = n, _ptr(entries), _ptr(out), _ptr(depth)
    if depth_in is not None:
        positions = np.ascontiguousarray(positions, np.float32)
        depth_in = np.ascontiguousarray(depth_in, np.float32)
        disc = np.zeros(n, np.int32)
        a.frag_position, a.frag_depth_in, a.frag_discarded = _ptr(positions), _ptr(depth_in), _ptr(disc)
        _call(f"frag_p{int(precomputed)}_s{skip}_e{int(ert)}_t{test}_d1", a)
        return out, depth, disc
    _call(f"frag_p{int(precomputed)}_s{skip}_e{int(ert)}_t{test}", a)
    return out, depth


def vertices(cu, ru, which: str) -> np.ndarray:
    """which = 'clipped' (8 cube vertices) or 'plane' (6 polygon vertices): rows of [gl_Position(4), ray_entry(3), clip distance]."""
    a = RefArgs()
    _set_camera(a, cu, ru)
    _call("vert_" + which, a)
    n = 8 if which == "clipped" else 6
    return np.array(list(a.vert_out), np.float32).reshape(8, 8)[:n]


def make_uniforms(cam, image_transform, clip_distance, dim_whd, dim_b_whd):
    """glm evaluation of volume_render_subpass.cpp:221-249 -> dict of float arrays."""
    vals = list(cam.translation) + list(cam.rotation) + [cam.yfov, cam.aspect, cam.znear, cam.zfar] + list(cam.node_translation) + \
        list(cam.node_rotation) + list(cam.node_scale) + [float(x) for x in image_transform] + [clip_distance] + \
        [float(x) for x in dim_whd] + [float(x) for x in dim_b_whd] + [0.0]
    inp = (C.c_float * len(vals))(*vals)
    out = (C.c_float * 97)()
    lib().ref_make_uniforms(inp, out)
    o = np.array(list(out), np.float32)
    return dict(view=o[0:16], proj=o[16:32], view_proj_inv=o[32:48], model=o[48:64], model_inv=o[64:80], plane=o[80:84],
                plane_tex=o[84:88], cam_pos_tex=o[88:92], block_size=o[92:96], front_index=int(o[96]))


def load_header(path: str):
    f = (C.c_float * 24)()
    t, e = C.create_string_buffer(16), C.create_string_buffer(16)
    if lib().ref_load_header(path.encode(), f, t, e) != 0:
        raise RuntimeError(lib().ref_loader_error().decode())
    f = list(f)
    return dict(extent=tuple(int(x) for x in f[0:3]), voxel_size=f[3:6], normalisation_range=f[6:8], image_transform=f[8:24],
                type=t.value.decode(), endianness=e.value.decode())


def load_data(path_header: str, path_data: str, n_voxels: int) -> np.ndarray:
    out = np.zeros(n_voxels, np.uint8)
    rc = lib().ref_load_data(path_header.encode(), path_data.encode(), _ptr(out), C.c_size_t(n_voxels))
    if rc != 0:
        raise RuntimeError(lib().ref_loader_error().decode())
    return out
