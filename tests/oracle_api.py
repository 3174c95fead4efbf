"""ctypes binding of the CPU oracle (oracle/_build/libvkv_oracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (vkvolume_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from vkvolume_b200.capi import (CameraDesc, CameraUniform, RayCastUniform, RenderOptions, SampleCounts,  # noqa: E402
                                TransferFunctionUniform, VolumeHeader, VolumeOptions)

ORACLE_LIB = ROOT / "oracle" / "_build" / "libvkv_oracle.so"
_P = C.c_void_p
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = ROOT / "oracle" / "vkv_oracle.c"
        if not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < src.stat().st_mtime:
            ORACLE_LIB.parent.mkdir(parents=True, exist_ok=True)
            subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c11",
                            "-o", str(ORACLE_LIB), str(src), "-lm"], check=True)
        h = C.CDLL(str(ORACLE_LIB))
        h.orc_occupied_voxel_count.restype = C.c_uint64
        h.orc_occupied_voxel_count_dispatch.restype = C.c_uint64
        h.orc_num_threads.restype = C.c_int
        _lib = h
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_P)


def _dims(t):
    return (C.c_uint32 * 3)(*[int(x) for x in t])


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def parse_header(text: str) -> VolumeHeader:
    h = VolumeHeader()
    rc = lib().orc_parse_header(text.encode(), C.byref(h))
    if rc != 0:
        raise ValueError("orc_parse_header failed")
    return h


def normalise(raw: np.ndarray, n_voxels: int, type_name: str, endianness: str, lo: float, hi: float) -> np.ndarray:
    raw = np.ascontiguousarray(raw)
    out = np.empty(n_voxels, dtype=np.uint8)
    rc = lib().orc_normalise(raw.ctypes.data_as(_P), C.c_size_t(n_voxels), type_name.encode(), endianness.encode(),
                             C.c_float(lo), C.c_float(hi), out.ctypes.data_as(_P))
    if rc != 0:
        raise ValueError("unsupported image data type")
    return out


def transfer_function_uniform(opt: VolumeOptions) -> TransferFunctionUniform:
    u = TransferFunctionUniform()
    lib().orc_transfer_function_uniform(C.byref(opt), C.byref(u))
    return u


def transfer_function_texture(opt: VolumeOptions) -> np.ndarray:
    out = np.empty((256, 256, 4), dtype=np.uint8)
    lib().orc_transfer_function_texture(C.byref(opt), out.ctypes.data_as(_P))
    return out


def gradient_map(V: np.ndarray, use_gradient: bool = True, modifier: float = 1.0, want_float: bool = False):
    """V is indexed [z, y, x]."""
    V, pv = _u8(V)
    D, H, W = V.shape
    G = np.empty_like(V)
    Gf = np.empty(V.shape, dtype=np.float32) if want_float else None
    lib().orc_gradient_map(pv, C.c_uint32(W), C.c_uint32(H), C.c_uint32(D), C.c_int(int(use_gradient)), C.c_float(modifier),
                           G.ctypes.data_as(_P), Gf.ctypes.data_as(_P) if want_float else None)
    return (G, Gf) if want_float else G


def map_extent(dim_whd, bs_requested: int):
    dim_b, bs = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
    lib().orc_map_extent(_dims(dim_whd), C.c_uint32(bs_requested), dim_b, bs)
    return tuple(dim_b), tuple(bs)


def occupancy_map(V, G, tf_rgba, bs_requested: int, use_gradient: bool, precomputed: bool = True) -> np.ndarray:
    V, pv = _u8(V)
    D, H, W = V.shape
    if G is None:
        G = np.zeros(1, dtype=np.uint8)
    G, pg = _u8(G)
    tf, pt = _u8(tf_rgba)
    dim_b, _ = map_extent((W, H, D), bs_requested)
    O = np.empty(dim_b[::-1], dtype=np.uint8)
    lib().orc_occupancy_map(pv, pg, pt, _dims((W, H, D)), C.c_uint32(bs_requested), C.c_int(int(use_gradient)),
                            C.c_int(int(precomputed)), O.ctypes.data_as(_P))
    return O


def occupied_voxel_count(V, G, tfu: TransferFunctionUniform, precomputed: bool = True, dispatch_subgroup: int = 0) -> int:
    V, pv = _u8(V)
    D, H, W = V.shape
    if G is None:
        G = np.zeros(1, dtype=np.uint8)
    G, pg = _u8(G)
    if dispatch_subgroup:
        return int(lib().orc_occupied_voxel_count_dispatch(pv, pg, _dims((W, H, D)), C.byref(tfu), C.c_int(int(precomputed)),
                                                           C.c_uint32(dispatch_subgroup)))
    return int(lib().orc_occupied_voxel_count(pv, pg, _dims((W, H, D)), C.byref(tfu), C.c_int(int(precomputed))))


def distance_map(O: np.ndarray) -> np.ndarray:
    O, po = _u8(O)
    Db, Hb, Wb = O.shape
    out = np.empty_like(O)
    lib().orc_distance_map(po, _dims((Wb, Hb, Db)), out.ctypes.data_as(_P))
    return out


def distance_map_anisotropic(O: np.ndarray) -> np.ndarray:
    O, po = _u8(O)
    Db, Hb, Wb = O.shape
    out = np.empty((8, Db, Hb, Wb), dtype=np.uint8)
    lib().orc_distance_map_anisotropic(po, _dims((Wb, Hb, Db)), out.ctypes.data_as(_P))
    return out


def distance_map_closed_form(O: np.ndarray, octant: int = -1) -> np.ndarray:
    O, po = _u8(O)
    Db, Hb, Wb = O.shape
    out = np.empty_like(O)
    lib().orc_distance_map_closed_form(po, _dims((Wb, Hb, Db)), C.c_int(octant), out.ctypes.data_as(_P))
    return out


def make_uniforms(dim_whd, dim_b_whd, cam: CameraDesc, image_transform, clip_distance: float):
    cu, ru = CameraUniform(), RayCastUniform()
    it = (C.c_float * 16)(*[float(x) for x in image_transform])
    lib().orc_make_uniforms(_dims(dim_whd), _dims(dim_b_whd), C.byref(cam), it, C.c_float(clip_distance), C.byref(cu), C.byref(ru))
    return cu, ru


def render(V, G, tf_rgba, maps, dim_b_whd, cu, ru, tfu, opt: RenderOptions, width: int, height: int,
           precomputed: bool = True, y_first: int = 0, y_count: int = -1, want_float=False, want_depth=False,
           rgba_init=None, depth_init=None):
    """Returns (rgba8 [H,W,4], counts, rgba_float or None, depth or None).  With opt.load_framebuffer the frame is blended and
    depth-tested over `rgba_init` / `depth_init` (the attachments' previous contents)."""
    V, pv = _u8(V)
    D, H, W = V.shape
    if G is None:
        G = np.zeros(1, dtype=np.uint8)
    G, pg = _u8(G)
    tf, pt = _u8(tf_rgba)
    if maps is None:
        maps = np.zeros(1, dtype=np.uint8)
    maps, pm = _u8(maps)
    rgba = np.zeros((height, width, 4), dtype=np.uint8) if rgba_init is None else np.ascontiguousarray(rgba_init, np.uint8).copy()
    rf = np.zeros((height, width, 4), dtype=np.float32) if want_float else None
    if depth_init is not None:
        want_depth = True
    dp = (np.zeros((height, width), dtype=np.float32) if depth_init is None else np.ascontiguousarray(depth_init, np.float32).copy()) if want_depth else None
    counts = SampleCounts()
    lib().orc_render(pv, pg, pt, pm, _dims((W, H, D)), _dims(dim_b_whd), C.byref(cu), C.byref(ru), C.byref(tfu), C.byref(opt),
                     C.c_int(int(precomputed)), C.c_int(width), C.c_int(height), C.c_int(y_first), C.c_int(y_count),
                     rgba.ctypes.data_as(_P), rf.ctypes.data_as(_P) if want_float else None,
                     dp.ctypes.data_as(_P) if want_depth else None, C.byref(counts))
    return rgba, counts, rf, dp


def synth_volume(kind: int, seed: int, width: int, height: int, depth: int) -> np.ndarray:
    """CPU twin of capi.synth_volume (synthetic bench inputs; not reference behaviour)."""
    out = np.empty((depth, height, width), dtype=np.uint8)
    rc = lib().orc_synth_volume(C.c_int(kind), C.c_uint64(seed), C.c_uint32(width), C.c_uint32(height), C.c_uint32(depth),
                                out.ctypes.data_as(_P))
    if rc != 0:
        raise ValueError("bad synthetic volume kind")
    return out
