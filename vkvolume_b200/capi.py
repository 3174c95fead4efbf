"""ctypes binding of libvkv.so (include/vkv.h) — the C ABI is the product boundary.

The library is loaded from vkvolume_b200/lib/ (built in-tree by vkvolume_b200.build).  There
is no fallback of any kind: if the shared library is missing, or a call fails, this module
raises.  Nothing here imports or calls the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvkv.so"

SKIP_NONE, SKIP_BLOCK, SKIP_DISTANCE, SKIP_ANISOTROPIC_DISTANCE = 0, 1, 2, 3
TEST_NONE, TEST_RAY_ENTRY, TEST_RAY_EXIT, TEST_NUM_TEXTURE_SAMPLES = 0, 1, 2, 3
FILTER_HARDWARE, FILTER_EXACT = 0, 1
IPC_HANDLE_BYTES = 72
GROUP_HANDLE_BYTES = 3 * IPC_HANDLE_BYTES


class VkvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libvkv error {code}: {message}")
        self.code = code


class TransferFunctionUniform(C.Structure):
    """TransferFunctionUniform (src/transfer_function.h:20-32)."""
    _fields_ = [("sampling_factor", C.c_float), ("voxel_alpha_factor", C.c_float),
                ("grad_magnitude_modifier", C.c_float), ("use_gradient", C.c_uint32),
                ("intensity_min", C.c_float), ("intensity_range_inv", C.c_float),
                ("gradient_min", C.c_float), ("gradient_range_inv", C.c_float)]


class CameraUniform(C.Structure):
    """CameraUniform (src/volume_render_subpass.h:32-39)."""
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("view_proj_inv", C.c_float * 16),
                ("model", C.c_float * 16), ("model_inv", C.c_float * 16)]


class RayCastUniform(C.Structure):
    """RayCastUniform (src/volume_render_subpass.h:46-53)."""
    _fields_ = [("plane", C.c_float * 4), ("plane_tex", C.c_float * 4), ("cam_pos_tex", C.c_float * 4),
                ("block_size", C.c_float * 4), ("front_index", C.c_int32), ("_pad", C.c_int32 * 3)]


class VolumeOptions(C.Structure):
    """Volume::Options (src/volume_component.h:45-56) with the reference's defaults."""
    _fields_ = [("sampling_factor", C.c_float), ("voxel_alpha_factor", C.c_float),
                ("use_precomputed_gradient", C.c_int32), ("intensity_min", C.c_float),
                ("intensity_max", C.c_float), ("gradient_min", C.c_float), ("gradient_max", C.c_float)]

    def __init__(self, sampling_factor=1.0, voxel_alpha_factor=1.0, use_precomputed_gradient=1,
                 intensity_min=0.0, intensity_max=1.0, gradient_min=0.0, gradient_max=1.0):
        super().__init__(sampling_factor, voxel_alpha_factor, use_precomputed_gradient, intensity_min,
                         intensity_max, gradient_min, gradient_max)


class RenderOptions(C.Structure):
    """VolumeRenderSubpass::Options (src/volume_render_subpass.h:74-81) + filter selector."""
    _fields_ = [("skipping_type", C.c_int32), ("clip_distance", C.c_float), ("early_ray_termination", C.c_int32),
                ("depth_attachment", C.c_int32), ("test", C.c_int32), ("filter", C.c_int32), ("load_framebuffer", C.c_int32)]

    def __init__(self, skipping_type=SKIP_DISTANCE, clip_distance=50.0, early_ray_termination=1,
                 depth_attachment=0, test=TEST_NONE, filter=FILTER_HARDWARE, load_framebuffer=0):
        super().__init__(skipping_type, clip_distance, early_ray_termination, depth_attachment, test, filter, load_framebuffer)


class CameraDesc(C.Structure):
    """Camera + volume-node description; defaults are the reference's Sponza main_camera and node scale 100."""
    _fields_ = [("translation", C.c_float * 3), ("rotation", C.c_float * 4), ("yfov", C.c_float),
                ("aspect", C.c_float), ("znear", C.c_float), ("zfar", C.c_float),
                ("node_translation", C.c_float * 3), ("node_rotation", C.c_float * 4), ("node_scale", C.c_float * 3)]

    def __init__(self, translation=(-705.01, 195.20, -119.93), rotation=(-0.004728, -0.775409, -0.005807, 0.631416),
                 yfov=1.0, aspect=1.0, znear=1.0, zfar=4000.0, node_translation=(0, 0, 0),
                 node_rotation=(0, 0, 0, 1), node_scale=(100, 100, 100)):
        super().__init__((C.c_float * 3)(*translation), (C.c_float * 4)(*rotation), yfov, aspect, znear, zfar,
                         (C.c_float * 3)(*node_translation), (C.c_float * 4)(*node_rotation),
                         (C.c_float * 3)(*node_scale))


class VolumeHeader(C.Structure):
    """LoadVolume::Header (src/load_volume.h:29-39)."""
    _fields_ = [("extent", C.c_uint32 * 3), ("voxel_size", C.c_float * 3), ("normalisation_range", C.c_float * 2),
                ("type", C.c_char * 16), ("endianness", C.c_char * 16), ("image_transform", C.c_float * 16)]


class SampleCounts(C.Structure):
    _fields_ = [("volume_samples", C.c_uint64), ("distance_samples", C.c_uint64),
                ("empty_samples", C.c_uint64), ("covered_pixels", C.c_uint64)]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes) — exactly the symbols include/vkv.h declares
    "vkv_last_error": (C.c_char_p, []),
    "vkv_version": (C.c_char_p, []),
    "vkv_context_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vkv_context_destroy": (None, [_P]),
    "vkv_context_device": (C.c_int, [_P]),
    "vkv_context_sm_count": (C.c_int, [_P]),
    "vkv_stream_synchronize": (C.c_int, [_P, _P]),
    "vkv_load_header": (C.c_int, [C.c_char_p, C.POINTER(VolumeHeader)]),
    "vkv_load_data": (C.c_int, [C.c_char_p, C.POINTER(VolumeHeader), _P, C.c_size_t]),
    "vkv_volume_create": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(_P)]),
    "vkv_volume_destroy": (None, [_P]),
    "vkv_volume_upload": (C.c_int, [_P, _P, _P]),
    "vkv_volume_upload_device": (C.c_int, [_P, _P, _P]),
    "vkv_volume_upload_raw": (C.c_int, [_P, _P, C.c_size_t, C.c_char_p, C.c_char_p, C.c_float, C.c_float, _P]),
    "vkv_volume_set_number_of_distance_maps": (C.c_int, [_P, C.c_size_t]),
    "vkv_transfer_function_uniform_from_options": (C.c_int, [C.POINTER(VolumeOptions), C.POINTER(TransferFunctionUniform)]),
    "vkv_volume_update_transfer_function_texture": (C.c_int, [_P, C.POINTER(VolumeOptions), _P]),
    "vkv_volume_set_transfer_function_texture": (C.c_int, [_P, _P, _P]),
    "vkv_compute_gradient_map": (C.c_int, [_P, C.POINTER(TransferFunctionUniform), _P]),
    "vkv_compute_occupied_voxel_count": (C.c_int, [_P, C.POINTER(TransferFunctionUniform), C.POINTER(C.c_uint64), _P]),
    "vkv_compute_distance_map": (C.c_int, [_P, C.POINTER(TransferFunctionUniform), C.c_int, _P]),
    "vkv_update_transfer_function": (C.c_int, [_P, C.POINTER(VolumeOptions), C.c_int, C.POINTER(C.c_uint64), _P]),
    "vkv_make_uniforms": (C.c_int, [_P, C.POINTER(CameraDesc), C.POINTER(C.c_float), C.c_float,
                                    C.POINTER(CameraUniform), C.POINTER(RayCastUniform)]),
    "vkv_make_uniforms_for_extent": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(CameraDesc), C.POINTER(C.c_float),
                                               C.c_float, C.POINTER(CameraUniform), C.POINTER(RayCastUniform)]),
    "vkv_render": (C.c_int, [_P, C.POINTER(CameraUniform), C.POINTER(RayCastUniform), C.POINTER(TransferFunctionUniform),
                             C.POINTER(RenderOptions), C.c_int, C.c_int, _P, _P, _P, _P]),
    "vkv_render_tiles": (C.c_int, [_P, C.POINTER(CameraUniform), C.POINTER(RayCastUniform), C.POINTER(TransferFunctionUniform),
                                   C.POINTER(RenderOptions), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   _P, _P, _P, _P]),
    "vkv_render_to_host": (C.c_int, [_P, C.POINTER(CameraUniform), C.POINTER(RayCastUniform), C.POINTER(TransferFunctionUniform),
                                     C.POINTER(RenderOptions), C.c_int, C.c_int, _P, C.POINTER(SampleCounts), _P]),
    "vkv_render_to_host_async": (C.c_int, [_P, C.POINTER(CameraUniform), C.POINTER(RayCastUniform), C.POINTER(TransferFunctionUniform),
                                           C.POINTER(RenderOptions), C.c_int, C.c_int, _P, _P, _P]),
    "vkv_render_to_host_wait": (C.c_int, [_P, _P]),
    "vkv_render_over_host": (C.c_int, [_P, C.POINTER(CameraUniform), C.POINTER(RayCastUniform), C.POINTER(TransferFunctionUniform),
                                       C.POINTER(RenderOptions), C.c_int, C.c_int, _P, _P, C.POINTER(SampleCounts), _P]),
    "vkv_volume_extent": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "vkv_volume_map_extent": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "vkv_volume_block_size": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "vkv_volume_number_of_distance_maps": (C.c_size_t, [_P]),
    "vkv_volume_device_voxels": (_P, [_P]),
    "vkv_volume_device_gradient": (_P, [_P]),
    "vkv_volume_device_distance_map": (_P, [_P, C.c_size_t]),
    "vkv_volume_device_transfer_function": (_P, [_P]),
    "vkv_volume_download_voxels": (C.c_int, [_P, _P, C.c_size_t]),
    "vkv_volume_download_gradient": (C.c_int, [_P, _P, C.c_size_t]),
    "vkv_volume_download_gradient_texture": (C.c_int, [_P, _P, C.c_size_t]),
    "vkv_volume_download_distance_map": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t]),
    "vkv_volume_download_transfer_function": (C.c_int, [_P, _P, C.c_size_t]),
    "vkv_volume_upload_gradient": (C.c_int, [_P, _P, _P]),
    "vkv_compute_occupancy_slab": (C.c_int, [_P, C.POINTER(TransferFunctionUniform), C.c_int, C.c_uint32, C.c_uint32, _P, _P]),
    "vkv_compute_distance_from_occupancy": (C.c_int, [_P, C.c_int, _P]),
    "vkv_volume_mark_occupancy_present": (C.c_int, [_P, C.c_int]),
    "vkv_volume_group_export": (C.c_int, [_P, _P]),
    "vkv_volume_group_open": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "vkv_volume_group_close": (C.c_int, [_P]),
    "vkv_update_transfer_function_sharded": (C.c_int, [_P, C.POINTER(VolumeOptions), C.c_int, C.POINTER(C.c_uint64), _P]),
    "vkv_ipc_export": (C.c_int, [_P, _P]),
    "vkv_ipc_open": (C.c_int, [_P, C.POINTER(_P)]),
    "vkv_ipc_close": (C.c_int, [_P]),
    "vkv_synth_volume": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P]),
    "vkv_bench_tex3d": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "vkv_kernel_launch_count": (C.c_uint64, []),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Loads libvkv.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -m vkvolume_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        # VKV_LIB: debug knob, an alternative build of the same library (A/B of compile-time kernel parameters)
        handle = C.CDLL(os.environ.get("VKV_LIB") or str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)        # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise VkvError(rc, lib().vkv_last_error().decode("utf-8", "replace"))


def _host_ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


class Context:
    def __init__(self, device: int = 0):
        self.handle = _P()
        check(lib().vkv_context_create(device, C.byref(self.handle)))
        self.device = device

    @property
    def sm_count(self) -> int:
        return lib().vkv_context_sm_count(self.handle)

    def synchronize(self, stream: int = 0):
        check(lib().vkv_stream_synchronize(self.handle, _P(stream)))

    def close(self):
        if self.handle:
            lib().vkv_context_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Volume:
    """Thin OO veneer over the vkv_volume_* / vkv_compute_* / vkv_render* entry points.

    Method names follow the reference's classes (Volume, ComputeGradientMap::compute,
    ComputeOccupiedVoxelCount::compute, ComputeDistanceMap::compute, VolumeRenderSubpass::draw).
    """

    def __init__(self, ctx: Context, width: int, height: int, depth: int, block_size: int = 4,
                 use_precomputed_gradient: bool = True):
        self.ctx = ctx
        self.handle = _P()
        check(lib().vkv_volume_create(ctx.handle, width, height, depth, block_size, int(use_precomputed_gradient),
                                      C.byref(self.handle)))
        self.extent = (width, height, depth)
        e = (C.c_uint32 * 3)()
        check(lib().vkv_volume_map_extent(self.handle, e))
        self.map_extent = tuple(e)
        check(lib().vkv_volume_block_size(self.handle, e))
        self.block_size = tuple(e)

    # -- resources -------------------------------------------------------------------------
    @property
    def n_voxels(self):
        w, h, d = self.extent
        return w * h * d

    @property
    def n_blocks(self):
        w, h, d = self.map_extent
        return w * h * d

    def upload(self, voxels: np.ndarray, stream: int = 0):
        voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
        assert voxels.size == self.n_voxels
        check(lib().vkv_volume_upload(self.handle, _host_ptr(voxels), _P(stream)))

    def upload_device(self, dev_ptr: int, stream: int = 0):
        check(lib().vkv_volume_upload_device(self.handle, _P(dev_ptr), _P(stream)))

    def upload_raw(self, raw: np.ndarray, type_name: str, endianness: str, lo: float, hi: float, stream: int = 0):
        raw = np.ascontiguousarray(raw)
        check(lib().vkv_volume_upload_raw(self.handle, _host_ptr(raw), raw.nbytes, type_name.encode(),
                                          endianness.encode(), lo, hi, _P(stream)))

    def upload_gradient(self, gradient: np.ndarray, stream: int = 0):
        gradient = np.ascontiguousarray(gradient, dtype=np.uint8)
        assert gradient.size == self.n_voxels
        check(lib().vkv_volume_upload_gradient(self.handle, _host_ptr(gradient), _P(stream)))
        self.ctx.synchronize(stream)

    def set_number_of_distance_maps(self, n: int):
        check(lib().vkv_volume_set_number_of_distance_maps(self.handle, n))

    def update_transfer_function_texture(self, options: VolumeOptions, stream: int = 0):
        check(lib().vkv_volume_update_transfer_function_texture(self.handle, C.byref(options), _P(stream)))

    def set_transfer_function_texture(self, rgba: np.ndarray, stream: int = 0):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert rgba.size == 256 * 256 * 4
        check(lib().vkv_volume_set_transfer_function_texture(self.handle, _host_ptr(rgba), _P(stream)))
        self.ctx.synchronize(stream)

    # -- compute components ----------------------------------------------------------------
    def compute_gradient_map(self, tfu: TransferFunctionUniform, stream: int = 0):
        check(lib().vkv_compute_gradient_map(self.handle, C.byref(tfu), _P(stream)))

    def compute_occupied_voxel_count(self, tfu: TransferFunctionUniform, stream: int = 0) -> int:
        out = C.c_uint64(0)
        check(lib().vkv_compute_occupied_voxel_count(self.handle, C.byref(tfu), C.byref(out), _P(stream)))
        return out.value

    def compute_distance_map(self, tfu: TransferFunctionUniform, skipping_type: int, stream: int = 0):
        check(lib().vkv_compute_distance_map(self.handle, C.byref(tfu), skipping_type, _P(stream)))

    def compute_occupancy_slab(self, tfu, skipping_type: int, zb_first: int, zb_count: int, count_dev: int = 0, stream: int = 0):
        check(lib().vkv_compute_occupancy_slab(self.handle, C.byref(tfu), skipping_type, zb_first, zb_count,
                                               _P(count_dev), _P(stream)))

    def compute_distance_from_occupancy(self, skipping_type: int, stream: int = 0):
        check(lib().vkv_compute_distance_from_occupancy(self.handle, skipping_type, _P(stream)))

    def update_transfer_function(self, options: VolumeOptions, skipping_type: int, count: bool = False, stream: int = 0):
        out = C.c_uint64(0)
        check(lib().vkv_update_transfer_function(self.handle, C.byref(options), skipping_type,
                                                 C.byref(out) if count else None, _P(stream)))
        return out.value if count else None

    # -- multi-GPU group (one process per GPU) ---------------------------------------------
    def group_export(self) -> bytes:
        buf = (C.c_uint8 * GROUP_HANDLE_BYTES)()
        check(lib().vkv_volume_group_export(self.handle, buf))
        return bytes(buf)

    def group_open(self, rank: int, world: int, all_handles: list):
        """all_handles: every rank's group_export() blob, in rank order."""
        blob = b"".join(all_handles)
        assert len(blob) == world * GROUP_HANDLE_BYTES
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(lib().vkv_volume_group_open(self.handle, rank, world, buf))

    def group_close(self):
        check(lib().vkv_volume_group_close(self.handle))

    def update_transfer_function_sharded(self, options: VolumeOptions, skipping_type: int, count: bool = False, stream: int = 0):
        """Collective over the volume's group: every rank calls it with the same arguments."""
        out = C.c_uint64(0)
        check(lib().vkv_update_transfer_function_sharded(self.handle, C.byref(options), skipping_type,
                                                         C.byref(out) if count else None, _P(stream)))
        return out.value if count else None

    # -- ray caster ------------------------------------------------------------------------
    def make_uniforms(self, cam: CameraDesc, image_transform, clip_distance: float):
        cu, ru = CameraUniform(), RayCastUniform()
        it = (C.c_float * 16)(*[float(x) for x in image_transform])
        check(lib().vkv_make_uniforms(self.handle, C.byref(cam), it, clip_distance, C.byref(cu), C.byref(ru)))
        return cu, ru

    def render(self, cu, ru, tfu, opt: RenderOptions, width: int, height: int, rgba8_dev: int, depth_dev: int = 0,
               counts_dev: int = 0, stream: int = 0):
        check(lib().vkv_render(self.handle, C.byref(cu), C.byref(ru), C.byref(tfu), C.byref(opt), width, height,
                               _P(rgba8_dev), _P(depth_dev), _P(counts_dev), _P(stream)))

    def render_tiles(self, cu, ru, tfu, opt, width, height, tile_w, tile_h, tile_first, tile_stride, rgba8_dev,
                     depth_dev: int = 0, counts_dev: int = 0, stream: int = 0):
        check(lib().vkv_render_tiles(self.handle, C.byref(cu), C.byref(ru), C.byref(tfu), C.byref(opt), width, height,
                                     tile_w, tile_h, tile_first, tile_stride, _P(rgba8_dev), _P(depth_dev),
                                     _P(counts_dev), _P(stream)))

    def render_to_host(self, cu, ru, tfu, opt, width, height, out: np.ndarray | None = None, want_counts=True, stream: int = 0):
        if out is None:
            out = np.empty((height, width, 4), dtype=np.uint8)
        counts = SampleCounts()
        check(lib().vkv_render_to_host(self.handle, C.byref(cu), C.byref(ru), C.byref(tfu), C.byref(opt), width, height,
                                       _host_ptr(out), C.byref(counts) if want_counts else None, _P(stream)))
        return out, counts

    def render_to_host_async(self, cu, ru, tfu, opt, width, height, rgba_host_ptr: int, counts_host_ptr: int = 0, stream: int = 0):
        """vkv_render_to_host_async: page-locked destination pointers; contents are defined after render_to_host_wait()."""
        check(lib().vkv_render_to_host_async(self.handle, C.byref(cu), C.byref(ru), C.byref(tfu), C.byref(opt), width, height,
                                             _P(rgba_host_ptr), _P(counts_host_ptr), _P(stream)))

    def render_to_host_wait(self, stream: int = 0):
        check(lib().vkv_render_to_host_wait(self.handle, _P(stream)))

    def render_over_host(self, cu, ru, tfu, opt, width, height, rgba: np.ndarray | None = None, depth: np.ndarray | None = None,
                         want_counts=True, stream: int = 0):
        """vkv_render_over_host: composites over `rgba` [H,W,4] u8 / `depth` [H,W] f32 in place when opt.load_framebuffer
        (VolumeRenderSubpass::draw's blend + depth test over what is already in the target); returns (rgba, depth, counts)."""
        if rgba is None:
            rgba = np.empty((height, width, 4), dtype=np.uint8)
        assert rgba.dtype == np.uint8 and rgba.flags.c_contiguous and rgba.shape == (height, width, 4)
        if depth is not None:
            assert depth.dtype == np.float32 and depth.flags.c_contiguous and depth.shape == (height, width)
        counts = SampleCounts()
        check(lib().vkv_render_over_host(self.handle, C.byref(cu), C.byref(ru), C.byref(tfu), C.byref(opt), width, height,
                                         _host_ptr(rgba), _host_ptr(depth) if depth is not None else None,
                                         C.byref(counts) if want_counts else None, _P(stream)))
        return rgba, depth, counts

    # -- read-backs ------------------------------------------------------------------------
    def download_voxels(self):
        out = np.empty(self.extent[::-1], dtype=np.uint8)
        check(lib().vkv_volume_download_voxels(self.handle, _host_ptr(out), out.nbytes))
        return out

    def download_gradient(self):
        out = np.empty(self.extent[::-1], dtype=np.uint8)
        check(lib().vkv_volume_download_gradient(self.handle, _host_ptr(out), out.nbytes))
        return out

    def download_gradient_texture(self):
        out = np.empty(self.extent[::-1], dtype=np.uint8)
        check(lib().vkv_volume_download_gradient_texture(self.handle, _host_ptr(out), out.nbytes))
        return out

    def download_distance_map(self, idx: int = 0):
        out = np.empty(self.map_extent[::-1], dtype=np.uint8)
        check(lib().vkv_volume_download_distance_map(self.handle, idx, _host_ptr(out), out.nbytes))
        return out

    def download_transfer_function(self):
        out = np.empty((256, 256, 4), dtype=np.uint8)
        check(lib().vkv_volume_download_transfer_function(self.handle, _host_ptr(out), out.nbytes))
        return out

    def device_voxels(self) -> int:
        return lib().vkv_volume_device_voxels(self.handle) or 0

    def device_gradient(self) -> int:
        return lib().vkv_volume_device_gradient(self.handle) or 0

    def device_distance_map(self, idx: int = 0) -> int:
        return lib().vkv_volume_device_distance_map(self.handle, idx) or 0

    def close(self):
        if self.handle:
            lib().vkv_volume_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def transfer_function_uniform(options: VolumeOptions) -> TransferFunctionUniform:
    u = TransferFunctionUniform()
    check(lib().vkv_transfer_function_uniform_from_options(C.byref(options), C.byref(u)))
    return u


def make_uniforms_for_extent(extent, map_extent, cam: CameraDesc, image_transform, clip_distance: float):
    """vkv_make_uniforms_for_extent: the host maths of VolumeRenderSubpass::draw without a device or a volume handle."""
    cu, ru = CameraUniform(), RayCastUniform()
    it = (C.c_float * 16)(*[float(x) for x in image_transform])
    check(lib().vkv_make_uniforms_for_extent((C.c_uint32 * 3)(*extent), (C.c_uint32 * 3)(*map_extent), C.byref(cam), it,
                                             clip_distance, C.byref(cu), C.byref(ru)))
    return cu, ru


def load_header(path: str) -> VolumeHeader:
    h = VolumeHeader()
    check(lib().vkv_load_header(path.encode(), C.byref(h)))
    return h


def load_data(path: str, header: VolumeHeader) -> np.ndarray:
    w, h, d = header.extent
    out = np.empty((d, h, w), dtype=np.uint8)
    check(lib().vkv_load_data(path.encode(), C.byref(header), _host_ptr(out), out.nbytes))
    return out


def synth_volume(ctx: Context, kind: int, seed: int, width: int, height: int, depth: int, dev_ptr: int, stream: int = 0):
    check(lib().vkv_synth_volume(ctx.handle, kind, seed, width, height, depth, _P(dev_ptr), _P(stream)))


def bench_tex3d(ctx: Context, extent: int, fetches_per_thread: int = 256, coherent: bool = True) -> float:
    out = C.c_double(0.0)
    check(lib().vkv_bench_tex3d(ctx.handle, extent, fetches_per_thread, int(coherent), C.byref(out)))
    return out.value


def kernel_launch_count() -> int:
    return lib().vkv_kernel_launch_count()
