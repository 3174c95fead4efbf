"""CPU-only: libvkv.so loads and exports every entry point include/vkv.h declares (no compute is called)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "vkv.h"
LIB = ROOT / "vkvolume_b200" / "lib" / "libvkv.so"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"VKV_API\s+[\w\s\*]+?\b(vkv_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 40
    for must in ("vkv_context_create", "vkv_volume_create", "vkv_compute_gradient_map", "vkv_compute_occupied_voxel_count",
                 "vkv_compute_distance_map", "vkv_update_transfer_function", "vkv_render", "vkv_render_tiles", "vkv_render_to_host",
                 "vkv_load_header", "vkv_load_data", "vkv_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert LIB.exists(), "libvkv.so is not built (python -m vkvolume_b200.build)"
    lib = ctypes.CDLL(str(LIB))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/vkv.h but not exported: {missing}"


def test_no_cpu_fallback_without_a_device():
    """Without a usable CUDA device the context fails loudly with VKV_ERR_CUDA / VKV_ERR_ARGUMENT — nothing computes on the CPU."""
    lib = ctypes.CDLL(str(LIB))
    lib.vkv_last_error.restype = ctypes.c_char_p
    ctx = ctypes.c_void_p()
    rc = lib.vkv_context_create(0, ctypes.byref(ctx))
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert rc != 0 and not ctx.value
    assert lib.vkv_last_error()


def test_product_does_not_reference_the_oracle():
    """Nothing under vkvolume_b200/ (the product) may import, link or execute anything under oracle/."""
    bad = []
    for path in (ROOT / "vkvolume_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and path.name != "build.py":
            text = path.read_text(errors="ignore")
            if re.search(r"oracle_api|ref_api|libvkv_oracle|libvkv_ref|#include\s+\"[^\"]*oracle", text):
                bad.append(str(path.relative_to(ROOT)))
    assert not bad, bad
