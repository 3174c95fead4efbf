"""In-tree build of libvkv.so (CUDA, sm_100a) and — as test infrastructure — the CPU oracle.

nvcc cross-compiles here without a GPU; the built .so files are git-ignored but travel to
the GPU box with the gpurun snapshot.  Run as ``python -m vkvolume_b200.build``.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "vkvolume_b200" / "csrc"
LIBDIR = ROOT / "vkvolume_b200" / "lib"
OBJDIR = ROOT / "build" / "obj"
LIBVKV = LIBDIR / "libvkv.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # parity: no fused multiply-add contraction, IEEE division and square root,
    # so fp32 results equal the CPU oracle's (-ffp-contract=off) operation for operation
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
]

ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "_build" / "libvkv_oracle.so"


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def _run(cmd, log: Path | None = None):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"command failed: {cmd[0]} ... ({proc.returncode})")
    return proc


def build_libvkv(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h")) + list((ROOT / "vkvolume_b200" / "host").glob("*.h"))
    OBJDIR.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if not force and _newer(obj, [src, *headers, Path(__file__)]):
            return obj
        if verbose:
            print(f"[build] nvcc {src.name}", flush=True)
        _run([NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)], log=OBJDIR / (src.stem + ".ptxas.log"))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or not _newer(LIBVKV, objs):
        if verbose:
            print("[build] link libvkv.so", flush=True)
        _run([NVCC, "-shared", "-o", str(LIBVKV), *map(str, objs), "-cudart", "static",
              "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIBVKV


def build_oracle(force: bool = False, verbose: bool = False) -> Path:
    """Builds the CPU oracle (test infrastructure; never loaded by the product)."""
    src = ORACLE_DIR / "vkv_oracle.c"
    deps = [src, ORACLE_DIR / "vkv_oracle.h", ROOT / "include" / "vkv.h"]
    ORACLE_LIB.parent.mkdir(parents=True, exist_ok=True)
    if force or not _newer(ORACLE_LIB, deps):
        if verbose:
            print("[build] gcc oracle", flush=True)
        _run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-Wall",
              "-o", str(ORACLE_LIB), str(src), "-lm"])
    return ORACLE_LIB


def build_host_tools(force: bool = False, verbose: bool = False):
    """C++ host layer (reference-named classes) + its CLI, linked against libvkv.so."""
    host = ROOT / "vkvolume_b200" / "host"
    out = []
    for name in ("vrender_b200", "host_selftest"):
        src = host / f"{name}.cpp"
        if not src.exists():
            continue
        exe = LIBDIR / name
        deps = [src, *host.glob("*.h"), ROOT / "include" / "vkv.h", LIBVKV]
        if force or not _newer(exe, deps):
            if verbose:
                print(f"[build] g++ {name}", flush=True)
            _run(["g++", "-O2", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-I", str(host), "-I", "/usr/local/cuda/include",
                  str(src), "-o", str(exe), "-L", str(LIBDIR), "-lvkv", "-L", "/usr/local/cuda/lib64", "-lcudart",
                  "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64"])
        out.append(exe)
    return out


def build_all(force: bool = False, verbose: bool = False):
    build_libvkv(force, verbose)
    build_oracle(force, verbose)
    build_host_tools(force, verbose)
    ref = ORACLE_DIR / "ref_shim" / "build_ref.py"
    if ref.exists() and Path("/root/reference").exists():
        _run([sys.executable, str(ref)])


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("ok:", LIBVKV, ORACLE_LIB)
