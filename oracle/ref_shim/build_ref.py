#!/usr/bin/env python
"""Builds oracle/_ref/libvkv_ref.so: the REFERENCE'S OWN SOURCES executed on the CPU.  Test infrastructure.

Nothing from /root/reference is copied into the repository.  At build time this script
  * reads the GLSL shaders where they lie (/root/reference/shaders), resolves their #includes, applies a
    purely syntactic GLSL -> C++ transliteration (strip layout()/uniform/in/out/precision qualifiers, turn
    interface blocks into structs or globals, array constructors into brace initialisers, main() ->
    shader_main()) and writes the result under oracle/_ref/gen/ (git-ignored);
  * compiles every shader VARIANT (the #define sets the host code selects, volume_render_subpass.cpp:57-92,
    compute_distance_map.cpp:115-119, compute_occupied_voxel_count.cpp:89-93,49) as its own translation unit
    against oracle/ref_shim/glsl_compat.h, wrapped by the small drivers in ref_harness.h that loop over
    invocations the way the reference's dispatch()/draw calls do;
  * compiles src/load_volume.cpp AS IS (with a 10-line stand-in for boost/endian/conversion.hpp, which is
    absent from the image, plus the vendored glm and vulkan headers);
  * compiles ref_glm.cpp, which evaluates the host maths of volume_render_subpass.cpp:221-249 with the
    reference's own vendored glm.
Two source patches are needed for C++ name lookup and are applied textually (both listed in oracle/README.md):
  1. distance_map_anisotropic.comp:47 `start.x` on an int (GLSL allows swizzling scalars) -> `start`;
  2. volume_render.frag:89 `float gradient = texture(gradient, pos).x;` — in GLSL the new name is not yet in
     scope inside its own initialiser, in C++ it is — the sampler is renamed `gradient_sampler`.
"""
from __future__ import annotations

import itertools
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
SHADERS = REF / "shaders"
OUT = ROOT / "oracle" / "_ref"
GEN = OUT / "gen"
GLM_INC = REF / "third_party" / "Vulkan-samples" / "third_party" / "glm"
VK_INC = REF / "third_party" / "Vulkan-samples" / "third_party" / "vulkan" / "include"


def _match_paren(s: str, i: int) -> int:
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parenthesis")


def transliterate(name: str, seen=None) -> str:
    """GLSL text -> C++ text (syntax only; no arithmetic is touched)."""
    src = (SHADERS / name).read_text()
    # includes (GL_GOOGLE_include_directive)
    src = re.sub(r'^\s*#include\s+"([^"]+)"\s*$', lambda m: transliterate(m.group(1)), src, flags=re.M)
    src = re.sub(r"^\s*#(version|extension)\b.*$", "", src, flags=re.M)
    src = re.sub(r"^\s*precision\s+\w+\s+\w+\s*;", "", src, flags=re.M)
    # layout(...) qualifiers (one level of nested parentheses is enough for these shaders)
    src = re.sub(r"layout\s*\((?:[^()]|\([^()]*\))*\)", "", src)
    src = re.sub(r"^\s*in\s*;", "", src, flags=re.M)        # what is left of `layout(local_size...) in;`

    # interface blocks
    def block(m):
        kind, tname, body, inst = m.group(1), m.group(2), m.group(3), m.group(4)
        body = re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1 *\2;", body)        # runtime-sized array -> pointer
        if inst:
            return f"struct {tname} {{{body}}} {inst};"
        return body        # members become globals
    src = re.sub(r"\b(uniform|buffer|out|in)\s+(\w+)\s*\{([^{}]*)\}\s*(\w*)\s*;", block, src, flags=re.S)
    # remaining storage / precision qualifiers on global declarations and parameters
    src = re.sub(r"\b(mediump|highp|lowp)\s+", "", src)
    src = re.sub(r"^(\s*)(uniform|in|out)\s+(?=\w)", r"\1", src, flags=re.M)
    src = re.sub(r"\bconst\s+in\b", "const", src)
    # array constructors  T[](a, b, ...)  ->  {a, b, ...}
    while True:
        m = re.search(r"\b\w+\s*\[\s*\]\s*\(", src)
        if not m:
            break
        close = _match_paren(src, m.end() - 1)
        src = src[:m.start()] + "{" + src[m.end():close] + "}" + src[close + 1:]
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    src = src.replace("discard;", "{ shader_discarded = true; return; }")
    # the two C++-name-lookup patches (see the module docstring)
    if name == "distance_map_anisotropic.comp":
        assert "pos.x = start.x;" in src
        src = src.replace("pos.x = start.x;", "pos.x = start;")
    if name == "volume_render.frag":
        assert "sampler3D gradient;" in src and "texture(gradient, pos)" in src
        src = src.replace("sampler3D gradient;", "sampler3D gradient_sampler;").replace("texture(gradient, pos)", "texture(gradient_sampler, pos)")
    return src


# variant name -> (shader file, defines, harness macro)
def variants():
    v = {}
    v["gradient_map"] = ("gradient_map.comp", [], "HARNESS_GRADIENT_MAP")
    for pre in (0, 1):
        d = ["PRECOMPUTED_GRADIENT"] if pre else []
        v[f"occupancy_map_p{pre}"] = ("occupancy_map.comp", d, "HARNESS_OCCUPANCY_MAP")
        v[f"occupied_voxel_count_p{pre}"] = ("occupied_voxel_count.comp", d, "HARNESS_VOXEL_COUNT")
    for s in (8, 32, 64):
        v[f"occupied_voxel_count_reduce_s{s}"] = ("occupied_voxel_count_reduce.comp", [f"SUBGROUP_SIZE {s}"], "HARNESS_VOXEL_COUNT_REDUCE")
    v["distance_map"] = ("distance_map.comp", [], "HARNESS_DISTANCE_MAP")
    v["distance_map_anisotropic"] = ("distance_map_anisotropic.comp", [], "HARNESS_DISTANCE_MAP_ANISO")
    v["vert_clipped"] = ("volume_render_clipped.vert", [], "HARNESS_VERT_CLIPPED")
    v["vert_plane"] = ("volume_render_plane_intersection.vert", [], "HARNESS_VERT_PLANE")
    skip_defs = {0: ["DISABLE_SKIP"], 1: ["BLOCK_SKIP"], 2: [], 3: ["ANISOTROPIC_DISTANCE"]}
    test_defs = {0: [], 1: ["SHOW_RAY_ENTRY"], 2: ["SHOW_RAY_EXIT"], 3: ["SHOW_NUM_SAMPLES"]}
    for pre, skip, ert, test in itertools.product((0, 1), (0, 1, 2, 3), (0, 1), (0, 1, 2, 3)):
        if pre == 0 and (test != 0 or skip not in (0, 2)):
            continue        # on-the-fly gradient: two representative variants only
        if test in (1, 2) and (skip != 2 or ert != 1):
            continue
        d = (["PRECOMPUTED_GRADIENT"] if pre else []) + skip_defs[skip] + ([] if ert else ["DISABLE_EARLY_RAY_TERMINATION"]) + test_defs[test]
        v[f"frag_p{pre}_s{skip}_e{ert}_t{test}"] = ("volume_render.frag", d, "HARNESS_FRAG")
    # DEPTH_ATTACHMENT (options.depth_attachment, volume_render_subpass.cpp:77-80): the march variants with ERT on + the ray-exit view
    for skip, test in ((0, 0), (1, 0), (2, 0), (3, 0), (2, 2)):
        d = ["PRECOMPUTED_GRADIENT", "DEPTH_ATTACHMENT"] + skip_defs[skip] + test_defs[test]
        v[f"frag_p1_s{skip}_e1_t{test}_d1"] = ("volume_render.frag", d, "HARNESS_FRAG")
    return v


def build(verbose=False) -> Path:
    if not SHADERS.exists():
        raise SystemExit("/root/reference is not present: oracle/_ref can only be built in the build container")
    GEN.mkdir(parents=True, exist_ok=True)
    incs = {}
    for vname, (shader, defs, harness) in variants().items():
        if shader not in incs:
            inc = GEN / (shader + ".inc")
            inc.write_text(transliterate(shader))
            incs[shader] = inc
        tu = GEN / f"{vname}.cpp"
        tu.write_text(
            "// generated by oracle/ref_shim/build_ref.py — do not commit\n"
            f'#include "{HERE / "glsl_compat.h"}"\n'
            + "".join(f"#define {d}\n" for d in defs)
            + f"#define VARIANT {vname}\n#define {harness} 1\n"
            "namespace {\nusing namespace glsl;\nstatic bool shader_discarded = false;\n"
            + ("static vec4 gl_Position;        // built-in of the `#version 320 es` vertex shader (not redeclared there)\n" if harness == "HARNESS_VERT_PLANE" else "")
            +
            f'#include "{incs[shader]}"\n'
            f'#include "{HERE / "ref_harness.h"}"\n'
            "}\n"
            f'#include "{HERE / "ref_exports.h"}"\n')
    srcs = sorted(GEN.glob("*.cpp"))
    objs = []

    def cc(src: Path, extra=()):
        obj = src.with_suffix(".o") if src.parent == GEN else GEN / (src.stem + ".o")
        cmd = ["g++", "-O2", "-ffp-contract=off", "-fPIC", "-std=gnu++17", "-w", "-c", str(src), "-o", str(obj), *extra]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + p.stderr[:6000])
            raise SystemExit(f"failed to compile {src.name}")
        return obj

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(cc, srcs))
    # the reference's loader, as is
    objs.append(cc(HERE / "ref_loader.cpp", ("-I", str(HERE / "boost_stub"), "-I", str(GLM_INC), "-I", str(VK_INC), "-I", str(REF / "src"),
                                             "-DGLM_FORCE_SWIZZLE", "-DGLM_FORCE_RADIANS", "-DGLM_FORCE_CTOR_INIT", "-DGLM_FORCE_DEPTH_ZERO_TO_ONE", "-DGLM_ENABLE_EXPERIMENTAL",
                                             "-D_MSC_EXTENSIONS",        # glm: member swizzles (`v.xyz`) as under MSVC, the reference's only toolchain
                                             f'-DREF_LOAD_VOLUME_CPP="{REF / "src" / "load_volume.cpp"}"')))
    objs.append(cc(HERE / "ref_glm.cpp", ("-I", str(GLM_INC), "-DGLM_FORCE_SWIZZLE", "-DGLM_FORCE_CTOR_INIT", "-DGLM_FORCE_RADIANS", "-DGLM_FORCE_DEPTH_ZERO_TO_ONE", "-DGLM_ENABLE_EXPERIMENTAL")))
    lib = OUT / "libvkv_ref.so"
    subprocess.run(["g++", "-shared", "-o", str(lib), *map(str, objs)], check=True)
    if verbose:
        print(f"built {lib} from {len(srcs)} shader variants + load_volume.cpp + glm host maths")
    return lib


if __name__ == "__main__":
    build(verbose=True)
