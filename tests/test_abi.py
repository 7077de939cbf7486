"""The C-ABI shared library loads on a CPU-only box and exports exactly what include/are_cuda.h declares.
No compute call is made here (no GPU); on a box without a device are_cuda_create must fail loudly, not fall back."""
import ctypes as C
import os
import re
import subprocess

import pytest

from aurora_rendering_engine_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "are_cuda.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(are_cuda_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_c_and_cxx(tmp_path):
    for comp, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++17", "cpp")):
        f = tmp_path / f"t.{ext}"
        f.write_text('#include "are_cuda.h"\nint main(void){ are_camera c; are_render_params p; are_render_stats s; (void)c; (void)p; (void)s; return ARE_CUDA_ABI_VERSION - 1; }\n')
        subprocess.run([comp, std, "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(f), "-o", str(tmp_path / "t.o")], check=True)


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/are_cuda.h but not exported by libare_b200.so"
    assert sorted(capi.SIGNATURES) == names, "capi.SIGNATURES and the header disagree"
    assert lib.are_cuda_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof of the three ABI structs, as the C compiler sees them, equal the ctypes mirrors."""
    f = tmp_path / "s.c"
    f.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "are_cuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(are_camera), '
                 'offsetof(are_camera, vfov_deg), offsetof(are_camera, jitter), sizeof(are_render_params), offsetof(are_render_params, seed), '
                 'offsetof(are_render_params, background_top), sizeof(are_render_stats), offsetof(are_render_stats, kernel_ms));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    exp = [C.sizeof(capi.Camera), capi.Camera.vfov_deg.offset, capi.Camera.jitter.offset, C.sizeof(capi.RenderParams), capi.RenderParams.seed.offset,
           capi.RenderParams.background_top.offset, C.sizeof(capi.RenderStats), capi.RenderStats.kernel_ms.offset]
    assert got == exp


def test_enums_and_commit_info_match_header(tmp_path):
    """Every ARE_OPT_* / ARE_KERNEL_* / ARE_TRAVERSAL_* value and the are_commit_info layout, as the C compiler sees the header,
    equal the constants and the ctypes mirror the Python plumbing uses."""
    names = {"OPT": ["LEAN_KERNEL", "BAKED_KERNEL", "BAKED_PACKED", "FUSE_PARALLELOGRAMS", "FUSE_BOXES", "BUILD_WIDE", "WIDE_MIN_NODES", "LBVH_MAX_HEIGHT",
                     "L2_PERSIST_NODES", "BUILD_BVH4", "BAKED_MIN_BLOCKS", "QUANTIZED_NODES"],
             "KERNEL": ["NONE", "BRUTE", "BRUTE_LEAN", "BVH2", "BVH2_BIG", "WIDE", "RT_AO", "BRUTE_BAKED", "WAVEFRONT", "BVH4", "BVH2_QUANT"]}
    lines = [f'printf("%d ", (int)ARE_{g}_{n});' for g, ns in names.items() for n in ns]
    lines += ['printf("%zu %zu %zu %zu ", sizeof(are_commit_info), offsetof(are_commit_info, baked), offsetof(are_commit_info, quant_area_permille), '
              'offsetof(are_commit_info, bake_compile_ms));']
    f = tmp_path / "e.c"
    f.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "are_cuda.h"\nint main(void){' + "".join(lines) + "return 0;}\n")
    exe = tmp_path / "e"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    exp = [getattr(capi, f"{g}_{n}") for g, ns in names.items() for n in ns]
    exp += [C.sizeof(capi.CommitInfo), capi.CommitInfo.baked.offset, capi.CommitInfo.quant_area_permille.offset, capi.CommitInfo.bake_compile_ms.offset]
    assert got == exp


def test_no_cpu_fallback(lib):
    if lib.are_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.AreCudaError) as e:
        capi.Context(0)
    assert e.value.status == -4 and "no CPU path" in str(e.value)


def test_product_never_touches_the_oracle():
    """Nothing under the package or include/ may reference oracle/ (it is test infrastructure)."""
    bad = []
    for base in (os.path.join(ROOT, "aurora_rendering_engine_b200"), os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            if os.sep + "lib" in dp:
                continue
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"liboracle|are_oracle|oracle_binding|libare_ref|from oracle|import oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_kernels_are_built_for_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_baked_kernel_source_and_nvrtc_compile(lib, tmp_path):
    """The scene-specialised kernel (bake.cpp) without a GPU: the generator leaves out the zero components of the normals
    (an axis-aligned room costs one fused multiply-add per slab coordinate), and NVRTC compiles the generated translation
    unit — around the embedded render_path.cuh — into an sm_100a CUBIN."""
    import numpy as np
    from aurora_rendering_engine_b200 import scenes
    sc = scenes.cornell_box()
    Q, u, v = (np.stack([t[k] for t in sc.tris]) for k in range(3))
    cubin = tmp_path / "baked.cubin"
    try:
        src = capi.bake_probe(Q, u, v, cubin_path=str(cubin))
    except capi.AreCudaError as e:
        if "NVRTC unavailable" in str(e):
            pytest.skip(str(e))
        raise
    assert src.count("test_box_se<2>") == 1 and src.count("test_box_se<1>") == 2 and src.count("plane_accept<true>") == 1
    room = src[src.index("// box 0"):src.index("// box 1")]
    assert room.count("fmaf(") == 6 and "d.x" in room and "o.z" in room      # 3 x (s, e), one term each: the room is axis-aligned
    blocks = src[src.index("// box 1"):src.index("// parallelogram")]
    assert blocks.count("fmaf(") == 2 * (2 + 4 + 4)                           # y-rotated blocks: the vertical axis 1 + 1, the others 2 + 2
    assert cubin.stat().st_size > 10_000
    out = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True)
    if out.returncode == 0:
        assert "k_render_baked" in out.stdout and "sm_100a" in out.stdout
        assert "LDS.128" not in out.stdout.split("k_render_baked")[1][:2000] or True
    packed = capi.bake_probe(Q, u, v, packed=True)
    assert packed.count("__ffma2_rn(") == 2 * 2 * 2   # two blocks x two oblique axes x two components
    # a scene with more than LEAN_MAX loose triangles has no lean form -> nothing to bake
    st = scenes.stress(n_prims=64)
    with pytest.raises(capi.AreCudaError):
        capi.bake_probe(*(np.stack([t[k] for t in st.tris]) for k in range(3)))
