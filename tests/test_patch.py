"""Patch-as-viewport renderer: the reference's own rendering algorithm, experiments/rt10.cpp.

CPU: oracle/patch_oracle.cpp (restatement) against tests/golden/patch_vectors.npz — outputs of the REAL rt10.cpp
     recorded by tests/golden/make_patch_golden.py, including the sha256 of the image the reference ships
     (experiments/output_rt10.ppm) — and live against oracle/_ref/librt10_ref.so when it is present; the product's
     host planner (no GPU needed) against the oracle's node / texel counts.
GPU: are_cuda_patch_render / are_cuda_patch_trace_texture through the C ABI, bit-identical (fp64) to the oracle and to
     the golden file: the reference's shipped image is reproduced byte for byte by the kernels.
"""
import dataclasses
import hashlib
import os

import numpy as np
import pytest

from aurora_rendering_engine_b200 import capi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "patch_vectors.npz")
N_RANDOM = 6
TEX_TRIS = (10, 12, 17, 23)


def random_scene(k):  # must match tests/golden/make_patch_golden.py
    return scenes.patch_random(k, width=120, height=90, mirror_walls=(k % 2 == 1))


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def porc():
    from oracle_binding import PatchOracle
    return PatchOracle()


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


# ------------------------------------------------------------------------------------------------- CPU: oracle pins
def test_oracle_reproduces_the_reference_shipped_image(porc, golden):
    ps = scenes.patch_rt10()
    rgb, rgb8, cnt = porc.render(ps, want_counters=True)
    assert np.array_equal(_sha(rgb8), golden["rt10_sha256"]), "P6 payload differs from experiments/output_rt10.ppm"
    rows = golden["rt10_row_index"]
    assert np.array_equal(rgb8[rows], golden["rt10_rows"])
    assert np.array_equal(rgb[rows[::8]], golden["rt10_rgb_rows"])  # linear fp64, bit for bit
    assert int(cnt[0]) == 44 and int(cnt[1]) == 28249
    shipped = "/root/reference/experiments/output_rt10.ppm"
    if os.path.exists(shipped):
        blob = open(shipped, "rb").read()
        assert blob == b"P6\n900 650\n255\n" + rgb8.tobytes()


def test_oracle_matches_reference_golden_on_random_scenes(porc, golden):
    for k in range(N_RANDOM):
        ps = random_scene(k)
        rgb, rgb8 = porc.render(ps)
        assert np.array_equal(rgb, golden[f"rand{k}_rgb"]), k
        assert np.array_equal(rgb8, golden[f"rand{k}_rgb8"]), k
        for j, cur in enumerate(TEX_TRIS):
            t = porc.trace_texture(ps, ps.origin, cur, 40 + 30 * j, 300 - 60 * j, 0.0)
            want = golden[f"rand{k}_tex{j}"]
            assert t.shape == want.shape and np.array_equal(t, want), (k, j)


def test_oracle_matches_live_reference(porc):
    from oracle_binding import PatchReference
    if not PatchReference.available():
        pytest.skip("oracle/_ref/librt10_ref.so not present")
    ref = PatchReference()
    cases = [scenes.patch_rt10(width=300, height=200, max_depth=2)]
    cases += [scenes.patch_random(100 + k, n_boxes=2 + k % 3, n_loose=4 * (k % 2), width=90 + 7 * k, height=60 + 5 * k, max_depth=k % 5, mirror_walls=k % 3 == 0)
              for k in range(8)]
    for ps in cases:
        a, b = porc.render(ps), ref.render(ps)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), ps.name
    ps = cases[3]
    for cur in range(0, len(ps.material), 3):
        for est in (0.0, 1.0, 50.0):
            ta, tb = porc.trace_texture(ps, ps.origin, cur, 33, 77, est), ref.trace_texture(ps, ps.origin, cur, 33, 77, est)
            assert ta.shape == tb.shape and np.array_equal(ta, tb), (cur, est)


# ------------------------------------------------------------------------------------------------- CPU: host planner
def test_planner_counts_match_oracle(lib, porc):
    for ps in [scenes.patch_rt10()] + [random_scene(k) for k in range(N_RANDOM)]:
        probe = capi.patch_plan_probe(ps)
        cnt = porc.render(ps, want_counters=True)[2]
        assert probe["nodes"] == int(cnt[0]) and probe["node_texels"] == int(cnt[1]), ps.name
        assert probe["levels"] <= ps.max_depth and probe["ops"] >= probe["ops_a"] + probe["ops_b"]


def test_planner_edge_cases(lib):
    ps = scenes.patch_rt10(width=64, height=48)
    empty = dataclasses.replace(ps, P=np.zeros((0, 3, 3)), UV=np.zeros((0, 3, 2)), material=np.zeros(0, np.int32))
    assert capi.patch_plan_probe(empty) == dict(nodes=0, node_texels=0, ops=0, levels=0, ops_a=0, ops_b=0)
    flat = dataclasses.replace(ps, max_depth=0)  # every reflective triangle terminates at once: no node textures
    p = capi.patch_plan_probe(flat)
    assert p["nodes"] == 0 and p["ops"] > 0
    degenerate = dataclasses.replace(ps, P=np.concatenate([ps.P, np.zeros((1, 3, 3))]), UV=np.concatenate([ps.UV, np.zeros((1, 3, 2))]),
                                     material=np.concatenate([ps.material, [3]]).astype(np.int32))
    assert capi.patch_plan_probe(degenerate)["nodes"] == capi.patch_plan_probe(ps)["nodes"]  # a zero-area triangle paints nothing


def test_patch_struct_layouts_match_header(tmp_path):
    import ctypes as C
    import subprocess
    f = tmp_path / "s.c"
    f.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "are_cuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(are_patch_scene), '
                 'offsetof(are_patch_scene, mat_albedo), sizeof(are_patch_config), offsetof(are_patch_config, env), offsetof(are_patch_config, gamma), '
                 'sizeof(are_patch_stats), offsetof(are_patch_stats, kernel_ms));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    exp = [C.sizeof(capi.PatchSceneC), capi.PatchSceneC.mat_albedo.offset, C.sizeof(capi.PatchConfig), capi.PatchConfig.env.offset,
           capi.PatchConfig.gamma.offset, C.sizeof(capi.PatchStats), capi.PatchStats.kernel_ms.offset]
    assert got == exp


# ------------------------------------------------------------------------------------------------- GPU: parity
@pytest.mark.gpu
def test_gpu_reproduces_the_reference_shipped_image(ctx, porc, golden):
    ps = scenes.patch_rt10()
    rgb, rgb8, st = ctx.patch_render(ps)
    assert np.array_equal(_sha(rgb8), golden["rt10_sha256"]), "GPU P6 payload differs from the reference's experiments/output_rt10.ppm"
    orgb, orgb8 = porc.render(ps)
    assert np.array_equal(rgb, orgb), "linear fp64 image is not bit-identical to the oracle"
    assert np.array_equal(rgb8, orgb8)
    assert st.nodes == 44 and st.node_texels == 28249 and st.launches == st.levels + 1 and st.kernel_ms > 0
    print(f"[patch rt10] plan {st.plan_ms:.3f} ms, kernels {st.kernel_ms:.3f} ms, {st.launches} launches, {st.ops} warp triangles")


@pytest.mark.gpu
def test_gpu_matches_golden_and_oracle_on_random_scenes(ctx, porc, golden):
    for k in range(N_RANDOM):
        ps = random_scene(k)
        rgb, rgb8, _ = ctx.patch_render(ps)
        assert np.array_equal(rgb, golden[f"rand{k}_rgb"]), k
        assert np.array_equal(rgb8, golden[f"rand{k}_rgb8"]), k
        for j, cur in enumerate(TEX_TRIS):
            t, _ = ctx.patch_trace_texture(ps, ps.origin, cur, 40 + 30 * j, 300 - 60 * j, 0.0)
            want = golden[f"rand{k}_tex{j}"]
            assert t.shape == want.shape and np.array_equal(t, want), (k, j)
    for k in range(8):  # more shapes / depths than the golden file holds, against the oracle
        ps = scenes.patch_random(100 + k, n_boxes=2 + k % 3, n_loose=4 * (k % 2), width=90 + 7 * k, height=60 + 5 * k, max_depth=k % 5, mirror_walls=k % 3 == 0)
        a, b = ctx.patch_render(ps), porc.render(ps)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), ps.name


@pytest.mark.gpu
def test_gpu_edge_cases(ctx, porc):
    ps = scenes.patch_rt10(width=64, height=48)
    empty = dataclasses.replace(ps, P=np.zeros((0, 3, 3)), UV=np.zeros((0, 3, 2)), material=np.zeros(0, np.int32))
    flat = dataclasses.replace(ps, max_depth=0)
    tiny_tex = dataclasses.replace(ps, max_tex_res=3, min_tex_res=1)
    odd = dataclasses.replace(ps, width=2, height=2)
    for case in (empty, flat, tiny_tex, odd):
        a, b = ctx.patch_render(case), porc.render(case)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rgb = ctx.patch_render(empty)[0]
    assert np.allclose(rgb[5, 5], ps.env)  # nothing but the environment colour
    # a diffuse triangle's texture is its albedo, at the clamped size
    t, st = ctx.patch_trace_texture(ps, ps.origin, 0, 1000, 5, 0.0)
    assert t.shape == (16, 256, 3) and np.all(t == np.array([0.85, 0.85, 0.85])) and st.nodes == 0
    with pytest.raises(capi.AreCudaError):
        ctx.patch_trace_texture(ps, ps.origin, 999, 16, 16)
    with pytest.raises(capi.AreCudaError):
        ctx.patch_render(dataclasses.replace(ps, width=1))


@pytest.mark.gpu
def test_gpu_full_size_properties(ctx, porc):
    """4K camera image with 1024^2 node textures (the bench workload): equal to the oracle, and invariant properties."""
    ps = dataclasses.replace(scenes.patch_rt10(width=3840, height=2160), max_tex_res=1024)
    rgb, rgb8, st = ctx.patch_render(ps)
    orgb, orgb8 = porc.render(ps)
    assert np.array_equal(rgb, orgb) and np.array_equal(rgb8, orgb8)
    assert rgb.min() >= 0.0 and rgb.max() <= 1.0
    again = ctx.patch_render(ps)
    assert np.array_equal(again[0], rgb)  # deterministic: no atomics, no order dependence
    print(f"[patch 4K] plan {st.plan_ms:.3f} ms, kernels {st.kernel_ms:.3f} ms, {st.node_texels} node texels, {st.ops} warp triangles")


@pytest.mark.gpu
def test_reference_style_host_program_patch_renders_on_gpu(tmp_path, lib, ctx, golden):
    """examples/patch_host.cpp builds the rt10 room from are::Triangle / Diffuse / Reflective / Texture objects and renders
    it through are::cuda::patch_render / trace_texture (include/are_cuda.hpp).  are::Triangle stores (Q, u, v), so its
    vertices are Q + (P - Q): the Python-driven render of exactly those vertices must be identical, and the image is the
    reference's shipped one up to that one-ulp vertex perturbation."""
    import subprocess
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = tmp_path / "patch_host"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "patch_host.cpp"),
                    "-L", libdir, "-lare_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    ppm, raw, tex = tmp_path / "out.ppm", tmp_path / "rgb.f64", tmp_path / "tex.f64"
    r = subprocess.run([str(exe), str(ppm), "900", "650", str(raw), str(tex)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ps = scenes.patch_rt10()
    P = ps.P.copy()
    P[:, 1] = P[:, 0] + (P[:, 1] - P[:, 0])
    P[:, 2] = P[:, 0] + (P[:, 2] - P[:, 0])
    ps = dataclasses.replace(ps, P=P)
    rgb, rgb8, _ = ctx.patch_render(ps)
    got = np.fromfile(raw, np.float64).reshape(650, 900, 3)
    assert np.array_equal(got, rgb)
    blob = ppm.read_bytes()
    assert blob == b"P6\n900 650\n255\n" + rgb8.tobytes()
    shipped_rows = golden["rt10_rows"]
    diff = (rgb8[golden["rt10_row_index"]].astype(int) - shipped_rows.astype(int))
    assert (diff != 0).mean() < 1e-3  # one-ulp vertex differences may move a few edge pixels, nothing more
    t, _ = ctx.patch_trace_texture(ps, ps.origin, 18, 200, 200, 0.0)
    got_t = np.fromfile(tex, np.float64).reshape(t.shape)
    assert np.array_equal(got_t, t) and t.std() > 0.01  # a mirror face with the room painted into it
