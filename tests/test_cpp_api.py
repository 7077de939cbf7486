"""The C++ class API (include/basic, include/material, include/object, texture.h) is a drop-in for the reference's.

CPU: tests/cpp/host_api_probe.cpp uses only the reference's public API; built against this repository's header-only
mirror it must print byte-for-byte what it prints when built against the REAL reference library (committed as
tests/golden/host_api_probe.txt; regenerated live when /root/reference is present).
GPU: examples/cornell_host.cpp builds the Cornell box with are::Triangle / are::ObjectSet like a reference user would
and renders it through are::cuda::Renderer; the image equals the one the Python driver renders from the same scene.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "tests", "cpp", "host_api_probe.cpp")
GOLDEN = os.path.join(ROOT, "tests", "golden", "host_api_probe.txt")
REF = "/root/reference"


def _run_probe(exe, workdir):
    return subprocess.run([exe, str(workdir)], capture_output=True, text=True, check=True).stdout


def test_mirror_headers_match_reference_golden(tmp_path):
    exe = tmp_path / "probe_mirror"
    subprocess.run(["g++", "-std=c++20", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), PROBE, "-o", str(exe)], check=True)
    out = _run_probe(str(exe), tmp_path)
    gold = open(GOLDEN).read()
    assert out == gold
    assert len(out.splitlines()) > 1000


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "texture.cpp")), reason="reference tree not present")
def test_mirror_headers_match_live_reference_build(tmp_path):
    import glob
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    subprocess.run(["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"), PROBE, "-o", str(a / "probe")], check=True)
    srcs = sorted(glob.glob(REF + "/src/basic/*.cpp") + glob.glob(REF + "/src/material/*.cpp") + glob.glob(REF + "/src/object/*.cpp")) + [REF + "/src/texture.cpp"]
    subprocess.run(["g++", "-std=c++20", "-O2", "-w", "-DARE_REFERENCE_BUILD", "-I", REF + "/include", PROBE] + srcs + ["-o", str(b / "probe")], check=True)
    oa, ob = _run_probe(str(a / "probe"), a), _run_probe(str(b / "probe"), b)
    assert oa == ob
    assert (a / "probe.ppm").read_bytes() == (b / "probe.ppm").read_bytes()
    assert ob == open(GOLDEN).read(), "tests/golden/host_api_probe.txt is stale: regenerate it from the reference build"


def test_new_types_compile_and_validate(tmp_path):
    """Sphere / Quad / new materials follow the reference's error conventions (invalid_argument from constructors)."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <object/object_set.h>
#include <material/lambertian.h>
#include <material/metal.h>
#include <material/dielectric.h>
#include <material/diffuse_light.h>
#include <camera.h>
#include <cstdio>
using namespace are;
int main() {
  Lambertian lam; Texture tex = Texture::solid(Color3(.5,.5,.5));
  int bad = 0;
  try { Sphere s(Point3(0,0,0), -1.0, &lam, &tex); } catch (const std::invalid_argument&) { ++bad; }
  try { Sphere s(Point3(0,0,0), 1.0, nullptr, &tex); } catch (const std::invalid_argument&) { ++bad; }
  try { Quad q(Point3(0,0,0), Vec3(1,0,0), Vec3(2,0,0), &lam, &tex); } catch (const std::invalid_argument&) { ++bad; }
  try { Dielectric d(0.0); } catch (const std::invalid_argument&) { ++bad; }
  Sphere s(Point3(0,0,-3), 1.0, &lam, &tex); Quad q(Point3(-1,-1,-2), Vec3(2,0,0), Vec3(0,2,0), &lam, &tex);
  Point3 h; Ray r(Point3(0,0,0), Vec3(0,0,-5));
  bool hs = s.intersect_ray(r, h); double zs = h.z();
  bool hq = q.intersect_ray(r, h); double zq = h.z();
  bool miss = q.intersect_ray(Ray(Point3(5,5,0), Vec3(0,0,-1)), h);
  Texture copy = tex;  // value semantics are safe here
  Metal m(2.0); DiffuseLight l(4.0); double p[8]; m.describe(p);
  std::printf("%d %d %.3f %d %.3f %d %d %.1f %d %d\n", bad, (int)hs, zs, (int)hq, zq, (int)miss, copy.kind(), p[0], m.kind(), l.kind());
  ObjectSet set; set.spheres.push_back(&s); set.quads.push_back(&q);
  return 0;
}''')
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["4", "1", "-2.000", "1", "-2.000", "0", "0", "1.0", "3", "5"]


@pytest.mark.gpu
def test_reference_style_host_program_renders_on_gpu(tmp_path, lib, ctx):
    from aurora_rendering_engine_b200 import capi, scenes
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = tmp_path / "cornell_host"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cornell_host.cpp"),
                    "-L", libdir, "-lare_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    W = H = 64
    spp = 8
    ppm, raw = tmp_path / "out.ppm", tmp_path / "sums.f32"
    r = subprocess.run([str(exe), str(ppm), str(W), str(H), str(spp), str(raw)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sums = np.fromfile(raw, np.float32).reshape(H, W, 3)
    sc = scenes.cornell_box(width=W, height=H)
    sc.feed(ctx)
    ctx.commit()
    img, st = ctx.render(capi.make_camera(**sc.camera_args()), capi.make_params(**sc.params_args(sample_count=spp)))
    assert np.array_equal(sums, img), "the C++ host program and the Python driver describe the same scene: images must be identical"
    # the same program over EVERY visible GPU (are::cuda::Renderer(devices) -> are_cuda_create_multi): same samples, sums equal up
    # to summation order (identical on a one-GPU box, where the group has one device)
    raw_all = tmp_path / "sums_all.f32"
    r = subprocess.run([str(exe), str(tmp_path / "out_all.ppm"), str(W), str(H), str(spp), str(raw_all), "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n_gpus = lib.are_cuda_device_count()
    assert f"on {n_gpus} GPU(s)" in r.stdout, r.stdout
    sums_all = np.fromfile(raw_all, np.float32).reshape(H, W, 3)
    assert np.allclose(sums_all, sums, rtol=1e-5, atol=1e-5) and (n_gpus > 1 or np.array_equal(sums_all, sums))
    # the PPM was written by are::Texture::save_texture: linear, truncating (reference src/texture.cpp:384-386)
    data = ppm.read_bytes()
    hdr = f"P6\n{W} {H}\n255\n".encode()
    assert data.startswith(hdr)
    px = np.frombuffer(data[len(hdr):], np.uint8).reshape(H, W, 3)
    # the shim divides the float sums by spp in double, exactly as done here
    mean = sums.astype(np.float64) * (1.0 / spp)
    exp = np.clip(mean * 255.0, 0.0, 255.0).astype(np.uint8)
    assert np.array_equal(px, exp)


def test_paste_restatement_matches_reference_golden(oracle):
    """oracle lib_texture_paste vs outputs of the real are::Texture::paste recorded in the golden file."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    changed = 0
    for c, want in zip(g["paste_corners"], g["paste_out"]):
        got = oracle.texture_paste(g["paste_dst"], g["paste_src"], c.reshape(4, 2))
        assert np.array_equal(got, want)
        changed += int((want != g["paste_dst"]).any(axis=2).sum())
    assert changed > 3000  # the cases really paste something
    assert np.array_equal(g["paste_out"][2], g["paste_dst"]) and np.array_equal(g["paste_out"][5], g["paste_dst"])  # degenerate / off-image


@pytest.mark.gpu
def test_gpu_paste_is_bit_identical(ctx, oracle):
    """are_cuda_texture_paste (fp64 kernel, -fmad=false) against the reference's outputs and against the restatement
    on fresh random quads."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    for c, want in zip(g["paste_corners"], g["paste_out"]):
        got = ctx.texture_paste(g["paste_dst"].copy(), g["paste_src"], c.reshape(4, 2))
        assert np.array_equal(got, want)
    rng = np.random.RandomState(11)
    pasted = 0
    for trial in range(30):
        dw, dh, sw, sh = rng.randint(16, 300), rng.randint(16, 300), rng.randint(1, 64), rng.randint(1, 64)
        dst, src = rng.uniform(0, 1, (dh, dw, 3)), rng.uniform(0, 1, (sh, sw, 3))
        corners = [tuple(int(v) for v in rng.randint(-20, max(dw, dh) + 20, 2)) for _ in range(4)]
        if trial % 3 == 0:
            corners = [(3, 2), (dw - 5, 4), (1, dh - 4), (dw - 2, dh - 3)]
        want = oracle.texture_paste(dst, src, corners)
        got = ctx.texture_paste(dst.copy(), src, corners)
        assert np.array_equal(got, want), (trial, corners)
        pasted += int((want != dst).any(axis=2).sum())
    assert pasted > 100_000
