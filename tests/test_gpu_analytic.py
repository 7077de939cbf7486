"""Closed-form pins (tests/analytic_cases.py) applied to the CUDA kernels through the C ABI: the parts of the path the
reference cannot pin (SURVEY.md §8a last row) are tied here to answers that come from no code of this repository —
energy conservation (white furnace), L = albedo for a convex Lambertian body, an emitter seen directly, Snell / Schlick /
total internal reflection, the moments of the cosine lobe, gradient noise vanishing on the lattice, camera geometry."""
import numpy as np
import pytest

import analytic_cases as ac
from aurora_rendering_engine_b200 import capi, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,traversal,variant", [
    ("mixed", 1, capi.KERNEL_BRUTE_BAKED), ("mixed", 1, capi.KERNEL_BRUTE), ("mixed", 2, capi.KERNEL_BVH2), ("mixed", 3, capi.KERNEL_WIDE),
    ("lean", 1, capi.KERNEL_BRUTE_BAKED), ("lean", 1, capi.KERNEL_BRUTE_LEAN), ("lean", 1, capi.KERNEL_BRUTE), ("lean", 2, capi.KERNEL_BVH2),
])
def test_white_furnace(ctx, kind, traversal, variant):
    """All albedos 1, uniform environment of radiance 1, nothing absorbs: EVERY sample returns exactly 1.0f, so every
    accumulator entry is exactly the sample count — through every render kernel."""
    sc = ac.furnace_scene(kind)
    spp = 16
    if kind == "lean":   # the three brute-force kernels a scene with a lean form can run on
        ctx.set_option(capi.OPT_BAKED_KERNEL, int(variant == capi.KERNEL_BRUTE_BAKED))
        ctx.set_option(capi.OPT_LEAN_KERNEL, int(variant != capi.KERNEL_BRUTE))
    else:                # spheres, glass, mirrors: the generic brute-force kernel, or its scene-specialised (baked) form
        ctx.set_option(capi.OPT_BAKED_KERNEL, int(variant != capi.KERNEL_BRUTE))
    img, st = ac.render(ctx, sc, traversal=traversal, sample_count=spp)
    assert st.kernel_variant == variant and st.rays > 2 * st.samples
    lost = int(round(float((1.0 - img).sum() / 3 * spp)))
    assert np.array_equal(img, np.ones_like(img)), f"{lost} of {st.samples} samples lost, max deviation {np.abs(img - 1).max()}"


@pytest.mark.parametrize("shape", ["sphere", "quad", "triangle"])
@pytest.mark.parametrize("traversal", [1, 2])
def test_convex_lambertian_returns_its_albedo(ctx, shape, traversal):
    rho = np.array([0.2, 0.5, 0.8])
    sc = ac.albedo_scene(shape, tuple(rho))
    img, _ = ac.render(ctx, sc, traversal=traversal, sample_count=8)
    inner = img[24:40, 24:40]
    assert np.abs(inner - rho.astype(np.float32)).max() < 1e-6, np.abs(inner - rho).max()
    assert np.array_equal(img[0, 0], np.ones(3)) or shape == "triangle"


def test_emitter_seen_directly(ctx):
    img, _ = ac.render(ctx, ac.emitter_scene(), sample_count=4)
    assert np.array_equal(img[20:28, 20:28], np.broadcast_to(np.array([3.0, 2.0, 0.5]), (8, 8, 3)))
    assert not img[0, 0].any()


@pytest.mark.parametrize("precision,tol", [(64, 1e-12), (32, 1e-5)])
def test_dielectric_closed_form(ctx, precision, tol):
    sc = scenes.SceneDesc("d")
    t = sc.solid(1, 1, 1)
    m = sc.mat(scenes.MAT_DIELECTRIC, 1.5)
    sc.sphere((0, 0, 0), 1.0, m, t)
    sc.feed(ctx)
    ctx.commit()
    rng = np.random.RandomState(3)
    n = 200_000
    nrm, wi, rnd = ac.unit(rng.normal(size=(n, 3))), ac.unit(rng.normal(size=(n, 3))), rng.uniform(size=(n, 4))
    wo, att, emit, alive = ctx.scatter_batch(np.full(n, m), np.full(n, t), wi, nrm, np.zeros((n, 3)), np.zeros((n, 2)), rnd, precision=precision)
    exp, refl, margin = ac.dielectric_expectation(wi, nrm, 1.5, rnd[:, 0])
    # decisions away from the Schlick threshold / the critical angle (fp32 rounds the 24-bit uniform and the cosine)
    cos_i = np.abs(np.sum(wi * nrm, axis=1))
    ok = (margin > 1e-5) & (np.abs(1.5 * 1.5 * (1 - cos_i ** 2) - 1.0) > 1e-4)
    assert ok.mean() > 0.99 and alive.all() and np.allclose(att, 1.0) and not emit.any()
    assert np.abs(wo[ok] - exp[ok]).max() < tol, np.abs(wo[ok] - exp[ok]).max()


@pytest.mark.parametrize("precision", [64, 32])
def test_cosine_lobe_moments(ctx, precision):
    sc = scenes.SceneDesc("c")
    t = sc.solid(1, 1, 1)
    m = sc.mat(scenes.MAT_LAMBERTIAN, -1)
    sc.sphere((0, 0, 0), 1.0, m, t)
    sc.feed(ctx)
    ctx.commit()
    n = 1_000_000
    rng = np.random.RandomState(5)
    nrm = np.tile(ac.unit(np.array([[0.3, -0.5, 0.8]])), (n, 1))
    wi = np.tile(ac.unit(np.array([[0.1, 0.2, -1.0]])), (n, 1))
    wo, att, emit, alive = ctx.scatter_batch(np.full(n, m), np.full(n, t), wi, nrm, np.zeros((n, 3)), np.zeros((n, 2)), rng.uniform(size=(n, 4)),
                                             precision=precision)
    c = wo @ nrm[0]
    assert alive.all() and (c >= -1e-6).all()
    assert abs(c.mean() - 2 / 3) < 1.5e-3 and abs((c * c).mean() - 0.5) < 1.5e-3
    tang = wo - c[:, None] * nrm[0]
    assert np.abs(tang.mean(axis=0)).max() < 2e-3 and abs((tang ** 2).sum(axis=1).mean() - 0.5) < 1.5e-3


@pytest.mark.parametrize("precision,tol", [(64, 1e-12), (32, 2e-5)])
def test_perlin_vanishes_on_the_lattice(ctx, precision, tol):
    sc = scenes.SceneDesc("n")
    t = sc.tex(scenes.TEX_NOISE, 4.0, 2)
    sc.sphere((0, 0, 0), 1.0, sc.mat(scenes.MAT_LAMBERTIAN, -1), t)
    sc.feed(ctx)
    ctx.commit()
    rng = np.random.RandomState(1)
    P = rng.randint(-40, 40, (20000, 3)).astype(np.float64)
    rgb = ctx.texture_batch(np.full(len(P), t), np.zeros((len(P), 2)), P, precision=precision)
    want = 0.5 * (1.0 + np.sin(4.0 * P[:, 2]))
    assert np.abs(rgb - want[:, None]).max() < tol, np.abs(rgb - want[:, None]).max()
    g = ctx.texture_batch(np.full(len(P), t), np.zeros((len(P), 2)), P + rng.uniform(0.05, 0.95, P.shape), precision=precision)[:, 0]
    assert (g >= 0).all() and (g <= 1).all() and g.std() > 0.1


@pytest.mark.parametrize("precision,tol", [(64, 1e-14), (32, 1e-6)])
def test_camera_closed_form(ctx, precision, tol):
    cam = capi.make_camera(pos=(1, 2, 3), target=(4, 2, -1), up=(0, 1, 0), vfov_deg=50.0, focus_dist=1.0, jitter=1)
    W, H = 64, 48
    px, py = np.array([W // 2, 0, W - 1]), np.array([H // 2, 0, H - 1])
    rnd = np.array([[0.0, 0.0, 0, 0], [0.0, 0.0, 0, 0], [1.0, 1.0, 0, 0]])
    Q, D = ctx.camera_rays(cam, W, H, px, py, rnd, precision=precision)
    fwd = ac.unit(np.array([3.0, 0.0, -4.0]))
    right = ac.unit(np.cross(fwd, [0, 1, 0]))
    up = np.cross(right, fwd)
    s = np.tan(np.radians(25.0))
    assert np.abs(Q - np.array([1, 2, 3])).max() == 0 and np.abs(D[0] - fwd).max() < tol
    for d, sx, sy in ((D[1], -1, 1), (D[2], 1, -1)):
        assert np.abs(d - ac.unit(fwd + right * (sx * s * W / H) + up * (sy * s))).max() < tol
