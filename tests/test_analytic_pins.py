"""Closed-form pins (tests/analytic_cases.py) applied to the CPU twin — so that the twin the GPU tests compare with is
itself tied to something outside this repository's code.  The same cases run against the CUDA kernels in
tests/test_gpu_analytic.py."""
import numpy as np
import pytest

import analytic_cases as ac
from aurora_rendering_engine_b200 import capi, scenes


@pytest.mark.parametrize("kind", ["mixed", "lean"])
def test_twin_white_furnace(oracle, kind):
    sc = ac.furnace_scene(kind, width=48, height=32)
    img, st = ac.render(oracle.scene(), sc, sample_count=4)
    assert st.rays > 2 * st.samples
    assert np.abs(img - 1.0).max() < 1e-12, np.abs(img - 1.0).max()


@pytest.mark.parametrize("shape", ["sphere", "quad", "triangle"])
def test_twin_convex_lambertian_returns_its_albedo(oracle, shape):
    rho = np.array([0.2, 0.5, 0.8])
    sc = ac.albedo_scene(shape, tuple(rho), width=32, height=32)
    img, _ = ac.render(oracle.scene(), sc, sample_count=4)
    inner = img[12:20, 12:20]
    assert np.abs(inner - rho).max() < 1e-12
    assert np.abs(img[0, 0] - 1.0).max() < 1e-12 or shape != "sphere"


def test_twin_emitter(oracle):
    img, _ = ac.render(oracle.scene(), ac.emitter_scene(), sample_count=2)
    assert np.abs(img[20:28, 20:28] - np.array([3.0, 2.0, 0.5])).max() < 1e-12
    assert not img[0, 0].any()


def _dielectric_inputs(n=20000, seed=3):
    rng = np.random.RandomState(seed)
    nrm = ac.unit(rng.normal(size=(n, 3)))
    wi = ac.unit(rng.normal(size=(n, 3)))
    rnd = rng.uniform(size=(n, 4))
    return wi, nrm, rnd


def test_twin_dielectric_closed_form(oracle):
    sc = scenes.SceneDesc("d")
    t = sc.solid(1, 1, 1)
    m = sc.mat(scenes.MAT_DIELECTRIC, 1.5)
    osc = sc.feed(oracle.scene())
    wi, nrm, rnd = _dielectric_inputs()
    n = len(wi)
    wo, att, emit, alive = osc.scatter_batch(np.full(n, m), np.full(n, t), wi, nrm, np.zeros((n, 3)), np.zeros((n, 2)), rnd)
    exp, refl, margin = ac.dielectric_expectation(wi, nrm, 1.5, rnd[:, 0])
    ok = margin > 1e-9
    assert alive.all() and np.allclose(att, 1.0) and not emit.any()
    assert np.abs(wo[ok] - exp[ok]).max() < 1e-12
    assert 0.02 < refl.mean() < 0.9 and np.abs(np.linalg.norm(wo, axis=1) - 1).max() < 1e-12


def test_twin_cosine_lobe_moments(oracle):
    sc = scenes.SceneDesc("c")
    t = sc.solid(1, 1, 1)
    m = sc.mat(scenes.MAT_LAMBERTIAN, -1)
    osc = sc.feed(oracle.scene())
    n = 400_000
    rng = np.random.RandomState(5)
    nrm = np.tile(ac.unit(np.array([[0.3, -0.5, 0.8]])), (n, 1))
    wi = np.tile(ac.unit(np.array([[0.1, 0.2, -1.0]])), (n, 1))
    wo, att, emit, alive = osc.scatter_batch(np.full(n, m), np.full(n, t), wi, nrm, np.zeros((n, 3)), np.zeros((n, 2)), rng.uniform(size=(n, 4)))
    c = wo @ nrm[0]
    assert alive.all() and (c >= 0).all()
    assert abs(c.mean() - 2 / 3) < 2e-3 and abs((c * c).mean() - 0.5) < 2e-3
    tang = wo - c[:, None] * nrm[0]
    assert np.abs(tang.mean(axis=0)).max() < 3e-3 and abs((tang ** 2).sum(axis=1).mean() - 0.5) < 2e-3


def test_twin_perlin_vanishes_on_the_lattice(oracle):
    sc = scenes.SceneDesc("n")
    t = sc.tex(scenes.TEX_NOISE, 4.0, 2)
    osc = sc.feed(oracle.scene())
    rng = np.random.RandomState(1)
    P = rng.randint(-40, 40, (5000, 3)).astype(np.float64)
    rgb = osc.texture_batch(np.full(len(P), t), np.zeros((len(P), 2)), P)
    want = 0.5 * (1.0 + np.sin(4.0 * P[:, 2]))
    assert np.abs(rgb - want[:, None]).max() < 1e-12
    # off the lattice it is noise: bounded, not constant
    Pf = P + rng.uniform(0.05, 0.95, P.shape)
    g = osc.texture_batch(np.full(len(P), t), np.zeros((len(P), 2)), Pf)[:, 0]
    assert (g >= 0).all() and (g <= 1).all() and g.std() > 0.1


def test_twin_camera_closed_form(oracle):
    cam = capi.make_camera(pos=(1, 2, 3), target=(4, 2, -1), up=(0, 1, 0), vfov_deg=50.0, focus_dist=1.0, jitter=1)
    W, H = 64, 48
    px, py = np.array([W // 2, 0, W - 1]), np.array([H // 2, 0, H - 1])
    rnd = np.array([[0.0, 0.0, 0, 0], [0.0, 0.0, 0, 0], [1.0, 1.0, 0, 0]])
    Q, D = oracle.camera_rays(cam, W, H, px, py, rnd)
    fwd = ac.unit(np.array([3.0, 0.0, -4.0]))
    assert np.abs(Q - np.array([1, 2, 3])).max() == 0 and np.abs(D[0] - fwd).max() < 1e-15
    # opposite image corners: the vertical half-angle is vfov/2, the horizontal one atan(aspect * tan(vfov/2))
    right = ac.unit(np.cross(fwd, [0, 1, 0]))
    up = np.cross(right, fwd)
    s = np.tan(np.radians(25.0))
    for d, sx, sy in ((D[1], -1, 1), (D[2], 1, -1)):
        want = ac.unit(fwd + right * (sx * s * W / H) + up * (sy * s))
        assert np.abs(d - want).max() < 1e-14
