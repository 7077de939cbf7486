"""The ray-pool kernel (csrc/pool.cu, ARE_OPT_POOL_KERNEL): the BVH2 path integrator with traversal decoupled from shading
inside the warp (lanes fetch rays from a shared-memory pool by ticket; hits are shaded 32 at a time, continuing rays
compacted in place).  Same estimator, same Philox counters, same device routines as the megakernel: identical ray counts,
images equal up to the order in which samples reach a pixel."""
import numpy as np
import pytest

from aurora_rendering_engine_b200 import capi, scenes
from oracle_binding import psnr

pytestmark = pytest.mark.gpu


def _pair(ctx, sc, spp, depth, **over):
    cam = capi.make_camera(**sc.camera_args())
    ctx.clear()
    sc.feed(ctx)
    ctx.commit()
    par = capi.make_params(**sc.params_args(sample_count=spp, traversal=2, max_depth=depth, **over))
    ctx.set_option(capi.OPT_POOL_KERNEL, 0)
    mega, sm = ctx.render(cam, par)
    ctx.set_option(capi.OPT_POOL_KERNEL, 1)
    acc = ctx.alloc_accum(sc.width, sc.height)
    sp = ctx.render_device(cam, par, acc, want_stats=True, count_tests=True)
    pool = ctx.download_accum(acc, sc.width, sc.height)
    ctx.free_accum(acc)
    pool2, sp2 = ctx.render(cam, par)   # the non-counting build
    ctx.set_option(capi.OPT_POOL_KERNEL, 0)
    assert np.array_equal(pool, pool2) and sp.rays == sp2.rays
    return mega, sm, pool, sp


@pytest.mark.parametrize("name,kw,spp,depth", [
    ("rtiow_final", dict(width=160, height=90), 8, 50),
    ("cornell_box", dict(width=96, height=96), 16, 50),
    ("textured", dict(width=128, height=72), 8, 12),
    ("stress", dict(n_prims=20_000, width=96, height=54), 4, 8),     # more than 16384 nodes: the 12-CTA build
    ("cornell_box", dict(width=37, height=23), 3, 6),                # ragged frame: border tiles skip tasks
    ("rtiow_final", dict(width=64, height=36), 1, 50),               # fewer tasks per tile than pool slots
])
def test_pool_kernel_equals_megakernel(ctx, oracle, name, kw, spp, depth):
    sc = scenes.by_name(name, **kw)
    mega, sm, pool, sp = _pair(ctx, sc, spp, depth)
    assert sp.kernel_variant == capi.KERNEL_POOL and sm.kernel_variant in (capi.KERNEL_BVH2, capi.KERNEL_BVH2_BIG)
    assert sp.rays == sm.rays and sp.samples == sm.samples, (sp.rays, sm.rays)
    assert sp.node_visits > 0
    assert np.allclose(pool, mega, rtol=2e-5, atol=2e-5 * spp), float(np.abs(pool - mega).max())
    if sc.width >= 96:
        cam = capi.make_camera(**sc.camera_args())
        oimg, ost = sc.feed(oracle.scene()).render(cam, capi.make_params(**sc.params_args(sample_count=spp, max_depth=depth)))
        assert abs(int(sp.rays) - int(ost.rays)) <= max(4, 1e-2 * ost.rays)
        assert psnr(np.clip(pool / spp, 0, 1), np.clip(oimg / spp, 0, 1)) >= 40.0


def test_pool_kernel_edge_cases(ctx):
    """Empty scene, one primitive (no hierarchy), zero samples, sample ranges adding up."""
    empty = scenes.SceneDesc("empty", width=37, height=23)
    empty.solid(0.5, 0.5, 0.5)
    empty.mat(scenes.MAT_LAMBERTIAN, -1)
    empty.camera = dict(pos=(0, 0, 0), target=(0, 0, -1), up=(0, 1, 0), vfov_deg=60.0, focus_dist=1.0, jitter=1)
    mega, sm, pool, sp = _pair(ctx, empty, 4, 8)
    assert sp.rays == sm.rays == 37 * 23 * 4 and np.allclose(pool, mega, rtol=1e-6, atol=1e-6)
    one = scenes.SceneDesc("one", width=40, height=30)
    g = one.solid(0.7, 0.3, 0.2)
    one.sphere((0, 0, -3), 1.0, one.mat(scenes.MAT_LAMBERTIAN, -1), g)
    one.camera = dict(pos=(0, 0, 0), target=(0, 0, -1), up=(0, 1, 0), vfov_deg=60.0, focus_dist=1.0, jitter=1)
    mega, sm, pool, sp = _pair(ctx, one, 8, 6)
    assert sp.rays == sm.rays and np.allclose(pool, mega, rtol=1e-5, atol=1e-5)
    sc = scenes.cornell_box(width=64, height=64)
    cam = capi.make_camera(**sc.camera_args())
    ctx.clear(); sc.feed(ctx); ctx.commit()
    ctx.set_option(capi.OPT_POOL_KERNEL, 1)
    acc = ctx.alloc_accum(64, 64)
    st0 = ctx.render_device(cam, capi.make_params(**sc.params_args(sample_count=0, traversal=2)), acc, want_stats=True)
    assert st0.rays == 0 and st0.launches == 0
    for b in (0, 5):
        ctx.render_device(cam, capi.make_params(**sc.params_args(sample_begin=b, sample_count=5 if b == 0 else 7, traversal=2, max_depth=10)), acc)
    two = ctx.download_accum(acc, 64, 64)
    ctx.free_accum(acc)
    one_go, _ = ctx.render(cam, capi.make_params(**sc.params_args(sample_begin=0, sample_count=12, traversal=2, max_depth=10)))
    ctx.set_option(capi.OPT_POOL_KERNEL, 0)
    assert np.allclose(two, one_go, rtol=2e-5, atol=1e-4)
