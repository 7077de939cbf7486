"""The wavefront schedule of the path integrator (csrc/wavefront.cu; north_star item 3): generate / extend / shade +
compact over ray queues.  It is the SAME estimator as the megakernel — same Philox counters, same device routines — so
the two must trace exactly the same rays and produce the same image up to the order in which samples reach a pixel."""
import numpy as np
import pytest

from aurora_rendering_engine_b200 import capi, scenes
from oracle_binding import psnr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kw,spp,depth", [
    ("rtiow_final", dict(width=160, height=90), 8, 50),
    ("cornell_box", dict(width=96, height=96), 16, 50),
    ("textured", dict(width=128, height=72), 8, 12),
    ("stress", dict(n_prims=20_000, width=96, height=54), 4, 8),
    ("cornell_box", dict(width=37, height=23), 3, 6),      # ragged frame, queue counts not multiples of a warp
])
def test_wavefront_equals_megakernel(ctx, oracle, name, kw, spp, depth):
    sc = scenes.by_name(name, **kw)
    cam = capi.make_camera(**sc.camera_args())
    sc.feed(ctx)
    ctx.commit()
    mega, sm = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=spp, traversal=2, max_depth=depth)))
    acc = ctx.alloc_accum(sc.width, sc.height)
    sw = ctx.render_device(cam, capi.make_params(**sc.params_args(sample_count=spp, traversal=2, max_depth=depth, integrator=scenes.INTEGRATOR_PATH_WAVEFRONT)),
                           acc, want_stats=True, count_tests=True)
    wave = ctx.download_accum(acc, sc.width, sc.height)
    ctx.free_accum(acc)
    assert sw.kernel_variant == capi.KERNEL_WAVEFRONT and sw.launches == 1 + 2 * depth
    assert sw.rays == sm.rays and sw.samples == sm.samples, (sw.rays, sm.rays)
    assert sw.node_visits > 0
    assert np.allclose(wave, mega, rtol=2e-5, atol=2e-5 * spp), float(np.abs(wave - mega).max())
    # and against the fp64 CPU twin, like every other kernel
    oimg, ost = sc.feed(oracle.scene()).render(cam, capi.make_params(**sc.params_args(sample_count=spp, max_depth=depth)))
    assert abs(int(sw.rays) - int(ost.rays)) <= max(4, 1e-2 * ost.rays)
    if sc.width >= 96:
        assert psnr(np.clip(wave / spp, 0, 1), np.clip(oimg / spp, 0, 1)) >= 40.0


def test_wavefront_batches_and_sample_ranges(ctx):
    """More samples than one queue batch holds, and disjoint sample ranges adding up."""
    sc = scenes.cornell_box(width=64, height=64)
    cam = capi.make_camera(**sc.camera_args())
    sc.feed(ctx)
    ctx.commit()
    wf = dict(integrator=scenes.INTEGRATOR_PATH_WAVEFRONT, traversal=2, max_depth=10)
    acc = ctx.alloc_accum(64, 64)
    for b in (0, 5):
        ctx.render_device(cam, capi.make_params(**sc.params_args(sample_begin=b, sample_count=5 if b == 0 else 7, **wf)), acc)
    two = ctx.download_accum(acc, 64, 64)
    ctx.free_accum(acc)
    one, st = ctx.render(cam, capi.make_params(**sc.params_args(sample_begin=0, sample_count=12, traversal=2, max_depth=10)))
    assert np.allclose(two, one, rtol=2e-5, atol=1e-4)
