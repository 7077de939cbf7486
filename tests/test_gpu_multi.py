"""Multi-device groups of the C ABI (are_cuda_create_multi): one context, several GPUs, samples sharded by range and
summed on the first device over peer memory.  A group of ONE device must be the single-device path bit for bit (runs on
any box); groups of 2+ devices need as many GPUs and are skipped below that."""
import numpy as np
import pytest

from aurora_rendering_engine_b200 import capi, scenes
from oracle_binding import psnr

pytestmark = pytest.mark.gpu


def _render(ctx, sc, spp, traversal=0, **over):
    cam = capi.make_camera(**sc.camera_args())
    ctx.clear()
    sc.feed(ctx)
    ctx.commit()
    return ctx.render(cam, capi.make_params(**sc.params_args(sample_count=spp, traversal=traversal, **over)))


def test_group_of_one_is_the_single_device_path(lib):
    sc = scenes.cornell_box(width=96, height=80)
    with capi.Context(0) as one, capi.Context([0]) as grp:
        assert grp.group_info() == (1, False)
        a, sa = _render(one, sc, 12)
        b, sb = _render(grp, sc, 12)
    assert sa.rays == sb.rays and np.array_equal(a, b)


def test_group_rejects_bad_device_lists(lib):
    n = lib.are_cuda_device_count()
    for devs in ([0, 0], [n], [-1], []):
        with pytest.raises(capi.AreCudaError):
            capi.Context(devs)


@pytest.mark.parametrize("name,kw,spp,traversal", [
    ("cornell_box", dict(width=160, height=128), 37, 0),      # odd count: shares of different sizes
    ("rtiow_final", dict(width=160, height=90), 8, 2),
    ("cornell_box", dict(width=33, height=17), 3, 0),         # fewer samples than devices on big boxes; frame not a multiple of 4 floats
])
def test_group_equals_one_device(lib, name, kw, spp, traversal):
    """the N-device image is the 1-device image up to float summation order (>= 60 dB; here ~1e-6 relative),
    the ray count is identical — the group draws exactly the samples one device would."""
    n = lib.are_cuda_device_count()
    if n < 2:
        pytest.skip("needs 2+ GPUs")
    sc = scenes.by_name(name, **kw)
    with capi.Context(0) as one:
        ref, sr = _render(one, sc, spp, traversal, max_depth=12)
    for k in sorted({2, n}):
        with capi.Context(list(range(k))) as grp:
            assert grp.group_info()[0] == k
            img, st = _render(grp, sc, spp, traversal, max_depth=12)
            # a second render into the same context: scratch buffers are reused, events re-armed
            img2, st2 = grp.render(capi.make_camera(**sc.camera_args()), capi.make_params(**sc.params_args(sample_count=spp, traversal=traversal, max_depth=12)))
        assert st.rays == sr.rays == st2.rays, (k, st.rays, sr.rays)
        assert st.samples == sr.samples
        p = psnr(np.clip(img / spp, 0, 1), np.clip(ref / spp, 0, 1))
        assert p >= 60.0, (k, p)
        assert np.allclose(img, ref, rtol=1e-5, atol=1e-5 * spp) and np.array_equal(img, img2)


def test_group_render_device_accumulates(lib):
    """are_cuda_render_device on a group ADDS the group's sum into the caller's accumulator on the first device."""
    n = lib.are_cuda_device_count()
    if n < 2:
        pytest.skip("needs 2+ GPUs")
    sc = scenes.cornell_box(width=64, height=64)
    cam = capi.make_camera(**sc.camera_args())
    with capi.Context(list(range(n))) as grp, capi.Context(0) as one:
        for c in (grp, one):
            sc.feed(c)
            c.commit()
        acc = grp.alloc_accum(64, 64)
        for b in (0, 8):
            grp.render_device(cam, capi.make_params(**sc.params_args(sample_begin=b, sample_count=8)), acc)
        two = grp.download_accum(acc, 64, 64)
        grp.free_accum(acc)
        ref, _ = one.render(cam, capi.make_params(**sc.params_args(sample_begin=0, sample_count=16)))
    assert np.allclose(two, ref, rtol=1e-5, atol=1e-4)
