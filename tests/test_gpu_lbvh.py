"""Device BVH builder (csrc/lbvh.cu) through the C ABI.

A BVH only prunes: whichever builder made it, a ray must meet the same primitive at the same distance, because the
leaves run the same per-primitive arithmetic.  So the device-built hierarchy (Morton codes + radix sort + Karras tree)
is checked against the host binned-SAH hierarchy and against brute force for IDENTICAL closest hits, and its images
against the host-built ones and the CPU twin.
"""
import numpy as np
import pytest

from aurora_rendering_engine_b200 import capi, scenes
from oracle_binding import psnr

pytestmark = pytest.mark.gpu


def _commit(ctx, sc, builder):
    ctx.clear()
    ctx.set_bvh_builder(builder)
    sc.feed(ctx)
    ctx.commit()
    return ctx.commit_info()


def test_lbvh_closest_hits_equal_host_sah_and_brute(ctx):
    sc = scenes.stress(n_prims=20_000, width=8, height=8)
    rng = np.random.RandomState(5)
    n = 300_000
    Q = rng.uniform(-12, 12, (n, 3))
    D = rng.normal(size=(n, 3))
    D[:1000, 0] = 0.0
    D[1000:2000, 1] = 0.0
    info_h = _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    ph, th, Ph, Nh, uvh = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
    info_d = _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
    pd, td, Pd, Nd, uvd = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
    assert info_h.builder == capi.BVH_BUILDER_HOST_SAH and info_d.builder == capi.BVH_BUILDER_DEVICE_LBVH
    assert info_d.bvh_nodes == info_h.bvh_nodes == info_d.hot_slots - 1  # one primitive per leaf, no boxes in this scene
    assert 0 < info_d.bvh_height <= 48 and info_d.device_bvh_ms > 0 and info_d.device_bvh_launches >= 5
    print(f"[lbvh 20k] host SAH {info_h.host_bvh_ms:.2f} ms (height {info_h.bvh_height}), device LBVH {info_d.device_bvh_ms:.3f} ms (height {info_d.bvh_height})")
    assert 0.01 < (ph >= 0).mean() < 0.99
    assert np.array_equal(ph, pd)
    hit = ph >= 0
    assert np.array_equal(th[hit], td[hit]) and np.array_equal(Ph[hit], Pd[hit]) and np.array_equal(uvh[hit], uvd[hit])


def test_lbvh_small_scene_equals_brute_force(ctx):
    """500 primitives: the brute-force list is the third witness."""
    sc = scenes.stress(n_prims=500, width=8, height=8)
    rng = np.random.RandomState(6)
    n = 100_000
    Q = rng.uniform(-4, 4, (n, 3))
    D = rng.normal(size=(n, 3))
    _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
    pb, tb, *_ = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=1)
    pd, td, *_ = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
    assert (pb >= 0).mean() > 0.01
    assert np.array_equal(pb, pd) and np.array_equal(tb[pb >= 0], td[pd >= 0])


@pytest.mark.parametrize("name,kw,spp", [
    ("cornell_box", dict(width=72, height=64), 8),     # boxes: two slots per leaf
    ("rtiow_final", dict(width=96, height=54), 4),     # one huge sphere among hundreds of small ones
    ("textured", dict(width=96, height=54), 8),
])
def test_lbvh_images_equal_host_sah_images(ctx, oracle, name, kw, spp):
    sc = scenes.by_name(name, **kw)
    cam = capi.make_camera(**sc.camera_args())
    par = capi.make_params(**sc.params_args(sample_count=spp, traversal=2, max_depth=12))
    _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    ih, sh = ctx.render(cam, par)
    info = _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
    idv, sd = ctx.render(cam, par)
    assert info.builder == capi.BVH_BUILDER_DEVICE_LBVH
    assert sh.rays == sd.rays, (sh.rays, sd.rays)
    # same paths; sliced traversals finish in a different order, so only the float summation order may differ
    assert np.allclose(ih, idv, rtol=1e-5, atol=1e-6), float(np.abs(ih - idv).max())
    osc = sc.feed(oracle.scene())
    oimg, ost = osc.render(cam, par)
    p = psnr(np.clip(idv / spp, 0, 1), np.clip(oimg / spp, 0, 1))
    assert p >= 40.0 and abs(int(sd.rays) - int(ost.rays)) < 1e-2 * ost.rays, (p, sd.rays, ost.rays)


def test_lbvh_edge_cases(ctx):
    """One primitive (the root is the leaf), two primitives, many primitives with one and the same centre (equal
    Morton codes are split on the index bits), an empty scene."""
    def scene(spheres):
        s = scenes.SceneDesc("edge", width=32, height=24, spp=4, max_depth=4)
        t = s.solid(0.7, 0.6, 0.5)
        m = s.mat(scenes.MAT_LAMBERTIAN, -1)
        for c, r in spheres:
            s.sphere(c, r, m, t)
        s.camera = dict(pos=(0, 0, 5), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=40.0, focus_dist=5.0, jitter=1)
        return s

    rng = np.random.RandomState(7)
    n = 20_000
    Q = np.tile(np.array([0.0, 0.0, 5.0]), (n, 1))
    D = np.concatenate([rng.uniform(-0.5, 0.5, (n, 2)), -np.ones((n, 1))], axis=1)
    cases = {
        "one": [((0, 0, 0), 1.0)],
        "two": [((-0.6, 0, 0), 0.5), ((0.6, 0, 0), 0.5)],
        "concentric": [((0, 0, 0), 0.2 + 0.01 * k) for k in range(100)],
        "coincident": [((0.3, 0.1, 0), 0.4)] * 64 + [((-0.8, 0, 0), 0.3)],
    }
    for label, sph in cases.items():
        sc = scene(sph)
        info_h = _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
        ph, th, *_ = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
        info_d = _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
        pd, td, *_ = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
        assert info_d.builder == capi.BVH_BUILDER_DEVICE_LBVH and info_d.bvh_nodes == len(sph) - 1, (label, info_d.as_dict())
        hit = ph >= 0
        assert hit.any() and np.array_equal(hit, pd >= 0), label
        assert np.array_equal(th[hit], td[hit]), label
        if label != "coincident":  # identical spheres: which of them reports the (identical) hit depends on the visiting order
            assert np.array_equal(ph, pd), label
    ctx.clear()
    ctx.set_bvh_builder(capi.BVH_BUILDER_DEVICE_LBVH)
    ctx.commit()
    p0, *_ = ctx.hit_batch(Q[:16], D[:16], t_min=1e-3, precision=32, traversal=2)
    assert (p0 < 0).all()


def test_lbvh_falls_back_to_host_builder_when_too_deep(ctx, monkeypatch):
    """A device-built tree deeper than the traversal stack is discarded and the host SAH builder takes over (the limit is
    lowered through the test hook ARE_OPT_LBVH_MAX_HEIGHT to provoke it); the scene renders as usual."""
    sc = scenes.stress(n_prims=2000, width=48, height=32)
    cam = capi.make_camera(**sc.camera_args())
    par = capi.make_params(**sc.params_args(sample_count=4, traversal=2))
    info = _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    ref, sr = ctx.render(cam, par)
    ctx.set_option(capi.OPT_LBVH_MAX_HEIGHT, 5)
    info = _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
    ctx.set_option(capi.OPT_LBVH_MAX_HEIGHT, 0)
    assert info.builder == capi.BVH_BUILDER_HOST_SAH and info.device_bvh_ms == 0 and info.bvh_nodes == 1999
    img, st = ctx.render(cam, par)
    assert st.rays == sr.rays and np.array_equal(img, ref)
    info = _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
    assert info.builder == capi.BVH_BUILDER_DEVICE_LBVH and info.bvh_height > 5


def test_refit_after_moving_primitives(ctx):
    """SURVEY §8f-2 refit: move a third of the primitives of a device-built scene, refit the hierarchy over its unchanged
    topology, and get exactly what a fresh commit of the moved scene gives — closest hits (fp32 traversal AND the fp64
    harness, i.e. the per-primitive arrays moved too), and the rendered image up to summation order."""
    sc = scenes.stress(n_prims=20_000, width=96, height=54)
    cam = capi.make_camera(**sc.camera_args())
    par = capi.make_params(**sc.params_args(sample_count=4, traversal=2))
    info = _commit(ctx, sc, capi.BVH_BUILDER_DEVICE_LBVH)
    assert info.builder == capi.BVH_BUILDER_DEVICE_LBVH
    before, _ = ctx.render(cam, par)
    rng = np.random.RandomState(3)
    ns, nt = len(sc.spheres), len(sc.tris)
    ms = rng.choice(ns, ns // 3, replace=False)          # stress(): spheres are primitives 0..ns-1, triangles ns..ns+nt-1
    mt = rng.choice(nt, nt // 3, replace=False)
    for i in ms:
        c, r, m, t = sc.spheres[i]
        sc.spheres[i] = (c + rng.normal(scale=0.3, size=3), r * rng.uniform(0.7, 1.4), m, t)
    for i in mt:
        Q, u, v, m, t, uv = sc.tris[i]
        sc.tris[i] = (Q + rng.normal(scale=0.3, size=3), u * rng.uniform(0.8, 1.3), v + 0.2 * u, m, t, uv)
    ctx.update_spheres(ms, np.stack([sc.spheres[i][0] for i in ms]), np.array([sc.spheres[i][1] for i in ms]))
    ctx.update_triangles(ns + mt, np.stack([sc.tris[i][0] for i in mt]), np.stack([sc.tris[i][1] for i in mt]), np.stack([sc.tris[i][2] for i in mt]))
    dev_ms = ctx.refit()
    assert 0 < dev_ms < 50
    img, st = ctx.render(cam, par)
    assert not np.array_equal(img, before)
    Q = rng.uniform(-12, 12, (100_000, 3))
    D = rng.normal(size=(100_000, 3))
    got32 = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
    got64 = ctx.hit_batch(Q, D, t_min=1e-3, precision=64)
    with capi.Context(0) as fresh:                        # the moved scene committed from scratch
        fresh.set_bvh_builder(capi.BVH_BUILDER_DEVICE_LBVH)
        sc.feed(fresh)
        fresh.commit()
        ref, sr = fresh.render(cam, par)
        want32 = fresh.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
        want64 = fresh.hit_batch(Q, D, t_min=1e-3, precision=64)
    for a, b in zip(got32, want32):
        assert np.array_equal(a, b, equal_nan=True)
    for a, b in zip(got64, want64):
        assert np.array_equal(a, b, equal_nan=True)
    assert st.rays == sr.rays and np.allclose(img, ref, rtol=1e-5, atol=1e-5)
    # a second refit on top of the first (dirty list re-armed), then moving things back gives the first image again
    ctx.update_spheres(ms[:10], np.stack([sc.spheres[i][0] for i in ms[:10]]) + 1.0, np.array([sc.spheres[i][1] for i in ms[:10]]))
    ctx.refit()
    ctx.update_spheres(ms[:10], np.stack([sc.spheres[i][0] for i in ms[:10]]), np.array([sc.spheres[i][1] for i in ms[:10]]))
    ctx.refit()
    again, _ = ctx.render(cam, par)
    assert np.allclose(again, ref, rtol=1e-5, atol=1e-5)


def test_refit_refuses_what_it_cannot_do(ctx):
    """Host-built hierarchies and scenes with fused items (the Cornell box: parallelograms, boxes) are not refittable: the
    call says so and the caller commits; wrong ids / types / degenerate geometry are rejected at update time."""
    sc = scenes.stress(n_prims=2000, width=32, height=32)
    _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    ctx.update_spheres([0], [[0.0, 0.0, 0.0]], [0.05])
    with pytest.raises(capi.AreCudaError, match="commit instead"):
        ctx.refit()
    ctx.commit()   # the update is kept: a commit picks it up
    with pytest.raises(capi.AreCudaError):
        ctx.update_spheres([len(sc.spheres)], [[0.0, 0.0, 0.0]], [0.05])      # that id is a triangle
    with pytest.raises(capi.AreCudaError):
        ctx.update_spheres([0], [[0.0, 0.0, 0.0]], [-1.0])
    with pytest.raises(capi.AreCudaError):
        ctx.update_triangles([len(sc.spheres)], [[0.0, 0.0, 0.0]], [[1.0, 0.0, 0.0]], [[2.0, 0.0, 0.0]])  # collinear edges
    box = scenes.cornell_box(width=32, height=32)
    _commit(ctx, box, capi.BVH_BUILDER_DEVICE_LBVH)
    Q, u, v, *_ = box.tris[0]
    ctx.update_triangles([0], [Q + 1.0], [u], [v])
    with pytest.raises(capi.AreCudaError, match="commit instead"):
        ctx.refit()


def test_bvh4_equals_bvh2(ctx):
    """The uncompressed 4-wide hierarchy (ARE_OPT_BUILD_BVH4 + ARE_TRAVERSAL_BVH4) is a collapse of the host-built BVH2 whose
    child boxes are the BVH2's own floats: a ray must meet the same primitive at the same distance, and the images agree up
    to summation order.  Without the option (or on a device-built tree) the request falls back to the BVH2."""
    sc = scenes.stress(n_prims=20_000, width=96, height=54)
    cam = capi.make_camera(**sc.camera_args())
    rng = np.random.RandomState(7)
    Q = rng.uniform(-12, 12, (200_000, 3))
    D = rng.normal(size=(200_000, 3))
    D[:1000, 0] = 0.0
    ctx.set_option(capi.OPT_BUILD_BVH4, 1)
    _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    want = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=2)
    got = ctx.hit_batch(Q, D, t_min=1e-3, precision=32, traversal=4)
    for a, b in zip(got, want):
        assert np.array_equal(a, b, equal_nan=True)
    i2, s2 = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=4, traversal=2)))
    i4, s4 = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=4, traversal=4)))
    acc = ctx.alloc_accum(sc.width, sc.height)
    c4 = ctx.render_device(cam, capi.make_params(**sc.params_args(sample_count=4, traversal=4)), acc, want_stats=True, count_tests=True)
    ctx.free_accum(acc)
    assert s4.kernel_variant == capi.KERNEL_BVH4 and s2.kernel_variant in (capi.KERNEL_BVH2, capi.KERNEL_BVH2_BIG, capi.KERNEL_BVH2_QUANT)
    assert s4.rays == s2.rays and np.allclose(i4, i2, rtol=1e-5, atol=1e-5)
    assert c4.node_visits > 0 and c4.sphere_tests > 0
    for name, kw in (("rtiow_final", dict(width=96, height=54)), ("cornell_box", dict(width=64, height=64))):
        s = scenes.by_name(name, **kw)
        _commit(ctx, s, capi.BVH_BUILDER_HOST_SAH)
        c = capi.make_camera(**s.camera_args())
        a2, t2 = ctx.render(c, capi.make_params(**s.params_args(sample_count=4, traversal=2, max_depth=12)))
        a4, t4 = ctx.render(c, capi.make_params(**s.params_args(sample_count=4, traversal=4, max_depth=12)))
        assert t4.kernel_variant == capi.KERNEL_BVH4 and t4.rays == t2.rays and np.allclose(a4, a2, rtol=1e-5, atol=1e-5)
    # not built: the request is served by the BVH2
    ctx.set_option(capi.OPT_BUILD_BVH4, 0)
    _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    _, sf = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=1, traversal=4)))
    assert sf.kernel_variant in (capi.KERNEL_BVH2, capi.KERNEL_BVH2_BIG, capi.KERNEL_BVH2_QUANT)


def test_quantised_nodes_equal_fp32_nodes(ctx):
    """Hierarchies beyond 8192 nodes are traversed through 32-byte nodes whose child boxes are 16-bit planes on a grid
    over the root box, padded outwards (ARE_OPT_QUANTIZED_NODES, dev_types.h: BvhNodeQ): a superset of every fp32 box, so
    each ray must end on the same primitive — identical ray counts, images equal up to summation order — at the price of
    a few more node visits.  Both builders; switching the option off renders through the fp32 nodes at once; a camera 150
    scene extents away sees the same picture through either."""
    sc = scenes.stress(n_prims=40_000, width=192, height=108)
    cam = capi.make_camera(**sc.camera_args())
    par = capi.make_params(**sc.params_args(sample_count=8, traversal=2))
    for builder in (capi.BVH_BUILDER_HOST_SAH, capi.BVH_BUILDER_DEVICE_LBVH):
        info = _commit(ctx, sc, builder)
        assert info.bvh_nodes > 16384 and 1000 <= info.quant_area_permille < 1100
        out = {}
        for q in (1, 0):
            ctx.set_option(capi.OPT_QUANTIZED_NODES, q)
            img, st = ctx.render(cam, par)
            acc = ctx.alloc_accum(sc.width, sc.height)
            cnt = ctx.render_device(cam, par, acc, want_stats=True, count_tests=True)
            ctx.free_accum(acc)
            out[q] = (img, st, cnt)
        ctx.set_option(capi.OPT_QUANTIZED_NODES, 1)
        (iq, sq, cq), (i32, s32, c32) = out[1], out[0]
        assert sq.kernel_variant == capi.KERNEL_BVH2_QUANT and s32.kernel_variant == capi.KERNEL_BVH2_BIG
        assert sq.rays == s32.rays == cq.rays == c32.rays
        assert np.allclose(iq, i32, rtol=1e-5, atol=1e-5)
        more = cq.node_visits / c32.node_visits
        print(f"[quantised nodes, builder {builder}] node visits x{more:.4f}, primitive tests x"
              f"{(cq.sphere_tests + cq.tri_tests) / (c32.sphere_tests + c32.tri_tests):.4f}")
        assert 0.999 <= more < 1.10
        assert cq.sphere_tests + cq.tri_tests >= 0.999 * (c32.sphere_tests + c32.tri_tests)
    # a ragged frame, a camera inside the cloud looking along an axis (zero direction components on the centre column)
    inside = capi.make_camera(pos=(0.25, 0.5, 0.125), target=(0.25, 0.5, -30.0), up=(0, 1, 0), vfov_deg=60.0, jitter=0)
    par2 = capi.make_params(**sc.params_args(width=61, height=35, sample_count=3, traversal=2))
    a, sa = ctx.render(inside, par2)
    ctx.set_option(capi.OPT_QUANTIZED_NODES, 0)
    b, sb = ctx.render(inside, par2)
    ctx.set_option(capi.OPT_QUANTIZED_NODES, 1)
    assert sa.kernel_variant == capi.KERNEL_BVH2_QUANT and sa.rays == sb.rays and np.allclose(a, b, rtol=1e-5, atol=1e-5)
    # far away (150 scene extents): the decode's B' is then rounded at the ulp of the distance, as the fp32 step's o / d is
    far = capi.make_camera(pos=(0, 0, 5000.0), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=0.5)
    a, sa = ctx.render(far, par)
    ctx.set_option(capi.OPT_QUANTIZED_NODES, 0)
    b, sb = ctx.render(far, par)
    ctx.set_option(capi.OPT_QUANTIZED_NODES, 1)
    assert sa.kernel_variant == capi.KERNEL_BVH2_QUANT and sb.kernel_variant == capi.KERNEL_BVH2_BIG
    # at this distance fp32 places a hit to ~5e-4, the size of the smallest primitives' features: grazing decisions may differ
    assert abs(sa.rays - sb.rays) <= 2e-3 * sb.rays and psnr(np.clip(a / 8, 0, 1), np.clip(b / 8, 0, 1)) > 35.0
    # a scene whose extent dwarfs its primitives: the grid is too coarse, the commit says so and the fp32 nodes render
    lam, grey = sc.spheres[0][2], sc.spheres[0][3]
    sc.spheres.append((np.array([0.0, -1e5 - 20.0, 0.0]), 1e5, lam, grey))
    sc.order.append(("s", len(sc.spheres) - 1))
    info = _commit(ctx, sc, capi.BVH_BUILDER_HOST_SAH)
    assert info.quant_area_permille > 1250
    _, sg = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=1, traversal=2)))
    assert sg.kernel_variant == capi.KERNEL_BVH2_BIG


@pytest.mark.parametrize("case", ["flat", "far_from_origin", "one_huge_axis"])
def test_quantised_nodes_on_awkward_extents(ctx, case):
    """The quantisation grid is derived from the root box: a scene with NO extent along an axis (coplanar triangles), a scene
    10^4 units from the origin (coordinates that fp32 resolves to 1e-3 only), and a scene 1000 times longer along one axis
    than along the others must all render what their fp32 nodes render."""
    rng = np.random.RandomState(5)
    n = 12_000
    s = scenes.SceneDesc(case, width=160, height=90, spp=4, max_depth=6, t_min=1e-3)
    lam = s.mat(scenes.MAT_LAMBERTIAN, -1)
    grey = s.solid(0.6, 0.5, 0.4)
    if case == "flat":  # small triangles in the plane z = 0 (disjoint cells of a grid, so no two are coplanar AND overlapping)
        k = int(np.ceil(np.sqrt(n)))
        for i in range(n):
            x, y = (i % k) * 0.1 - 0.05 * k, (i // k) * 0.1 - 0.05 * k
            s.tri((x, y, 0.0), (0.08, 0.0, 0.0), (0.0, 0.08, 0.0), lam, grey)
        cam = dict(pos=(0.3, -0.2, 9.0), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=50.0)
    else:
        off = np.array([1e4, -2e4, 1.5e4]) if case == "far_from_origin" else np.zeros(3)
        scale = np.array([1.0, 1.0, 1.0]) if case == "far_from_origin" else np.array([1000.0, 1.0, 1.0])
        c = rng.uniform(-5, 5, (n, 3)) * scale + off
        r = rng.uniform(0.03, 0.08, n)
        for i in range(n):
            s.sphere(c[i], float(r[i]), lam, grey)
        cam = dict(pos=tuple(off + np.array([0.0, 0.0, 16.0])), target=tuple(off), up=(0, 1, 0), vfov_deg=45.0)
    s.camera = dict(cam, focus_dist=10.0, jitter=1)
    camera = capi.make_camera(**s.camera_args())
    par = capi.make_params(**s.params_args(sample_count=4, traversal=2))
    for builder in (capi.BVH_BUILDER_HOST_SAH, capi.BVH_BUILDER_DEVICE_LBVH):
        info = _commit(ctx, s, builder)
        assert info.bvh_nodes > 8192 and info.quant_area_permille >= 1000, info.as_dict()
        # 1000 : 1 extents put 0.06-wide spheres on a 0.15-wide grid along x: the commit measures that and keeps the fp32 nodes
        used = info.quant_area_permille <= 1250
        assert used == (case != "one_huge_axis"), info.as_dict()
        iq, sq = ctx.render(camera, par)
        ctx.set_option(capi.OPT_QUANTIZED_NODES, 0)
        i32, s32 = ctx.render(camera, par)
        ctx.set_option(capi.OPT_QUANTIZED_NODES, 1)
        assert sq.kernel_variant == (capi.KERNEL_BVH2_QUANT if used else capi.KERNEL_BVH2_BIG) and s32.kernel_variant == capi.KERNEL_BVH2_BIG
        assert sq.rays == s32.rays and sq.rays > s.width * s.height * 4
        assert np.allclose(iq, i32, rtol=1e-5, atol=1e-5)
