#!/bin/bash
# tools/sanitize.sh — compute-sanitizer (memcheck, racecheck, initcheck) over one small invocation of every kernel family:
# the megakernel (lean and generic brute force) + fp64 harness, a BVH2 / wide-BVH render, the device BVH builder, the RT_AO integrator, Texture::paste and
# the patch renderer.  Writes gpurun_out/sanitizer_<tool>.log; exit status 0 only if every tool reports 0 errors.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/are_sanitize_workload.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from aurora_rendering_engine_b200 import capi, scenes
with capi.Context(0) as ctx:
    for name, kw, trav in (("cornell_box", dict(width=48, height=48), 0), ("rtiow_final", dict(width=48, height=27), 2), ("rtiow_final", dict(width=48, height=27), 3),
                           ("textured", dict(width=48, height=27), 0), ("rt_cornell", dict(width=48, height=48), 0), ("stress", dict(n_prims=30000, width=48, height=27), 0)):
        sc = scenes.by_name(name, **kw)
        ctx.clear(); sc.feed(ctx); ctx.commit()
        img, st = ctx.render(capi.make_camera(**sc.camera_args()), capi.make_params(**sc.params_args(sample_count=2, traversal=trav, max_depth=min(sc.max_depth, 8))))
        assert np.isfinite(img).all()
        Q = np.random.RandomState(0).uniform(-1, 1, (2000, 3)); D = np.random.RandomState(1).normal(size=(2000, 3))
        ctx.hit_batch(Q, D, precision=64); ctx.hit_batch(Q, D, precision=32, traversal=trav)
    # the generic brute-force kernel on a scene that has a lean form, and the device BVH builder (lbvh.cu) + a render of its tree
    sc = scenes.by_name("cornell_box", width=48, height=48)
    ctx.clear(); sc.feed(ctx); ctx.commit()
    for opts, want in (((1, 1), capi.KERNEL_BRUTE_BAKED), ((1, 0), capi.KERNEL_BRUTE_LEAN), ((0, 0), capi.KERNEL_BRUTE)):
        ctx.set_option(capi.OPT_LEAN_KERNEL, opts[0]); ctx.set_option(capi.OPT_BAKED_KERNEL, opts[1])
        img, st = ctx.render(capi.make_camera(**sc.camera_args()), capi.make_params(**sc.params_args(sample_count=2, traversal=1, max_depth=8)))
        assert st.kernel_variant == want, (st.kernel_variant, want)
    ctx.set_option(capi.OPT_LEAN_KERNEL, 1); ctx.set_option(capi.OPT_BAKED_KERNEL, 1)
    ctx.set_bvh_builder(capi.BVH_BUILDER_DEVICE_LBVH)
    for name, kw in (("cornell_box", dict(width=48, height=48)), ("stress", dict(n_prims=30000, width=48, height=27))):
        sc = scenes.by_name(name, **kw)
        ctx.clear(); sc.feed(ctx); ctx.commit()
        assert ctx.commit_info().builder == capi.BVH_BUILDER_DEVICE_LBVH
        img, st = ctx.render(capi.make_camera(**sc.camera_args()), capi.make_params(**sc.params_args(sample_count=2, traversal=2, max_depth=8)))
        assert np.isfinite(img).all()
        if name == "stress":  # 30 000 nodes: the quantised 32-byte nodes (k_quant_grid, k_quant_nodes, bvhq_step), then the fp32 nodes (256-bit loads)
            assert st.kernel_variant == capi.KERNEL_BVH2_QUANT, st.kernel_variant
            ctx.set_option(capi.OPT_QUANTIZED_NODES, 0)
            img, st = ctx.render(capi.make_camera(**sc.camera_args()), capi.make_params(**sc.params_args(sample_count=2, traversal=2, max_depth=8)))
            assert st.kernel_variant == capi.KERNEL_BVH2_BIG and np.isfinite(img).all()
            ctx.set_option(capi.OPT_QUANTIZED_NODES, 1)
    ctx.set_bvh_builder(capi.BVH_BUILDER_HOST_SAH)
    # round 2: wavefront schedule, 4-wide BVH, refit, tone-map, L2 probe
    ctx.set_option(capi.OPT_BUILD_BVH4, 1)
    sc = scenes.by_name("stress", n_prims=3000, width=48, height=27)
    ctx.clear(); sc.feed(ctx); ctx.commit()
    cam = capi.make_camera(**sc.camera_args())
    for kw in (dict(traversal=4), dict(traversal=2, integrator=scenes.INTEGRATOR_PATH_WAVEFRONT)):
        img, st = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=2, max_depth=6, **kw)))
        assert np.isfinite(img).all() and st.rays > 0
    ctx.set_option(capi.OPT_BUILD_BVH4, 0)
    ctx.set_bvh_builder(capi.BVH_BUILDER_DEVICE_LBVH)
    ctx.clear(); sc.feed(ctx); ctx.commit()
    ctx.update_spheres([0, 5, 9], [[0.1, 0.2, 0.3]] * 3, [0.04, 0.05, 0.03])
    Q, u, v, *_ = sc.tris[0]
    ctx.update_triangles([len(sc.spheres)], [Q + 0.5], [u], [v])
    ctx.refit()
    img, st = ctx.render(cam, capi.make_params(**sc.params_args(sample_count=2, traversal=2, max_depth=6)))
    acc = ctx.alloc_accum(48, 27)
    ctx.render_device(cam, capi.make_params(**sc.params_args(sample_count=2, traversal=2, max_depth=6)), acc)
    ctx.tonemap(acc, 48, 27, 0.5, 0); ctx.free_accum(acc)
    ctx.set_bvh_builder(capi.BVH_BUILDER_HOST_SAH)
    ctx.measure_l2_peak()
    with capi.Context([0]) as grp:   # a group of one device: the multi-device driver code without a second GPU
        sc2 = scenes.by_name("cornell_box", width=48, height=48)
        sc2.feed(grp); grp.commit()
        grp.render(capi.make_camera(**sc2.camera_args()), capi.make_params(**sc2.params_args(sample_count=3, max_depth=6)))
    ps = scenes.patch_random(1, width=64, height=48, mirror_walls=True)
    ctx.patch_render(ps); ctx.patch_trace_texture(ps, ps.origin, 12, 40, 40)
    dst = np.zeros((40, 50, 3)); src = np.random.RandomState(2).uniform(size=(20, 30, 3))
    ctx.texture_paste(dst, src, [(5, 5), (40, 8), (8, 30), (45, 35)])
print("workload ok")
PY
rc=0
for tool in memcheck racecheck initcheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/are_sanitize_workload.py > gpurun_out/sanitizer_$tool.log 2>&1 || rc=1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1) $(grep -c 'workload ok' gpurun_out/sanitizer_$tool.log) run(s) completed"
done
exit $rc
