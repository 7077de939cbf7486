#!/usr/bin/env python
"""Host binned-SAH build vs device LBVH build (csrc/lbvh.cu): build time and the render time of the tree each one makes.

  python tools/bench_lbvh.py [n_prims ...]      one JSON line per (scene size, builder)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aurora_rendering_engine_b200 import capi, scenes  # noqa: E402


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [100_000, 1_000_000]
    for n in sizes:
        sc = scenes.stress(n_prims=n, width=1920, height=1080) if n > 0 else scenes.rtiow_final(width=1200, height=675)
        cam = capi.make_camera(**sc.camera_args())
        par = capi.make_params(**sc.params_args(sample_count=2 if n > 0 else 16, traversal=2))
        for builder, label in ((capi.BVH_BUILDER_HOST_SAH, "host_sah"), (capi.BVH_BUILDER_DEVICE_LBVH, "device_lbvh")):
            with capi.Context(0) as ctx:
                ctx.set_bvh_builder(builder)
                sc.feed(ctx)
                commits = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    ctx.commit()
                    commits.append((time.perf_counter() - t0) * 1e3)
                info = ctx.commit_info()
                acc = ctx.alloc_accum(sc.width, sc.height)
                ms = []
                for k in range(4):
                    st = ctx.render_device(cam, par, acc, want_stats=True)
                    ms.append(st.kernel_ms)
                stc = ctx.render_device(cam, par, acc, want_stats=True, count_tests=True)
                ctx.free_accum(acc)
                print(json.dumps({"scene": sc.name, "prims": sc.num_prims, "builder": label, "used": info.builder, "commit_ms_best": min(commits),
                                  "host_compile_ms": info.host_compile_ms, "host_bvh_ms": info.host_bvh_ms, "device_bvh_ms": info.device_bvh_ms,
                                  "bvh_nodes": info.bvh_nodes, "bvh_height": info.bvh_height, "render_ms_best": min(ms[1:]),
                                  "msamples_per_s": st.samples / min(ms[1:]) / 1e3, "node_visits_per_ray": stc.node_visits / max(1, stc.rays),
                                  "prim_tests_per_ray": (stc.tri_tests + stc.quad_tests + stc.sphere_tests + stc.box_tests) / max(1, stc.rays)}), flush=True)


if __name__ == "__main__":
    main()
