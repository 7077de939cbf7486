#!/usr/bin/env python
"""tools/bench_patch.py — patch-as-viewport renderer (experiments/rt10.cpp algorithm): GPU path vs the CPU checkers.

Prints one JSON line per workload: end-to-end wall time of are_cuda_patch_render with host buffers (plan + upload +
kernels + download), its device time (CUDA events), the CPU restatement (oracle/libpatch_oracle.so) and — where
oracle/_ref/librt10_ref.so was built — the real reference, all on the same box, outputs compared byte for byte."""
import dataclasses
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aurora_rendering_engine_b200 import capi, scenes  # noqa: E402
from oracle_binding import PatchOracle, PatchReference  # noqa: E402


def best(f, n):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        r = f()
        ts.append(time.perf_counter() - t)
    return min(ts), r


def main():
    orc = PatchOracle()
    ref = PatchReference() if PatchReference.available() else None
    cases = [("rt10 900x650 (the reference program's own configuration)", scenes.patch_rt10()),
             ("rt10 3840x2160, node textures up to 1024^2", dataclasses.replace(scenes.patch_rt10(width=3840, height=2160), max_tex_res=1024)),
             ("rt10 7680x4320, node textures up to 2048^2, depth 6", dataclasses.replace(scenes.patch_rt10(width=7680, height=4320, max_depth=6), max_tex_res=2048))]
    with capi.Context(0) as ctx:
        for name, ps in cases:
            for _ in range(3):
                ctx.patch_render(ps)
            t_gpu, (rgb, rgb8, st) = best(lambda: ctx.patch_render(ps), 5)
            t_gpu8, _ = best(lambda: ctx.patch_render(ps, want_rgb=False), 5)
            t_orc, (orgb, orgb8) = best(lambda: orc.render(ps), 2)
            line = {"workload": name, "gpu_e2e_ms": t_gpu * 1e3, "gpu_e2e_rgb8_only_ms": t_gpu8 * 1e3, "gpu_kernel_ms": st.kernel_ms, "gpu_plan_ms": st.plan_ms,
                    "launches": st.launches, "nodes": st.nodes, "node_texels": st.node_texels, "warp_triangles": st.ops,
                    "pixels": ps.width * ps.height, "cpu_port_ms": t_orc * 1e3, "identical_to_port": bool(np.array_equal(rgb, orgb) and np.array_equal(rgb8, orgb8))}
            if ref is not None:
                t_ref, (rrgb, rrgb8) = best(lambda: ref.render(ps), 2)
                line.update(cpu_reference_ms=t_ref * 1e3, identical_to_reference=bool(np.array_equal(rgb, rrgb) and np.array_equal(rgb8, rrgb8)))
            line["speedup_e2e_vs_port"] = line["cpu_port_ms"] / line["gpu_e2e_ms"]
            print(json.dumps(line))


if __name__ == "__main__":
    main()
