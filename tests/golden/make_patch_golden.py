#!/usr/bin/env python
"""Generate tests/golden/patch_vectors.npz from the REAL reference patch renderer (experiments/rt10.cpp).

Run in the build container (needs /root/reference):   python tests/golden/make_patch_golden.py
oracle/Makefile compiles the reference translation unit where it lies (behind oracle/ref_patch_harness.cpp) into
oracle/_ref/librt10_ref.so; this script records what IT computes:

  rt10_sha256 / rt10_size   sha256 of the P6 payload the reference renders for its built-in scene — and the script
                            asserts that payload equals the reference's shipped experiments/output_rt10.ppm;
  rt10_rows                 64 full rows of that payload (so a mismatch can be located without the reference);
  rand{k}_rgb / _rgb8       linear fp64 image + P6 payload of scenes.patch_random(k) (mirror walls on odd k);
  rand{k}_tex{j}            renderTriangleWithTriangle textures of a few triangles of those scenes.

The reference has no tests of its own; outputs of the reference itself are the fixtures.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_binding import PatchReference, build_oracle  # noqa: E402
from aurora_rendering_engine_b200 import scenes  # noqa: E402

N_RANDOM = 6
TEX_TRIS = (10, 12, 17, 23)


def random_scene(k):
    return scenes.patch_random(k, width=120, height=90, mirror_walls=(k % 2 == 1))


def main():
    build_oracle()
    ref = PatchReference()
    g = {}
    ps = scenes.patch_rt10()
    rgb, rgb8 = ref.render(ps)
    shipped = open("/root/reference/experiments/output_rt10.ppm", "rb").read()
    header = b"P6\n900 650\n255\n"
    assert shipped[: len(header)] == header and shipped[len(header):] == rgb8.tobytes(), "reference build does not reproduce its shipped image"
    g["rt10_sha256"] = np.frombuffer(hashlib.sha256(rgb8.tobytes()).digest(), dtype=np.uint8)
    g["rt10_size"] = np.array([900, 650])
    rows = np.arange(5, 650, 10)[:64]
    g["rt10_row_index"] = rows
    g["rt10_rows"] = rgb8[rows]
    g["rt10_rgb_rows"] = rgb[rows[::8]]
    for k in range(N_RANDOM):
        r = random_scene(k)
        a, b = ref.render(r)
        g[f"rand{k}_rgb"], g[f"rand{k}_rgb8"] = a, b
        for j, cur in enumerate(TEX_TRIS):
            g[f"rand{k}_tex{j}"] = ref.trace_texture(r, r.origin, cur, 40 + 30 * j, 300 - 60 * j, 0.0)
    out = os.path.join(HERE, "patch_vectors.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
