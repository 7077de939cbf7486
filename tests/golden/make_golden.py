#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz from the REAL reference library.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
It builds oracle/_ref/libare_ref.so (oracle/Makefile: the reference's own sources compiled where they lie, plus
oracle/ref_harness.cpp) and records inputs + the reference's outputs for every library routine on the path-tracing
hot path.  The reference has no tests or golden vectors of its own, so
these fixtures — outputs of the reference itself — are what pins oracle/are_oracle.c, and through it the kernels.
The file is small (a few hundred kB) and committed; the GPU box never needs /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_binding import Reference, build_oracle  # noqa: E402
from aurora_rendering_engine_b200 import scenes  # noqa: E402


def main():
    build_oracle()
    ref = Reference()
    rng = np.random.RandomState(20261017)
    g = {}
    n = 512
    a, b = rng.uniform(-3, 3, (n, 3)), rng.uniform(-3, 3, (n, 3))
    s = rng.uniform(-2, 2, n)
    a[0] = 0.0          # normalize(0) -> NaN vector
    s[1] = 0.0          # v / 0 -> NaN vector
    a[2] = [1e-9, -1e-9, 5e-9]  # near_zero
    g["vec_a"], g["vec_b"], g["vec_s"] = a, b, s
    for op in range(8):
        g[f"vec_binary_{op}"] = ref.vec3_binary(op, a, b, s)
    for op in range(4):
        g[f"vec_scalar_{op}"] = ref.vec3_scalar(op, a, b)
    # reflect / refract, incl. eta > 1 at grazing incidence (the TIR case the reference does not detect)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    nn = rng.normal(size=(n, 3))
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    eta = rng.choice([1 / 1.5, 1.5, 1 / 1.33, 3.0, 1.0], n)
    g["rr_v"], g["rr_n"], g["rr_eta"] = v, nn, eta
    g["reflect"] = ref.reflect(v, nn)
    g["refract"] = ref.refract(v, nn, eta)
    # Ray ctor + at()
    Q, D, t = rng.uniform(-2, 2, (n, 3)), rng.uniform(-2, 2, (n, 3)), rng.uniform(-1, 5, n)
    D[0] = 0.0
    g["ray_Q"], g["ray_D"], g["ray_t"] = Q, D, t
    g["ray_outD"], g["ray_at"] = ref.ray(Q, D, t)
    # Plane
    pp, pn = rng.uniform(-2, 2, (n, 3)), rng.uniform(-2, 2, (n, 3))
    g["plane_p"], g["plane_n"] = pp, pn
    g["plane4"] = ref.plane_from_point_normal(pp, pn)
    D2 = rng.uniform(-1, 1, (n, 3))
    D2[:8] = np.cross(pn[:8], rng.uniform(-1, 1, (8, 3)))  # parallel to the plane -> |denom| < eps
    g["plane_rayQ"], g["plane_rayD"] = Q, D2
    g["plane_hit"], g["plane_P"] = ref.plane_intersect(g["plane4"], Q, D2)
    # Material::reflect (viewport-origin mirroring)
    planes = g["plane4"].copy()
    planes[:4, :3] = 0.0  # illegal plane -> false
    g["mat_planes"], g["mat_origin"] = planes, Q
    g["mat_refl_ok"], g["mat_refl_out"] = ref.material_reflect(1, 0.9, planes, Q)
    g["mat_diff_ok"], g["mat_diff_out"] = ref.material_reflect(0, 0.0, planes, Q)
    # Triangle ctor validation
    ctor_in = [((0, 0, 0), (1, 0, 0), (0, 1, 0), 0), ((0, 0, 0), (0, 0, 0), (0, 1, 0), 0), ((0, 0, 0), (1, 0, 0), (0, 0, 0), 0),
               ((0, 0, 0), (1, 0, 0), (2, 0, 0), 0), ((0, 0, 0), (1e-9, 0, 0), (0, 1, 0), 0), ((1, 2, 3), (1, 0, 0), (0, 1, 0), 1),
               ((1, 2, 3), (1, 0, 0), (0, 1, 0), 2), ((1, 2, 3), (1e-5, 0, 0), (0, 1e-5, 0), 0), ((5, 5, 5), (1, 1, 0), (1, 1, 1e-7), 0)]
    g["ctor_in"] = np.array([list(q) + list(u) + list(v_) + [f] for q, u, v_, f in ctor_in], float)
    res = [ref.triangle_ctor(q, u, v_, f) for q, u, v_, f in ctor_in]
    g["ctor_status"] = np.array([r[0] for r in res], np.int32)
    g["ctor_verts"] = np.stack([r[1] for r in res])
    # Triangle::intersect_ray / point_in on the 34-triangle rt.cpp Cornell set and on random triangles
    sc = scenes.rt_cornell()
    TQ = np.stack([t_[0] for t_ in sc.tris]); Tu = np.stack([t_[1] for t_ in sc.tris]); Tv = np.stack([t_[2] for t_ in sc.tris])
    g["cornell_TQ"], g["cornell_Tu"], g["cornell_Tv"] = TQ, Tu, Tv
    rq, rd = 0.3 * rng.uniform(-1, 1, (2000, 3)), rng.uniform(-1, 1, (2000, 3))
    # rays aimed exactly at vertices / along edges: the +-eps decisions
    rq[:34] = [0.1, 0.2, 0.3]
    rd[:34] = TQ - rq[:34]
    rq[34:68] = [0.0, 0.0, 0.5]
    rd[34:68] = (TQ + 0.5 * Tu) - rq[34:68]
    g["cornell_rayQ"], g["cornell_rayD"] = rq, rd
    ts = ref.triset(TQ, Tu, Tv)
    hit, P = ts.hit_matrix(rq[:256], rd[:256])
    g["cornell_hit_matrix"], g["cornell_hit_P"] = hit, P
    nh, prim, tt, PP = ts.closest_hit(rq, rd)
    g["cornell_closest_n"] = np.array([nh]); g["cornell_closest_prim"] = prim; g["cornell_closest_t"] = tt; g["cornell_closest_P"] = PP
    pts = np.concatenate([TQ[3] + rng.uniform(-0.2, 1.2, (200, 1)) * Tu[3] + rng.uniform(-0.2, 1.2, (200, 1)) * Tv[3],
                          np.array([TQ[3], TQ[3] + Tu[3], TQ[3] + Tv[3], TQ[3] + 0.5 * Tu[3] + 0.5 * Tv[3]])])
    g["pointin_pts"] = pts
    g["pointin"] = ts.point_in(3, pts)
    ts.close()
    # texture quantisation on save (values incl. out-of-range and exact byte boundaries) and load
    import tempfile
    tex = rng.uniform(-0.2, 1.3, (6, 9, 3))
    tex[0, 0] = [0.0, 1.0, 0.5]
    tex[0, 1] = [254.0 / 255.0, 255.0 / 255.0, 1.0 / 255.0]
    tex[0, 2] = [0.999999, 1e-9, 2.0]
    g["tex_rgb"] = tex
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.ppm")
        assert ref.texture_save(p, tex) == 1
        raw = open(p, "rb").read()
        g["tex_ppm_bytes"] = np.frombuffer(raw, np.uint8)
        st, back = ref.texture_load(p)
        assert st == 0
        g["tex_loaded"] = back
        g["tex_save_bad_suffix"] = np.array([ref.texture_save(os.path.join(d, "t.png"), tex)])
        g["tex_load_missing"] = np.array([ref.texture_load(os.path.join(d, "nope.ppm"))[0]])
        open(os.path.join(d, "p3.ppm"), "wb").write(b"P3\n1 1\n255\n0 0 0\n")
        g["tex_load_p3"] = np.array([ref.texture_load(os.path.join(d, "p3.ppm"))[0]])
        open(os.path.join(d, "short.ppm"), "wb").write(b"P6\n2 2\n255\n" + bytes(5))
        g["tex_load_short"] = np.array([ref.texture_load(os.path.join(d, "short.ppm"))[0]])
    # Texture::paste (src/texture.cpp:85-360): general quad, partly outside, degenerate, flipped, full-frame
    pdst, psrc = rng.uniform(0, 1, (40, 56, 3)), rng.uniform(0, 1, (9, 13, 3))
    pcorners = np.array([[3, 4, 50, 2, 2, 33, 52, 30], [-9, -6, 70, 5, 4, 50, 55, 39], [5, 5, 5, 5, 5, 20, 5, 20],
                         [40, 3, 6, 7, 45, 36, 2, 31], [0, 0, 55, 0, 0, 39, 55, 39], [200, 200, 260, 200, 200, 260, 260, 260]], np.int32)
    g["paste_dst"], g["paste_src"], g["paste_corners"] = pdst, psrc, pcorners
    g["paste_out"] = np.stack([ref.texture_paste(pdst, psrc, c.reshape(4, 2)) for c in pcorners])
    g["tex_fill_ctor"] = np.array([ref.texture_fill_ctor(4, 4), ref.texture_fill_ctor(0, 4), ref.texture_fill_ctor(4, -1)])
    g["sizeof"] = np.array(ref.sizeof())
    g["geometry_epsilon"] = np.array([ref.geometry_epsilon()])
    out = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
