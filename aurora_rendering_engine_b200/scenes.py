"""Synthetic scene generators for the BASELINE.json configurations.

A :class:`SceneDesc` is a neutral, numpy-only description (textures, materials, primitives, camera, render
defaults).  ``SceneDesc.feed(sink)`` replays it into anything that exposes the ``add_*`` vocabulary of
``include/are_cuda.h`` — the CUDA context (:class:`aurora_rendering_engine_b200.capi.Context`) in the product,
or the CPU oracle wrapper in ``tests/``.  Nothing here computes a pixel.

Scenes
------
``rt_cornell``      config 0 — the 34-triangle mirror Cornell box of the reference's only stochastic renderer
                    (/root/reference/experiments/rt.cpp:153-185), camera (0,0,4)->(0,0,0), fov 50 (rt.cpp:399).
``rtiow_final``     config 1 — "Ray Tracing in One Weekend" final scene, ~485 spheres, 1200x675.
``textured``        config 2 — spatial checker ground, Perlin sphere, image sphere, emissive quads, 1920x1080.
``cornell_box``     config 3 — 555-unit Cornell box (walls + light + two rotated boxes as triangles), 2048x2048.
``stress``          config 4 — n/2 spheres + n/2 triangles in a 100^3 cube (n = 1 M in BASELINE), 3840x2160.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

MAT_DIFFUSE, MAT_REFLECTIVE, MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT = range(6)
TEX_SOLID, TEX_CHECKER_UV, TEX_CHECKER_3D, TEX_NOISE, TEX_IMAGE = range(5)
INTEGRATOR_PATH, INTEGRATOR_RT_AO, INTEGRATOR_PATH_WAVEFRONT = 0, 1, 2


def _p8(*vals):
    p = np.zeros(8, dtype=np.float64)
    p[: len(vals)] = vals
    return p


@dataclass
class SceneDesc:
    name: str
    textures: list = field(default_factory=list)   # (kind, params[8], rgb ndarray (h,w,3) float64 | None)
    materials: list = field(default_factory=list)  # (kind, params[8])
    tris: list = field(default_factory=list)       # (Q, u, v, mat, tex, uv6 | None)   -> prim ids in add order
    quads: list = field(default_factory=list)      # (Q, u, v, mat, tex)
    spheres: list = field(default_factory=list)    # (c, r, mat, tex)
    order: list = field(default_factory=list)      # ('t'|'q'|'s', index) in add order (defines primitive ids)
    camera: dict = field(default_factory=dict)
    width: int = 512
    height: int = 512
    spp: int = 64
    max_depth: int = 50
    t_min: float = 1e-3
    background_bottom: tuple = (1.0, 1.0, 1.0)
    background_top: tuple = (0.5, 0.7, 1.0)
    integrator: int = INTEGRATOR_PATH
    ao_samples: int = 32

    # -- construction helpers -------------------------------------------------------------------------
    def tex(self, kind, *params, rgb=None):
        self.textures.append((kind, _p8(*params), rgb))
        return len(self.textures) - 1

    def solid(self, r, g, b):
        return self.tex(TEX_SOLID, r, g, b)

    def mat(self, kind, *params):
        self.materials.append((kind, _p8(*params)))
        return len(self.materials) - 1

    def tri(self, Q, u, v, mat, tex, uv=None):
        self.tris.append((np.asarray(Q, float), np.asarray(u, float), np.asarray(v, float), mat, tex,
                          None if uv is None else np.asarray(uv, float)))
        self.order.append(("t", len(self.tris) - 1))

    def quad(self, Q, u, v, mat, tex):
        self.quads.append((np.asarray(Q, float), np.asarray(u, float), np.asarray(v, float), mat, tex))
        self.order.append(("q", len(self.quads) - 1))

    def sphere(self, c, r, mat, tex):
        self.spheres.append((np.asarray(c, float), float(r), mat, tex))
        self.order.append(("s", len(self.spheres) - 1))

    def quad_as_tris(self, p0, p1, p2, p3, mat, tex):
        """Two triangles (p0,p1,p2), (p0,p2,p3) with the uv assignment of rt.cpp pushQuad (:168-175)."""
        p0, p1, p2, p3 = (np.asarray(p, float) for p in (p0, p1, p2, p3))
        self.tri(p0, p1 - p0, p2 - p0, mat, tex, uv=(0, 0, 1, 0, 1, 1))
        self.tri(p0, p2 - p0, p3 - p0, mat, tex, uv=(0, 0, 1, 1, 0, 1))

    @property
    def num_prims(self):
        return len(self.order)

    # -- replay ---------------------------------------------------------------------------------------
    def feed(self, sink, bulk_threshold=4096):
        """Replay into ``sink`` (add_texture / add_material / add_triangle / set_triangle_uv / add_quad /
        add_sphere [/ add_triangles / add_spheres]).  Primitive ids follow ``order``."""
        for kind, p, rgb in self.textures:
            sink.add_texture(kind, p, rgb)
        for kind, p in self.materials:
            sink.add_material(kind, p)
        n = len(self.order)
        homogeneous_runs = n >= bulk_threshold and hasattr(sink, "add_triangles")
        if homogeneous_runs:
            i = 0
            while i < n:
                k = self.order[i][0]
                j = i
                while j < n and self.order[j][0] == k:
                    j += 1
                idx = [self.order[m][1] for m in range(i, j)]
                if k == "t" and all(self.tris[m][5] is None for m in idx):
                    sink.add_triangles(np.stack([self.tris[m][0] for m in idx]), np.stack([self.tris[m][1] for m in idx]),
                                       np.stack([self.tris[m][2] for m in idx]),
                                       np.array([self.tris[m][3] for m in idx], np.int32), np.array([self.tris[m][4] for m in idx], np.int32))
                elif k == "s":
                    sink.add_spheres(np.stack([self.spheres[m][0] for m in idx]), np.array([self.spheres[m][1] for m in idx]),
                                     np.array([self.spheres[m][2] for m in idx], np.int32), np.array([self.spheres[m][3] for m in idx], np.int32))
                else:
                    for m in range(i, j):
                        self._feed_one(sink, *self.order[m])
                i = j
        else:
            for k, m in self.order:
                self._feed_one(sink, k, m)
        return sink

    def _feed_one(self, sink, k, m):
        if k == "t":
            Q, u, v, mat, tex, uv = self.tris[m]
            pid = sink.add_triangle(Q, u, v, mat, tex)
            if uv is not None:
                sink.set_triangle_uv(pid, uv)
        elif k == "q":
            sink.add_quad(*self.quads[m])
        else:
            sink.add_sphere(*self.spheres[m])

    def camera_args(self):
        c = dict(pos=(0, 0, 1), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=40.0, focus_dist=1.0, defocus_angle_deg=0.0, jitter=1)
        c.update(self.camera)
        return c

    def params_args(self, **over):
        p = dict(width=self.width, height=self.height, sample_begin=0, sample_count=self.spp, max_depth=self.max_depth,
                 integrator=self.integrator, traversal=0, ao_samples=self.ao_samples, seed=1, t_min=self.t_min,
                 background_bottom=self.background_bottom, background_top=self.background_top)
        p.update(over)
        return p


# ---------------------------------------------------------------------------------------------------------
# config 0 — rt.cpp's scene (experiments/rt.cpp:153-185)
# ---------------------------------------------------------------------------------------------------------
def rt_cornell(width=512, height=512, diffuse_walls=False):
    """diffuse_walls=True: the five room walls are MAT_DIFFUSE instead of MAT_METAL 0.90 (the boxes stay mirrors), which is
    what makes rt.cpp's cosine gather + Russian roulette (rt.cpp:278-329) run at all — in the scene as shipped it is dead
    code.  oracle/_ref/rt_ref_diffuse_counted is the reference program with exactly that edit."""
    s = SceneDesc("rt_cornell_diffuse" if diffuse_walls else "rt_cornell", width=width, height=height, spp=1, max_depth=2, integrator=INTEGRATOR_RT_AO,
                  ao_samples=32, t_min=1e-5, background_bottom=(0.06, 0.09, 0.14), background_top=(0.06, 0.09, 0.14))
    red, green, white = s.solid(0.8, 0.15, 0.15), s.solid(0.15, 0.8, 0.15), s.solid(0.8, 0.8, 0.8)
    checker = s.tex(TEX_CHECKER_UV, 8, 0.9, 0.9, 0.9, 0.1, 0.1, 0.1)
    meta = s.solid(0.93, 0.95, 1.0)
    if diffuse_walls:
        wall = s.mat(MAT_DIFFUSE)
    else:
        wall = s.mat(MAT_REFLECTIVE, 0.90, 0.0, 0.0, 0.0)    # MAT_METAL 0.90 with the default (black) tint, rt.cpp:160-164
    box = s.mat(MAT_REFLECTIVE, 0.90, 0.92, 0.94, 1.0)       # rt.cpp:165-166
    f = np.float32  # the original holds fp32 coordinates
    s.quad_as_tris((-1, -1, -1), (-1, 1, -1), (-1, 1, 1), (-1, -1, 1), wall, red)
    s.quad_as_tris((1, -1, 1), (1, 1, 1), (1, 1, -1), (1, -1, -1), wall, green)
    s.quad_as_tris((-1, -1, 1), (1, -1, 1), (1, -1, -1), (-1, -1, -1), wall, checker)
    s.quad_as_tris((-1, 1, 1), (-1, 1, -1), (1, 1, -1), (1, 1, 1), wall, white)
    s.quad_as_tris((-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), wall, white)

    def push_box(cx, cy, cz, r, tex):
        X = (float(f(cx) - f(r)), float(f(cx) + f(r)))
        Y = (float(f(cy) - f(r)), float(f(cy) + f(r)))
        Z = (float(f(cz) - f(r)), float(f(cz) + f(r)))
        q = s.quad_as_tris
        q((X[0], Y[0], Z[0]), (X[1], Y[0], Z[0]), (X[1], Y[1], Z[0]), (X[0], Y[1], Z[0]), box, tex)
        q((X[0], Y[0], Z[1]), (X[1], Y[0], Z[1]), (X[1], Y[1], Z[1]), (X[0], Y[1], Z[1]), box, tex)
        q((X[0], Y[0], Z[0]), (X[0], Y[1], Z[0]), (X[0], Y[1], Z[1]), (X[0], Y[0], Z[1]), box, tex)
        q((X[1], Y[0], Z[0]), (X[1], Y[1], Z[0]), (X[1], Y[1], Z[1]), (X[1], Y[0], Z[1]), box, tex)
        q((X[0], Y[1], Z[0]), (X[1], Y[1], Z[0]), (X[1], Y[1], Z[1]), (X[0], Y[1], Z[1]), box, tex)
        q((X[0], Y[0], Z[0]), (X[1], Y[0], Z[0]), (X[1], Y[0], Z[1]), (X[0], Y[0], Z[1]), box, tex)

    push_box(-0.5, -0.9, 0.4, 0.15, white)
    push_box(0.0, -0.9, 0.4, 0.10, meta)
    s.camera = dict(pos=(0, 0, 4), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=50.0, jitter=0)
    return s


# ---------------------------------------------------------------------------------------------------------
# config 3 — 555-unit Cornell box
# ---------------------------------------------------------------------------------------------------------
def _box_tris(s, a, b, mat, tex, rot_y_deg, offset):
    a, b = np.minimum(a, b).astype(float), np.maximum(a, b).astype(float)
    th = math.radians(rot_y_deg)
    c, sn = math.cos(th), math.sin(th)

    def xf(p):
        p = np.asarray(p, float)
        return np.array([c * p[0] + sn * p[2], p[1], -sn * p[0] + c * p[2]]) + np.asarray(offset, float)

    dx, dy, dz = np.array([b[0] - a[0], 0, 0]), np.array([0, b[1] - a[1], 0]), np.array([0, 0, b[2] - a[2]])
    faces = [
        (np.array([a[0], a[1], b[2]]), dx, dy),    # front
        (np.array([b[0], a[1], b[2]]), -dz, dy),   # right
        (np.array([b[0], a[1], a[2]]), -dx, dy),   # back
        (np.array([a[0], a[1], a[2]]), dz, dy),    # left
        (np.array([a[0], b[1], b[2]]), dx, -dz),   # top
        (np.array([a[0], a[1], a[2]]), dx, dz),    # bottom
    ]
    for Q, u, v in faces:
        s.quad_as_tris(xf(Q), xf(Q + u), xf(Q + u + v), xf(Q + v), mat, tex)


def cornell_box(width=2048, height=2048, spp=16384, as_quads=False):
    """RTIOW-dimension Cornell box.  BASELINE config 3 asks for the box "as triangles" (as_quads=False:
    5 walls + light + 2 boxes = 36 triangles); as_quads=True keeps the six flat surfaces as are quads."""
    s = SceneDesc("cornell_box", width=width, height=height, spp=spp, max_depth=50, t_min=1e-3,
                  background_bottom=(0, 0, 0), background_top=(0, 0, 0))
    red, white, green, light = s.solid(.65, .05, .05), s.solid(.73, .73, .73), s.solid(.12, .45, .15), s.solid(15, 15, 15)
    lam = s.mat(MAT_LAMBERTIAN, -1)
    emit = s.mat(MAT_DIFFUSE_LIGHT, -1, 1.0)
    flat = [
        ((555, 0, 0), (0, 555, 0), (0, 0, 555), lam, green),
        ((0, 0, 0), (0, 555, 0), (0, 0, 555), lam, red),
        ((343, 554, 332), (-130, 0, 0), (0, 0, -105), emit, light),
        ((0, 0, 0), (555, 0, 0), (0, 0, 555), lam, white),
        ((555, 555, 555), (-555, 0, 0), (0, 0, -555), lam, white),
        ((0, 0, 555), (555, 0, 0), (0, 555, 0), lam, white),
    ]
    for Q, u, v, m, t in flat:
        Q, u, v = np.array(Q, float), np.array(u, float), np.array(v, float)
        if as_quads:
            s.quad(Q, u, v, m, t)
        else:
            s.quad_as_tris(Q, Q + u, Q + u + v, Q + v, m, t)
    _box_tris(s, np.array([0, 0, 0]), np.array([165, 330, 165]), lam, white, 15.0, (265, 0, 295))
    _box_tris(s, np.array([0, 0, 0]), np.array([165, 165, 165]), lam, white, -18.0, (130, 0, 65))
    s.camera = dict(pos=(278, 278, -800), target=(278, 278, 0), up=(0, 1, 0), vfov_deg=40.0, focus_dist=10.0, jitter=1)
    return s


# ---------------------------------------------------------------------------------------------------------
# config 1 — RTIOW final scene
# ---------------------------------------------------------------------------------------------------------
def rtiow_final(width=1200, height=675, spp=500, seed=1):
    rng = np.random.RandomState(seed)
    s = SceneDesc("rtiow_final", width=width, height=height, spp=spp, max_depth=50, t_min=1e-3)
    lam = s.mat(MAT_LAMBERTIAN, -1)
    glass = s.mat(MAT_DIELECTRIC, 1.5)
    s.sphere((0, -1000, 0), 1000, lam, s.solid(0.5, 0.5, 0.5))
    white = s.solid(1, 1, 1)
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = rng.rand()
            c = np.array([a + 0.9 * rng.rand(), 0.2, b + 0.9 * rng.rand()])
            col = rng.rand(3)
            col2 = rng.rand(3)
            fuzz = 0.5 * rng.rand()
            if np.linalg.norm(c - np.array([4, 0.2, 0])) <= 0.9:
                continue
            if choose < 0.8:
                s.sphere(c, 0.2, lam, s.solid(*(col * col2)))
            elif choose < 0.95:
                s.sphere(c, 0.2, s.mat(MAT_METAL, fuzz, -1), s.solid(*(0.5 + 0.5 * col)))
            else:
                s.sphere(c, 0.2, glass, white)
    s.sphere((0, 1, 0), 1.0, glass, white)
    s.sphere((-4, 1, 0), 1.0, lam, s.solid(0.4, 0.2, 0.1))
    s.sphere((4, 1, 0), 1.0, s.mat(MAT_METAL, 0.0, -1), s.solid(0.7, 0.6, 0.5))
    s.camera = dict(pos=(13, 2, 3), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=20.0, focus_dist=10.0,
                    defocus_angle_deg=0.6, jitter=1)
    return s


# ---------------------------------------------------------------------------------------------------------
# config 2 — textured scene
# ---------------------------------------------------------------------------------------------------------
def synthetic_image(w=1024, h=512):
    """The synthetic 'earth' of SURVEY.md §8d: r = x%256, g = y%256, b = (x^y)%256, as bytes."""
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    return np.stack([np.broadcast_to(x % 256, (h, w)), np.broadcast_to(y % 256, (h, w)), (x ^ y) % 256], axis=-1).astype(np.uint8)


def textured(width=1920, height=1080, spp=1024, image_rgb=None):
    """image_rgb: (h,w,3) float64 in [0,1] as are::Texture holds it; default = synthetic_image()/255."""
    s = SceneDesc("textured", width=width, height=height, spp=spp, max_depth=50, t_min=1e-3,
                  background_bottom=(0, 0, 0), background_top=(0, 0, 0))
    if image_rgb is None:
        image_rgb = synthetic_image().astype(np.float64) / 255.0
    lam = s.mat(MAT_LAMBERTIAN, -1)
    emit = s.mat(MAT_DIFFUSE_LIGHT, -1, 1.0)
    checker = s.tex(TEX_CHECKER_3D, 0.32, .2, .3, .1, .9, .9, .9)
    noise = s.tex(TEX_NOISE, 4.0, 2)
    image = s.tex(TEX_IMAGE, rgb=np.ascontiguousarray(image_rgb))
    light = s.solid(4, 4, 4)
    # the ground sits at y = -0.1, inside a checker cell: a surface lying exactly ON a cell boundary (y = 0) would make
    # floor(y / scale) flip with the sign of the last rounding error, in fp64 and fp32 alike
    s.quad((-30, -0.1, -30), (60, 0, 0), (0, 0, 60), lam, checker)
    s.sphere((0, 1.9, 0), 2.0, lam, noise)
    s.sphere((4.5, 1.9, 1.5), 2.0, lam, image)
    s.sphere((-4.5, 1.4, 2.0), 1.5, s.mat(MAT_METAL, 0.05, -1), s.solid(0.8, 0.8, 0.9))
    s.quad((3, 1, -2), (2, 0, 0), (0, 2, 0), emit, light)
    s.quad((-3, 6, -3), (6, 0, 0), (0, 0, 6), emit, light)
    s.camera = dict(pos=(0, 4, 14), target=(0, 2, 0), up=(0, 1, 0), vfov_deg=35.0, focus_dist=10.0, jitter=1)
    return s


# ---------------------------------------------------------------------------------------------------------
# config 4 — traversal stress
# ---------------------------------------------------------------------------------------------------------
def stress(n_prims=1_000_000, width=3840, height=2160, spp=256, seed=4, extent=50.0):
    rng = np.random.RandomState(seed)
    s = SceneDesc("stress", width=width, height=height, spp=spp, max_depth=8, t_min=1e-3)
    lam = s.mat(MAT_LAMBERTIAN, -1)
    grey = s.solid(0.5, 0.5, 0.5)
    ns = n_prims // 2
    nt = n_prims - ns
    # density-preserving extent when n is scaled down for tests
    ext = extent * (n_prims / 1_000_000.0) ** (1.0 / 3.0) if n_prims < 1_000_000 else extent
    cs = rng.uniform(-ext, ext, size=(ns, 3))
    rs = rng.uniform(0.02, 0.05, size=ns)
    ct = rng.uniform(-ext, ext, size=(nt, 3))
    e1 = rng.normal(size=(nt, 3))
    e1 *= 0.1 / np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = rng.normal(size=(nt, 3))
    e2 -= e1 * (np.sum(e1 * e2, axis=1, keepdims=True) / 0.01) * 0.5   # keep well away from collinear
    e2 *= 0.1 / np.linalg.norm(e2, axis=1, keepdims=True)
    for i in range(ns):
        s.spheres.append((cs[i], float(rs[i]), lam, grey))
        s.order.append(("s", i))
    for i in range(nt):
        s.tris.append((ct[i], e1[i], e2[i], lam, grey, None))
        s.order.append(("t", i))
    s.camera = dict(pos=(0, 0, 2.6 * ext), target=(0, 0, 0), up=(0, 1, 0), vfov_deg=40.0, focus_dist=10.0, jitter=1)
    return s


def by_name(name, **kw):
    if name == "rt_cornell_diffuse":
        return rt_cornell(diffuse_walls=True, **kw)
    return {"rt_cornell": rt_cornell, "cornell_box": cornell_box, "rtiow_final": rtiow_final, "textured": textured,
            "stress": stress}[name](**kw)


# ---------------------------------------------------------------------------------------------------------
# patch-as-viewport renderer: scenes for are_cuda_patch_render
# ---------------------------------------------------------------------------------------------------------
PATCH_DIFFUSE, PATCH_REFLECTIVE = 0, 1


@dataclass
class PatchScene:
    """Input of the reference's patch renderer (/root/reference/experiments/rt10.cpp): uv-mapped triangles with
    (type, albedo, metalness) materials, an eye point, a two-triangle viewport and RenderConfig (rt10.cpp:536-542)."""
    name: str
    P: np.ndarray            # (n,3,3) vertices
    UV: np.ndarray           # (n,3,2) texture coordinates
    material: np.ndarray     # (n,) int32 material index, -1 = none (white, diffuse)
    mat_type: np.ndarray     # (m,) int32 PATCH_DIFFUSE | PATCH_REFLECTIVE
    mat_albedo: np.ndarray   # (m,3)
    mat_metalness: np.ndarray  # (m,)
    origin: np.ndarray       # (3,)
    vp_P: np.ndarray         # (2,3,3) the two viewport triangles
    vp_UV: np.ndarray        # (2,3,2)
    width: int = 900
    height: int = 650
    max_depth: int = 4
    min_area_px: float = 6.0
    max_tex_res: int = 256
    min_tex_res: int = 16
    env: tuple = (0.06, 0.07, 0.09)
    gamma: float = 2.2

    def cfg8(self):
        return np.array([self.max_depth, self.min_area_px, self.max_tex_res, self.min_tex_res, *self.env, self.gamma], dtype=np.float64)


class _PatchBuilder:
    def __init__(self):
        self.P, self.UV, self.M = [], [], []

    def quad(self, p00, p10, p11, p01, mat):  # rt10.cpp:778-800: two triangles, uv (0,0)-(1,1) per quad
        self.P += [[p00, p10, p01], [p10, p11, p01]]
        self.UV += [[(0, 0), (1, 0), (0, 1)], [(1, 0), (1, 1), (0, 1)]]
        self.M += [mat, mat]

    def box(self, lo, hi, mat):  # rt10.cpp:802-830, faces -X +X -Y +Y -Z +Z
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        p = {(i, j, k): ((x0, x1)[i], (y0, y1)[j], (z0, z1)[k]) for i in (0, 1) for j in (0, 1) for k in (0, 1)}
        self.quad(p[0, 0, 0], p[0, 0, 1], p[0, 1, 1], p[0, 1, 0], mat)
        self.quad(p[1, 0, 0], p[1, 1, 0], p[1, 1, 1], p[1, 0, 1], mat)
        self.quad(p[0, 0, 0], p[1, 0, 0], p[1, 0, 1], p[0, 0, 1], mat)
        self.quad(p[0, 1, 0], p[0, 1, 1], p[1, 1, 1], p[1, 1, 0], mat)
        self.quad(p[0, 0, 0], p[0, 1, 0], p[1, 1, 0], p[1, 0, 0], mat)
        self.quad(p[0, 0, 1], p[1, 0, 1], p[1, 1, 1], p[0, 1, 1], mat)

    def arrays(self):
        return (np.array(self.P, dtype=np.float64).reshape(-1, 3, 3), np.array(self.UV, dtype=np.float64).reshape(-1, 3, 2),
                np.array(self.M, dtype=np.int32))


def _patch_viewport(width, height, center, vp_h):
    """The camera of rt10.cpp:895-921: viewport rectangle in the plane z = center.z, u right, v down."""
    aspect = float(width) / float(height)
    vp_w = vp_h * aspect
    cx, cy, cz = center
    TL = (cx - vp_w * 0.5, cy - vp_h * 0.5, cz)
    TR = (cx + vp_w * 0.5, cy - vp_h * 0.5, cz)
    BL = (cx - vp_w * 0.5, cy + vp_h * 0.5, cz)
    BR = (cx + vp_w * 0.5, cy + vp_h * 0.5, cz)
    vp_P = np.array([[TL, TR, BL], [TR, BR, BL]], dtype=np.float64)
    vp_UV = np.array([[(0, 0), (1, 0), (0, 1)], [(1, 0), (1, 1), (0, 1)]], dtype=np.float64)
    return vp_P, vp_UV


def patch_rt10(width=900, height=650, max_depth=4):
    """The scene hard-wired into the reference program (rt10.cpp:832-925): open-front Cornell room x[-1,1] y[0,2] z[0,2],
    a blue and a yellow metal box, eye (0,1,-3), viewport plane z=-2 of world height 1.6.  At the defaults the
    reference's output is its shipped experiments/output_rt10.ppm."""
    b = _PatchBuilder()
    white, red, green, blue, yellow = 0, 1, 2, 3, 4
    b.quad((-1, 0, 0), (1, 0, 0), (1, 0, 2), (-1, 0, 2), white)
    b.quad((-1, 2, 0), (-1, 2, 2), (1, 2, 2), (1, 2, 0), white)
    b.quad((-1, 0, 2), (1, 0, 2), (1, 2, 2), (-1, 2, 2), white)
    b.quad((-1, 0, 0), (-1, 0, 2), (-1, 2, 2), (-1, 2, 0), red)
    b.quad((1, 0, 0), (1, 2, 0), (1, 2, 2), (1, 0, 2), green)
    b.box((-0.70, 0.0, 0.80), (-0.15, 0.60, 1.30), blue)
    b.box((0.15, 0.0, 1.00), (0.70, 1.10, 1.65), yellow)
    P, UV, M = b.arrays()
    vp_P, vp_UV = _patch_viewport(width, height, (0.0, 1.0, -2.0), 1.6)
    return PatchScene("patch_rt10", P, UV, M,
                      mat_type=np.array([0, 0, 0, 1, 1], dtype=np.int32),
                      mat_albedo=np.array([[0.85, 0.85, 0.85], [0.85, 0.25, 0.25], [0.25, 0.85, 0.25], [0.25, 0.45, 1.0], [1.0, 0.92, 0.20]]),
                      mat_metalness=np.array([0.0, 0.0, 0.0, 0.90, 0.92]),
                      origin=np.array([0.0, 1.0, -3.0]), vp_P=vp_P, vp_UV=vp_UV, width=width, height=height, max_depth=max_depth)


def patch_random(seed, n_boxes=3, n_loose=6, width=160, height=120, max_depth=3, mirror_walls=False):
    """Random rooms for parity tests: the rt10 room with random boxes (random metal / diffuse materials), a few loose
    triangles (some without a material), optional mirror walls (mirror facing mirror: exercises the cycle guard and the
    depth / area cut-offs)."""
    rng = np.random.RandomState(seed)
    n_mat = 6
    mat_type = rng.randint(0, 2, n_mat).astype(np.int32)
    mat_type[0], mat_type[1] = 0, 1
    mat_albedo = rng.uniform(0.1, 1.0, (n_mat, 3))
    mat_metal = rng.uniform(-0.1, 1.1, n_mat)  # outside [0,1] on purpose: the renderer clamps (rt10.cpp:651)
    b = _PatchBuilder()
    wall = 1 if mirror_walls else 0
    b.quad((-1, 0, 0), (1, 0, 0), (1, 0, 2), (-1, 0, 2), 0)
    b.quad((-1, 2, 0), (-1, 2, 2), (1, 2, 2), (1, 2, 0), int(rng.randint(0, n_mat)))
    b.quad((-1, 0, 2), (1, 0, 2), (1, 2, 2), (-1, 2, 2), int(rng.randint(0, n_mat)))
    b.quad((-1, 0, 0), (-1, 0, 2), (-1, 2, 2), (-1, 2, 0), wall)
    b.quad((1, 0, 0), (1, 2, 0), (1, 2, 2), (1, 0, 2), wall)
    for _ in range(n_boxes):
        lo = np.array([rng.uniform(-0.9, 0.4), 0.0, rng.uniform(0.3, 1.3)])
        hi = lo + np.array([rng.uniform(0.2, 0.5), rng.uniform(0.2, 1.2), rng.uniform(0.2, 0.5)])
        b.box(tuple(lo), tuple(hi), int(rng.randint(0, n_mat)))
    P, UV, M = b.arrays()
    if n_loose:
        c = rng.uniform([-0.8, 0.2, 0.3], [0.8, 1.8, 1.8], (n_loose, 1, 3))
        LP = c + rng.uniform(-0.35, 0.35, (n_loose, 3, 3))
        LUV = rng.uniform(0, 1, (n_loose, 3, 2))
        LM = rng.randint(-1, n_mat, n_loose).astype(np.int32)
        P, UV, M = np.concatenate([P, LP]), np.concatenate([UV, LUV]), np.concatenate([M, LM])
    vp_P, vp_UV = _patch_viewport(width, height, (rng.uniform(-0.1, 0.1), 1.0 + rng.uniform(-0.1, 0.1), -2.0), 1.6)
    origin = np.array([rng.uniform(-0.3, 0.3), 1.0 + rng.uniform(-0.2, 0.2), -3.0])
    return PatchScene(f"patch_random_{seed}", P, UV, M.astype(np.int32), mat_type, mat_albedo, mat_metal, origin, vp_P, vp_UV,
                      width=width, height=height, max_depth=max_depth, min_area_px=float(rng.choice([2.0, 6.0, 20.0])),
                      max_tex_res=int(rng.choice([64, 128, 256])), min_tex_res=int(rng.choice([4, 16])),
                      env=tuple(rng.uniform(0, 0.2, 3)), gamma=2.2)
