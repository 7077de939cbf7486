"""ctypes bindings of the CPU checkers — TEST INFRASTRUCTURE ONLY.

``Oracle``   : oracle/liboracle.so   (oracle/are_oracle.c — the C restatement + fp64 twin)
``Reference``: oracle/_ref/libare_ref.so (the real reference library behind oracle/ref_harness.cpp)

The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from aurora_rendering_engine_b200.capi import Camera, RenderParams, RenderStats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libare_ref.so")
RT_REF = os.path.join(ORACLE_DIR, "_ref", "rt_ref")
RT_REF_COUNTED = os.path.join(ORACLE_DIR, "_ref", "rt_ref_counted")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


def build_oracle():
    """(Re)build the checkers: the C restatement always, the real reference when /root/reference exists."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


class _VecLib:
    """The flat per-function face shared (same names modulo prefix) by the restatement and the real reference."""

    def __init__(self, lib, prefix):
        self.lib, self.pre = lib, prefix

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def vec3_binary(self, op, a, b=None, s=None):
        a = _d(a)
        out = np.empty_like(a)
        b_ = _d(b) if b is not None else None
        s_ = _d(s) if s is not None else None
        self._f("vec3_binary")(C.c_int(op), C.c_int(len(a)), _p(a), _p(b_) if b_ is not None else None, _p(s_) if s_ is not None else None, _p(out))
        return out

    def vec3_scalar(self, op, a, b=None):
        a = _d(a)
        out = np.empty(len(a))
        b_ = _d(b) if b is not None else None
        self._f("vec3_scalar")(C.c_int(op), C.c_int(len(a)), _p(a), _p(b_) if b_ is not None else None, _p(out))
        return out

    def reflect(self, v, n):
        v, n = _d(v), _d(n)
        out = np.empty_like(v)
        self._f("reflect")(C.c_int(len(v)), _p(v), _p(n), _p(out))
        return out

    def refract(self, uv, n, eta):
        uv, n, eta = _d(uv), _d(n), _d(eta)
        out = np.empty_like(uv)
        self._f("refract")(C.c_int(len(uv)), _p(uv), _p(n), _p(eta), _p(out))
        return out

    def ray(self, Q, D, t):
        Q, D, t = _d(Q), _d(D), _d(t)
        oD, oA = np.empty_like(Q), np.empty_like(Q)
        self._f("ray")(C.c_int(len(Q)), _p(Q), _p(D), _p(t), _p(oD), _p(oA))
        return oD, oA

    def plane_from_point_normal(self, p, n):
        p, n = _d(p), _d(n)
        out = np.empty((len(p), 4))
        self._f("plane_from_point_normal")(C.c_int(len(p)), _p(p), _p(n), _p(out))
        return out

    def plane_intersect(self, plane4, Q, D):
        plane4, Q, D = _d(plane4), _d(Q), _d(D)
        hit = np.empty(len(Q), np.int32)
        P = np.empty_like(Q)
        self._f("plane_intersect")(C.c_int(len(Q)), _p(plane4), _p(Q), _p(D), _p(hit, _ip), _p(P))
        return hit, P

    def material_reflect(self, kind, reflectivity, plane4, origin):
        plane4, origin = _d(plane4), _d(origin)
        ok = np.empty(len(origin), np.int32)
        out = np.empty_like(origin)
        self._f("material_reflect")(C.c_int(kind), C.c_double(reflectivity), C.c_int(len(origin)), _p(plane4), _p(origin), _p(ok, _ip), _p(out))
        return ok, out

    def triangle_ctor(self, Q, u, v, flags=0):
        Q, u, v = _d(Q), _d(u), _d(v)
        verts = np.full(9, np.nan)
        fn = self._f("triangle_ctor")
        fn.restype = C.c_int
        st = fn(_p(Q), _p(u), _p(v), C.c_int(flags), _p(verts))
        return st, verts.reshape(3, 3)

    def texture_load(self, path):
        fn = self._f("texture_load")
        fn.restype = C.c_int
        w, h = C.c_int(0), C.c_int(0)
        st = fn(os.fsencode(path), C.byref(w), C.byref(h), None, C.c_long(0))
        if st != 0:
            return st, None
        rgb = np.empty((h.value, w.value, 3))
        st = fn(os.fsencode(path), C.byref(w), C.byref(h), _p(rgb), C.c_long(rgb.size))
        return st, rgb

    def texture_paste(self, dst, src, corners):
        """are::Texture::paste; returns the updated copy of dst (h,w,3). corners = (lt, rt, lb, rb) integer pixel pairs."""
        out = np.array(dst, dtype=np.float64, order="C", copy=True)
        src = _d(src)
        c = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(8))
        self._f("texture_paste")(_p(out), C.c_int(out.shape[1]), C.c_int(out.shape[0]), _p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), _p(c, _ip))
        return out

    def texture_save(self, path, rgb):
        rgb = _d(rgb)
        fn = self._f("texture_save")
        fn.restype = C.c_int
        return fn(os.fsencode(path), C.c_int(rgb.shape[1]), C.c_int(rgb.shape[0]), _p(rgb))


class Reference(_VecLib):
    """The REAL reference library (only present where oracle/_ref was built or shipped)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        super().__init__(C.CDLL(REF_SO), "ref_")
        self.lib.ref_triset_create.restype = _vp
        self.lib.ref_triset_closest_hit.restype = C.c_long
        self.lib.ref_geometry_epsilon.restype = C.c_double

    def sizeof(self):
        out = (C.c_int * 5)()
        self.lib.ref_sizeof(out)
        return list(out)

    def geometry_epsilon(self):
        return self.lib.ref_geometry_epsilon()

    def texture_fill_ctor(self, w, h):
        return self.lib.ref_texture_fill_ctor(C.c_int(w), C.c_int(h))

    class TriSet:
        def __init__(self, ref, TQ, Tu, Tv):
            self.ref = ref
            TQ, Tu, Tv = _d(TQ), _d(Tu), _d(Tv)
            self.n = len(TQ)
            self.h = ref.lib.ref_triset_create(C.c_int(self.n), _p(TQ), _p(Tu), _p(Tv))
            if not self.h:
                raise ValueError("reference rejected a triangle")

        def close(self):
            if self.h:
                self.ref.lib.ref_triset_destroy(_vp(self.h))
                self.h = None

        def __del__(self):
            self.close()

        def hit_matrix(self, Q, D):
            Q, D = _d(Q), _d(D)
            hit = np.empty((len(Q), self.n), np.int32)
            P = np.empty((len(Q), self.n, 3))
            self.ref.lib.ref_triset_hit_matrix(_vp(self.h), C.c_int(len(Q)), _p(Q), _p(D), _p(hit, _ip), _p(P))
            return hit, P

        def closest_hit(self, Q, D):
            Q, D = _d(Q), _d(D)
            prim, t, P = np.empty(len(Q), np.int32), np.empty(len(Q)), np.empty_like(Q)
            n = self.ref.lib.ref_triset_closest_hit(_vp(self.h), C.c_int(len(Q)), _p(Q), _p(D), _p(prim, _ip), _p(t), _p(P))
            return n, prim, t, P

        def point_in(self, tri, pts):
            pts = _d(pts)
            inside = np.empty(len(pts), np.int32)
            self.ref.lib.ref_triset_point_in(_vp(self.h), C.c_int(tri), C.c_int(len(pts)), _p(pts), _p(inside, _ip))
            return inside

    def triset(self, TQ, Tu, Tv):
        return Reference.TriSet(self, TQ, Tu, Tv)


class Oracle(_VecLib):
    """oracle/are_oracle.c."""

    def __init__(self, path=None):
        """path: another build of the same source (bench.py times oracle/liboracle_fast.so, -O3 -march=native)."""
        if path is None and not os.path.exists(ORACLE_SO):
            build_oracle()
        super().__init__(C.CDLL(path or ORACLE_SO), "lib_")
        L = self.lib
        L.lib_triset_closest_hit.restype = C.c_long
        L.orc_scene_create.restype = _vp
        for n in ("orc_add_texture", "orc_add_material", "orc_add_triangle", "orc_set_triangle_uv", "orc_add_quad", "orc_add_sphere",
                  "orc_num_primitives", "orc_render", "orc_render_window"):
            getattr(L, n).restype = C.c_int

    # -- triangle list functions taking raw arrays --
    def triset_hit_matrix(self, TQ, Tu, Tv, Q, D):
        TQ, Tu, Tv, Q, D = _d(TQ), _d(Tu), _d(Tv), _d(Q), _d(D)
        hit = np.empty((len(Q), len(TQ)), np.int32)
        P = np.empty((len(Q), len(TQ), 3))
        self.lib.lib_triset_hit_matrix(C.c_int(len(TQ)), _p(TQ), _p(Tu), _p(Tv), C.c_int(len(Q)), _p(Q), _p(D), _p(hit, _ip), _p(P))
        return hit, P

    def triset_closest_hit(self, TQ, Tu, Tv, Q, D):
        TQ, Tu, Tv, Q, D = _d(TQ), _d(Tu), _d(Tv), _d(Q), _d(D)
        prim, t, P = np.empty(len(Q), np.int32), np.empty(len(Q)), np.empty_like(Q)
        n = self.lib.lib_triset_closest_hit(C.c_int(len(TQ)), _p(TQ), _p(Tu), _p(Tv), C.c_int(len(Q)), _p(Q), _p(D), _p(prim, _ip), _p(t), _p(P))
        return n, prim, t, P

    def triset_point_in(self, TQ, Tu, Tv, tri, pts):
        TQ, Tu, Tv, pts = _d(TQ), _d(Tu), _d(Tv), _d(pts)
        inside = np.empty(len(pts), np.int32)
        self.lib.lib_triset_point_in(_p(TQ), _p(Tu), _p(Tv), C.c_int(tri), C.c_int(len(pts)), _p(pts), _p(inside, _ip))
        return inside

    def encode_linear(self, c):
        c = _d(c).ravel()
        out = np.empty(c.size, np.uint8)
        self.lib.lib_encode_linear(C.c_long(c.size), _p(c), _p(out, C.POINTER(C.c_uint8)))
        return out

    def encode_gamma22(self, c):
        c = np.ascontiguousarray(c, np.float32).ravel()
        out = np.empty(c.size, np.uint8)
        self.lib.lib_encode_gamma22(C.c_long(c.size), _p(c, C.POINTER(C.c_float)), _p(out, C.POINTER(C.c_uint8)))
        return out

    def encode_sqrt(self, c):
        c = np.ascontiguousarray(c, np.float32).ravel()
        out = np.empty(c.size, np.uint8)
        self.lib.lib_encode_sqrt(C.c_long(c.size), _p(c, C.POINTER(C.c_float)), _p(out, C.POINTER(C.c_uint8)))
        return out

    def philox(self, seed, counters):
        c = np.ascontiguousarray(counters, np.uint32)
        out = np.empty_like(c)
        self.lib.orc_philox(C.c_int(len(c)), C.c_uint64(seed), _p(c, C.POINTER(C.c_uint32)), _p(out, C.POINTER(C.c_uint32)))
        return out

    def camera_rays(self, cam: Camera, W, H, px, py, rnd):
        px, py, rnd = np.ascontiguousarray(px, np.int32), np.ascontiguousarray(py, np.int32), _d(rnd)
        Q, D = np.empty((len(px), 3)), np.empty((len(px), 3))
        self.lib.orc_camera_rays(C.byref(cam), C.c_int(W), C.c_int(H), C.c_int(len(px)), _p(px, _ip), _p(py, _ip), _p(rnd), _p(Q), _p(D))
        return Q, D

    def scene(self):
        return OracleScene(self)


class OracleScene:
    """Same add_* vocabulary as aurora_rendering_engine_b200.capi.Context, on the CPU twin."""

    def __init__(self, orc: Oracle):
        self.o = orc
        self.L = orc.lib
        self.h = self.L.orc_scene_create()

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(_vp(self.h))
            self.h = None

    def __del__(self):
        self.close()

    def add_texture(self, kind, params, rgb=None):
        p = _d(params)
        if rgb is not None:
            img = _d(rgb)
            return self.L.orc_add_texture(_vp(self.h), C.c_int(kind), _p(p), _p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]))
        return self.L.orc_add_texture(_vp(self.h), C.c_int(kind), _p(p), None, C.c_int(0), C.c_int(0))

    def add_material(self, kind, params):
        p = _d(params)
        return self.L.orc_add_material(_vp(self.h), C.c_int(kind), _p(p))

    def add_triangle(self, Q, u, v, mat, tex):
        Q, u, v = _d(Q), _d(u), _d(v)
        return self.L.orc_add_triangle(_vp(self.h), _p(Q), _p(u), _p(v), C.c_int(mat), C.c_int(tex))

    def set_triangle_uv(self, prim, uv):
        uv = _d(uv)
        return self.L.orc_set_triangle_uv(_vp(self.h), C.c_int(prim), _p(uv))

    def add_quad(self, Q, u, v, mat, tex):
        Q, u, v = _d(Q), _d(u), _d(v)
        return self.L.orc_add_quad(_vp(self.h), _p(Q), _p(u), _p(v), C.c_int(mat), C.c_int(tex))

    def add_sphere(self, c, r, mat, tex):
        c = _d(c)
        return self.L.orc_add_sphere(_vp(self.h), _p(c), C.c_double(r), C.c_int(mat), C.c_int(tex))

    def num_primitives(self):
        return self.L.orc_num_primitives(_vp(self.h))

    def hit_batch(self, Q, D, t_min=0.0):
        Q, D = _d(Q), _d(D)
        n = len(Q)
        prim = np.empty(n, np.int32)
        t, P, N, uv = np.empty(n), np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 2))
        self.L.orc_hit_batch(_vp(self.h), C.c_int(n), _p(Q), _p(D), C.c_double(t_min), _p(prim, _ip), _p(t), _p(P), _p(N), _p(uv))
        return prim, t, P, N, uv

    def scatter_batch(self, mat, tex, wi, N, P, uv, rnd):
        mat, tex = np.ascontiguousarray(mat, np.int32), np.ascontiguousarray(tex, np.int32)
        wi, N, P, uv, rnd = _d(wi), _d(N), _d(P), _d(uv), _d(rnd)
        n = len(mat)
        wo, att, emit, alive = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3)), np.empty(n, np.int32)
        self.L.orc_scatter_batch(_vp(self.h), C.c_int(n), _p(mat, _ip), _p(tex, _ip), _p(wi), _p(N), _p(P), _p(uv), _p(rnd),
                                 _p(wo), _p(att), _p(emit), _p(alive, _ip))
        return wo, att, emit, alive

    def texture_batch(self, tex, uv, P):
        tex = np.ascontiguousarray(tex, np.int32)
        uv, P = _d(uv), _d(P)
        rgb = np.empty((len(tex), 3))
        self.L.orc_texture_batch(_vp(self.h), C.c_int(len(tex)), _p(tex, _ip), _p(uv), _p(P), _p(rgb))
        return rgb

    def render(self, cam: Camera, params: RenderParams, nthreads=None, window=None):
        """Returns (accum float64 (H,W,3) of sample SUMS, RenderStats)."""
        nthreads = nthreads or os.cpu_count() or 1
        acc = np.zeros((params.height, params.width, 3))
        st = RenderStats()
        if window is None:
            window = (0, 0, params.width, params.height)
        self.L.orc_render_window(_vp(self.h), C.byref(cam), C.byref(params), C.c_int(window[0]), C.c_int(window[1]), C.c_int(window[2]),
                                 C.c_int(window[3]), _p(acc), C.c_int(nthreads), C.byref(st))
        return acc, st


def psnr(a, b, peak=1.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(peak * peak / mse)


# ---------------------------------------------------------------------------------------------------------
# patch-as-viewport renderer (experiments/rt10.cpp): restatement (oracle/libpatch_oracle.so) and the real thing
# (oracle/_ref/librt10_ref.so)
# ---------------------------------------------------------------------------------------------------------
_ip = C.POINTER(C.c_int)


class _PatchLib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix

    def _scene_args(self, ps):
        self._keep = [_d(ps.P), _d(ps.UV), np.ascontiguousarray(ps.material, dtype=np.int32), np.ascontiguousarray(ps.mat_type, dtype=np.int32),
                      _d(ps.mat_albedo), _d(ps.mat_metalness)]
        P, UV, M, T, A, Mt = self._keep
        return [C.c_int(len(M)), _p(P), _p(UV), _p(M, _ip), C.c_int(len(T)), _p(T, _ip), _p(A), _p(Mt)]

    def render(self, ps, want_counters=False):
        """-> (rgb (H,W,3) float64 linear, rgb8 (H,W,3) uint8 gamma-encoded P6 payload[, counters])"""
        rgb = np.zeros((ps.height, ps.width, 3), dtype=np.float64)
        rgb8 = np.zeros((ps.height, ps.width, 3), dtype=np.uint8)
        o, vP, vUV, cfg = _d(ps.origin), _d(ps.vp_P), _d(ps.vp_UV), ps.cfg8()
        args = self._scene_args(ps) + [_p(o), _p(vP), _p(vUV), C.c_int(ps.width), C.c_int(ps.height), _p(cfg), _p(rgb),
                                       rgb8.ctypes.data_as(C.POINTER(C.c_uint8))]
        cnt = np.zeros(3, dtype=np.uint64)
        if self.prefix == "orc":
            args.append(cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
        rc = getattr(self.lib, self.prefix + "_patch_render")(*args)
        assert rc == 0, rc
        return (rgb, rgb8, cnt) if want_counters else (rgb, rgb8)

    def trace_texture(self, ps, origin, current, tex_w, tex_h, est_area_px=0.0):
        out = np.zeros((ps.max_tex_res * ps.max_tex_res * 3 + 3), dtype=np.float64)
        wh = np.zeros(2, dtype=np.int32)
        o, cfg = _d(origin), ps.cfg8()
        rc = getattr(self.lib, self.prefix + "_patch_trace_texture")(*self._scene_args(ps), _p(o), C.c_int(current), C.c_int(tex_w), C.c_int(tex_h),
                                                                     C.c_double(est_area_px), _p(cfg), _p(out), _p(wh, _ip))
        assert rc == 0, rc
        return out[: wh[0] * wh[1] * 3].reshape(wh[1], wh[0], 3).copy()


class PatchOracle(_PatchLib):
    def __init__(self):
        build_oracle()
        super().__init__(os.path.join(ORACLE_DIR, "libpatch_oracle.so"), "orc")


class PatchReference(_PatchLib):
    """The real experiments/rt10.cpp; only available where oracle/_ref/librt10_ref.so was built."""
    PATH = os.path.join(ORACLE_DIR, "_ref", "librt10_ref.so")

    def __init__(self):
        super().__init__(self.PATH, "ref")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)


# ---------------------------------------------------------------------------------------------------------
# experiments/rt.cpp with diffuse walls (oracle/_ref/rt_ref_diffuse_counted): the reference's cosine gather + RR
# ---------------------------------------------------------------------------------------------------------
RT_REF_DIFFUSE_COUNTED = os.path.join(ORACLE_DIR, "_ref", "rt_ref_diffuse_counted")


def run_rt_reference(binary, workdir, seed=1, timeout=600):
    """Run one of the compiled rt.cpp programs in `workdir` -> (image (512,512,3) float64 in [0,1] = bytes/255 of its
    out_diffuse.ppm, rays cast or None when the build does not count)."""
    import re
    env = dict(os.environ, ARE_RT_SEED=str(seed))
    r = subprocess.run([binary], cwd=str(workdir), check=True, capture_output=True, timeout=timeout, env=env)
    raw = open(os.path.join(str(workdir), "out_diffuse.ppm"), "rb").read()
    hdr = b"P6\n512 512\n255\n"
    assert raw.startswith(hdr)
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(512, 512, 3).astype(np.float64) / 255.0
    m = re.search(rb"ARE_COUNT rays=(\d+)", r.stdout)
    return img, (int(m.group(1)) if m else None)


def block_mean(img, b):
    h, w, c = img.shape
    return img.reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))
