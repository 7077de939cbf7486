"""ctypes binding of the C ABI in ``include/are_cuda.h`` (libare_b200.so).

This is plumbing only: every function here forwards to one ``are_cuda_*`` entry point.  There is no Python or
CPU implementation of any of them — if the shared library is missing, or no CUDA device is present, the calls
raise instead of falling back.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARE_B200_LIB") or os.path.join(_HERE, "lib", "libare_b200.so")  # override: A/B of kernel builds

ARE_OK = 0
STATUS_NAMES = {0: "ARE_OK", -1: "ARE_ERR_INVALID_ARGUMENT", -2: "ARE_ERR_RUNTIME", -3: "ARE_ERR_CUDA",
                -4: "ARE_ERR_NO_DEVICE", -5: "ARE_ERR_NOT_COMMITTED", -6: "ARE_ERR_IO"}


class Camera(C.Structure):
    """are_camera"""
    _fields_ = [("pos", C.c_double * 3), ("target", C.c_double * 3), ("up", C.c_double * 3), ("vfov_deg", C.c_double),
                ("focus_dist", C.c_double), ("defocus_angle_deg", C.c_double), ("jitter", C.c_int32), ("pad_", C.c_int32)]


class RenderParams(C.Structure):
    """are_render_params"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("sample_begin", C.c_int32), ("sample_count", C.c_int32),
                ("max_depth", C.c_int32), ("integrator", C.c_int32), ("traversal", C.c_int32), ("ao_samples", C.c_int32),
                ("seed", C.c_uint64), ("t_min", C.c_double), ("background_bottom", C.c_double * 3),
                ("background_top", C.c_double * 3)]


class RenderStats(C.Structure):
    """are_render_stats"""
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("tri_tests", C.c_uint64), ("quad_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("node_visits", C.c_uint64), ("box_tests", C.c_uint64), ("kernel_ms", C.c_double), ("launches", C.c_uint64),
                ("kernel_variant", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CommitInfo(C.Structure):
    """are_commit_info"""
    _fields_ = [("builder", C.c_int), ("bvh_nodes", C.c_int), ("bvh_height", C.c_int), ("hot_slots", C.c_int), ("host_compile_ms", C.c_double),
                ("host_bvh_ms", C.c_double), ("device_bvh_ms", C.c_double), ("device_bvh_launches", C.c_uint64), ("baked", C.c_int32), ("quant_area_permille", C.c_int32),
                ("bake_compile_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


BVH_BUILDER_HOST_SAH, BVH_BUILDER_DEVICE_LBVH = 0, 1
KERNEL_NONE, KERNEL_BRUTE, KERNEL_BRUTE_LEAN, KERNEL_BVH2, KERNEL_BVH2_BIG, KERNEL_WIDE, KERNEL_RT_AO, KERNEL_BRUTE_BAKED, KERNEL_WAVEFRONT, KERNEL_BVH4, KERNEL_BVH2_QUANT = range(11)
(OPT_LEAN_KERNEL, OPT_BAKED_KERNEL, OPT_BAKED_PACKED, OPT_FUSE_PARALLELOGRAMS, OPT_FUSE_BOXES, OPT_BUILD_WIDE, OPT_WIDE_MIN_NODES,
 OPT_LBVH_MAX_HEIGHT, OPT_L2_PERSIST_NODES, OPT_BUILD_BVH4, OPT_BAKED_MIN_BLOCKS, OPT_QUANTIZED_NODES) = range(1, 13)


def make_camera(pos, target, up=(0, 1, 0), vfov_deg=40.0, focus_dist=1.0, defocus_angle_deg=0.0, jitter=1) -> Camera:
    c = Camera()
    c.pos[:] = [float(x) for x in pos]
    c.target[:] = [float(x) for x in target]
    c.up[:] = [float(x) for x in up]
    c.vfov_deg, c.focus_dist, c.defocus_angle_deg, c.jitter = float(vfov_deg), float(focus_dist), float(defocus_angle_deg), int(jitter)
    return c


def make_params(width, height, sample_begin=0, sample_count=1, max_depth=50, integrator=0, traversal=0, ao_samples=32,
                seed=1, t_min=1e-3, background_bottom=(1, 1, 1), background_top=(0.5, 0.7, 1.0)) -> RenderParams:
    p = RenderParams()
    p.width, p.height, p.sample_begin, p.sample_count = int(width), int(height), int(sample_begin), int(sample_count)
    p.max_depth, p.integrator, p.traversal, p.ao_samples = int(max_depth), int(integrator), int(traversal), int(ao_samples)
    p.seed, p.t_min = int(seed), float(t_min)
    p.background_bottom[:] = [float(x) for x in background_bottom]
    p.background_top[:] = [float(x) for x in background_top]
    return p


class PatchSceneC(C.Structure):
    _fields_ = [("n_tri", C.c_int32), ("n_mat", C.c_int32), ("P", C.POINTER(C.c_double)), ("UV", C.POINTER(C.c_double)),
                ("material", C.POINTER(C.c_int32)), ("mat_type", C.POINTER(C.c_int32)), ("mat_albedo", C.POINTER(C.c_double)),
                ("mat_metalness", C.POINTER(C.c_double))]


class PatchConfig(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("max_tex_res", C.c_int32), ("min_tex_res", C.c_int32), ("pad_", C.c_int32),
                ("min_area_px", C.c_double), ("env", C.c_double * 3), ("gamma", C.c_double)]


class PatchStats(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("node_texels", C.c_uint64), ("ops", C.c_uint64), ("levels", C.c_uint64), ("launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("plan_ms", C.c_double), ("kernel_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class _PatchArgs:
    """Keeps the numpy arrays behind an are_patch_scene alive for the duration of a call."""

    def __init__(self, ps):
        self.P = np.ascontiguousarray(ps.P, np.float64)
        self.UV = np.ascontiguousarray(ps.UV, np.float64)
        self.M = np.ascontiguousarray(ps.material, np.int32)
        self.T = np.ascontiguousarray(ps.mat_type, np.int32)
        self.A = np.ascontiguousarray(ps.mat_albedo, np.float64)
        self.Mt = np.ascontiguousarray(ps.mat_metalness, np.float64)
        i32 = C.POINTER(C.c_int32)
        self.scene = PatchSceneC(len(self.M), len(self.T), self.P.ctypes.data_as(_dp), self.UV.ctypes.data_as(_dp), self.M.ctypes.data_as(i32),
                                 self.T.ctypes.data_as(i32), self.A.ctypes.data_as(_dp), self.Mt.ctypes.data_as(_dp))
        self.cfg = PatchConfig(int(ps.max_depth), int(ps.max_tex_res), int(ps.min_tex_res), 0, float(ps.min_area_px),
                               (C.c_double * 3)(*[float(x) for x in ps.env]), float(ps.gamma))
        self.origin = np.ascontiguousarray(ps.origin, np.float64)
        self.vp_P = np.ascontiguousarray(ps.vp_P, np.float64)
        self.vp_UV = np.ascontiguousarray(ps.vp_UV, np.float64)


class AreCudaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes).  Every symbol include/are_cuda.h declares appears here (tests check that).
SIGNATURES = {
    "are_cuda_abi_version": (C.c_int, []),
    "are_cuda_device_count": (C.c_int, []),
    "are_cuda_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "are_cuda_create_multi": (C.c_int, [C.POINTER(_vp), _ip, C.c_int]),
    "are_cuda_group_info": (C.c_int, [_vp, _ip, _ip]),
    "are_cuda_destroy": (None, [_vp]),
    "are_cuda_last_error": (C.c_char_p, [_vp]),
    "are_cuda_set_stream": (C.c_int, [_vp, _vp]),
    "are_cuda_set_bvh_builder": (C.c_int, [_vp, C.c_int]),
    "are_cuda_get_commit_info": (C.c_int, [_vp, C.POINTER(CommitInfo)]),
    "are_cuda_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "are_cuda_set_build_threads": (None, [C.c_int]),
    "are_cuda_get_baked_cubin": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "are_cuda_add_texture": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int, C.c_int]),
    "are_cuda_add_material": (C.c_int, [_vp, C.c_int, _dp]),
    "are_cuda_add_triangle": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int]),
    "are_cuda_set_triangle_uv": (C.c_int, [_vp, C.c_int, _dp]),
    "are_cuda_add_quad": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_int]),
    "are_cuda_add_sphere": (C.c_int, [_vp, _dp, C.c_double, C.c_int, C.c_int]),
    "are_cuda_add_triangles": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _ip, _ip]),
    "are_cuda_add_spheres": (C.c_int, [_vp, C.c_int, _dp, _dp, _ip, _ip]),
    "are_cuda_clear": (C.c_int, [_vp]),
    "are_cuda_num_primitives": (C.c_int, [_vp]),
    "are_cuda_commit": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "are_cuda_update_triangles": (C.c_int, [_vp, C.c_int, _ip, _dp, _dp, _dp]),
    "are_cuda_update_spheres": (C.c_int, [_vp, C.c_int, _ip, _dp, _dp]),
    "are_cuda_refit": (C.c_int, [_vp, _dp]),
    "are_cuda_compile_probe": (C.c_int, [C.c_int, _dp, _dp, _dp, _ip]),
    "are_cuda_compile_probe_digest": (C.c_int, [C.c_int, _dp, _dp, _dp, _ip, C.POINTER(C.c_uint64)]),
    "are_cuda_compile_probe_forms": (C.c_int, [C.c_int, _dp, _dp, _dp, _ip]),
    "are_cuda_bake_probe": (C.c_int, [C.c_int, _dp, _dp, _dp, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_char_p]),
    "are_cuda_hit_batch": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp]),
    "are_cuda_scatter_batch": (C.c_int, [_vp, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, _ip]),
    "are_cuda_texture_batch": (C.c_int, [_vp, C.c_int, _ip, _dp, _dp, C.c_int, _dp]),
    "are_cuda_camera_rays": (C.c_int, [_vp, C.POINTER(Camera), C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _dp, _dp]),
    "are_cuda_plane_batch": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _ip, _dp]),
    "are_cuda_point_in_batch": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, _dp, _ip]),
    "are_cuda_material_reflect_batch": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _ip, _dp]),
    "are_cuda_philox_batch": (C.c_int, [_vp, C.c_int, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "are_cuda_render_device": (C.c_int, [_vp, C.POINTER(Camera), C.POINTER(RenderParams), _vp, C.POINTER(RenderStats), C.c_int]),
    "are_cuda_render": (C.c_int, [_vp, C.POINTER(Camera), C.POINTER(RenderParams), _fp, C.POINTER(RenderStats)]),
    "are_cuda_tonemap": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_uint8)]),
    "are_cuda_texture_paste": (C.c_int, [_vp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _ip]),
    "are_cuda_patch_render": (C.c_int, [_vp, C.POINTER(PatchSceneC), _dp, _dp, _dp, C.c_int, C.c_int, C.POINTER(PatchConfig), _dp,
                                        C.POINTER(C.c_uint8), C.POINTER(PatchStats)]),
    "are_cuda_patch_trace_texture": (C.c_int, [_vp, C.POINTER(PatchSceneC), _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(PatchConfig), _dp,
                                               _ip, C.POINTER(PatchStats)]),
    "are_cuda_patch_plan_probe": (C.c_int, [C.POINTER(PatchSceneC), _dp, _dp, _dp, C.c_int, C.c_int, C.POINTER(PatchConfig), C.POINTER(C.c_uint64)]),
    "are_cuda_write_ppm": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint8)]),
    "are_cuda_alloc_accum": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "are_cuda_zero_accum": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "are_cuda_download_accum": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _fp]),
    "are_cuda_free_accum": (C.c_int, [_vp, _vp]),
    "are_cuda_synchronize": (C.c_int, [_vp]),
    "are_cuda_measure_fp32_peak": (C.c_int, [_vp, _dp, _ip, _ip]),
    "are_cuda_measure_l2_peak": (C.c_int, [_vp, _dp, C.POINTER(C.c_uint64)]),
}


def load_library(path: str | None = None):
    """dlopen libare_b200.so and declare the signatures.  Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not built — run `python -c 'import __graft_entry__ as g; g.build()'` (there is no fallback path)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def patch_plan_probe(ps) -> dict:
    """Host-only planner probe (no GPU): node / texel / warp-triangle counts of a PatchScene's camera render."""
    lib = load_library()
    a = _PatchArgs(ps)
    out = (C.c_uint64 * 6)()
    st = lib.are_cuda_patch_plan_probe(C.byref(a.scene), _ptr(a.origin), _ptr(a.vp_P), _ptr(a.vp_UV), int(ps.width), int(ps.height), C.byref(a.cfg), out)
    if st < 0:
        raise AreCudaError(st, "patch_plan_probe")
    return dict(zip(("nodes", "node_texels", "ops", "levels", "ops_a", "ops_b"), [int(x) for x in out]))


def set_build_threads(n: int):
    """Host threads of the scene compiler, process-wide (0 = all hardware threads)."""
    load_library().are_cuda_set_build_threads(int(n))


def compile_probe(Q, u, v) -> dict:
    """Host-only scene-compiler probe (are_cuda_compile_probe): no GPU involved."""
    lib = load_library()
    Q, u, v = (np.ascontiguousarray(x, dtype=np.float64) for x in (Q, u, v))
    out = np.zeros(8, np.int32)
    digest = C.c_uint64(0)
    st = lib.are_cuda_compile_probe_digest(len(Q), Q.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp), out.ctypes.data_as(_ip),
                                           C.byref(digest))
    if st != ARE_OK:
        raise AreCudaError(st, "compile probe rejected the triangles")
    keys = ("hot_slots", "fused_pairs", "boxes", "bvh_nodes", "bvh_depth", "brute_quads", "brute_tris", "brute_boxes")
    d = dict(zip(keys, (int(x) for x in out)))
    d["digest"] = int(digest.value)
    return d


def compile_probe_forms(Q, u, v) -> dict:
    """Host-only probe of the lean form and the device-builder input (are_cuda_compile_probe_forms): no GPU involved."""
    lib = load_library()
    Q, u, v = (np.ascontiguousarray(x, dtype=np.float64) for x in (Q, u, v))
    out = np.zeros(8, np.int32)
    st = lib.are_cuda_compile_probe_forms(len(Q), Q.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp), out.ctypes.data_as(_ip))
    if st != ARE_OK:
        raise AreCudaError(st, "compile probe rejected the triangles")
    keys = ("lean_ok", "lean_records", "lean_open_boxes", "lbvh_items", "lbvh_slots", "lbvh_conservative", "host_bvh_nodes", "brute_boxes")
    return dict(zip(keys, (int(x) for x in out)))


def bake_probe(Q, u, v, packed=False, cubin_path=None, generic=False) -> str:
    """Host-only: the CUDA source generated for the lean form of a triangle scene (are_cuda_bake_probe); with cubin_path
    also the NVRTC-compiled sm_100a CUBIN.  No GPU involved."""
    lib = load_library()
    Q, u, v = (np.ascontiguousarray(x, dtype=np.float64) for x in (Q, u, v))
    n = C.c_uint64(0)
    buf = C.create_string_buffer(1 << 18)
    st = lib.are_cuda_bake_probe(len(Q), Q.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp), int(bool(packed)) | (2 if generic else 0), buf, len(buf),
                                 C.byref(n), os.fsencode(cubin_path) if cubin_path else None)
    if st != ARE_OK:
        raise AreCudaError(st, (lib.are_cuda_last_error(None) or b"").decode())
    return buf.value.decode()


def _d(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a, t=_dp):
    return a.ctypes.data_as(t)


class Context:
    """One are_cuda_ctx (one GPU).  Mirrors the header one-to-one; numpy arrays in, numpy arrays out."""

    def __init__(self, device=0):
        """device: an index (one GPU), or a sequence of indices (a multi-device group, are_cuda_create_multi: renders are
        sharded by sample range over the devices and summed on the first)."""
        self.lib = load_library()
        h = _vp()
        if isinstance(device, (list, tuple)):
            devs = np.ascontiguousarray(device, np.int32)
            st = self.lib.are_cuda_create_multi(C.byref(h), _ptr(devs, _ip), len(devs))
            device = int(devs[0]) if len(devs) else -1
        else:
            st = self.lib.are_cuda_create(C.byref(h), int(device))
        if st != ARE_OK:
            msg = self.lib.are_cuda_last_error(None)
            raise AreCudaError(st, (msg or b"").decode())
        self.h = h
        self.device = device

    # -- housekeeping ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.are_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st < 0:
            raise AreCudaError(st, (self.lib.are_cuda_last_error(self.h) or b"").decode())
        return st

    def group_info(self):
        n, p = C.c_int(0), C.c_int(0)
        self._ck(self.lib.are_cuda_group_info(self.h, C.byref(n), C.byref(p)))
        return n.value, bool(p.value)

    def set_stream(self, cuda_stream_handle: int):
        self._ck(self.lib.are_cuda_set_stream(self.h, _vp(cuda_stream_handle)))

    # -- scene ----------------------------------------------------------------------------------------
    def set_bvh_builder(self, builder: int):
        self._ck(self.lib.are_cuda_set_bvh_builder(self.h, int(builder)))

    def set_option(self, option: int, value: int):
        self._ck(self.lib.are_cuda_set_option(self.h, int(option), int(value)))

    def baked_cubin(self) -> bytes:
        """The CUBIN of the committed scene's scene-specialised kernel (raises when the scene has none)."""
        n = C.c_uint64(0)
        self._ck(self.lib.are_cuda_get_baked_cubin(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._ck(self.lib.are_cuda_get_baked_cubin(self.h, buf, n.value, C.byref(n)))
        return buf.raw

    def commit_info(self) -> CommitInfo:
        info = CommitInfo()
        self._ck(self.lib.are_cuda_get_commit_info(self.h, C.byref(info)))
        return info

    def add_texture(self, kind, params, rgb=None):
        p = _d(params, (8,))
        if rgb is not None:
            img = _d(rgb)
            hh, ww = img.shape[0], img.shape[1]
            return self._ck(self.lib.are_cuda_add_texture(self.h, int(kind), _ptr(p), _ptr(img), ww, hh))
        return self._ck(self.lib.are_cuda_add_texture(self.h, int(kind), _ptr(p), None, 0, 0))

    def add_material(self, kind, params):
        p = _d(params, (8,))
        return self._ck(self.lib.are_cuda_add_material(self.h, int(kind), _ptr(p)))

    def add_triangle(self, Q, u, v, mat, tex):
        Q, u, v = _d(Q, (3,)), _d(u, (3,)), _d(v, (3,))
        return self._ck(self.lib.are_cuda_add_triangle(self.h, _ptr(Q), _ptr(u), _ptr(v), int(mat), int(tex)))

    def set_triangle_uv(self, prim, uv):
        uv = _d(uv, (6,))
        return self._ck(self.lib.are_cuda_set_triangle_uv(self.h, int(prim), _ptr(uv)))

    def add_quad(self, Q, u, v, mat, tex):
        Q, u, v = _d(Q, (3,)), _d(u, (3,)), _d(v, (3,))
        return self._ck(self.lib.are_cuda_add_quad(self.h, _ptr(Q), _ptr(u), _ptr(v), int(mat), int(tex)))

    def add_sphere(self, c, r, mat, tex):
        c = _d(c, (3,))
        return self._ck(self.lib.are_cuda_add_sphere(self.h, _ptr(c), float(r), int(mat), int(tex)))

    def add_triangles(self, Q, u, v, mat, tex):
        Q, u, v = _d(Q), _d(u), _d(v)
        mat, tex = np.ascontiguousarray(mat, np.int32), np.ascontiguousarray(tex, np.int32)
        return self._ck(self.lib.are_cuda_add_triangles(self.h, len(Q), _ptr(Q), _ptr(u), _ptr(v), _ptr(mat, _ip), _ptr(tex, _ip)))

    def add_spheres(self, c, r, mat, tex):
        c, r = _d(c), _d(r)
        mat, tex = np.ascontiguousarray(mat, np.int32), np.ascontiguousarray(tex, np.int32)
        return self._ck(self.lib.are_cuda_add_spheres(self.h, len(c), _ptr(c), _ptr(r), _ptr(mat, _ip), _ptr(tex, _ip)))

    def clear(self):
        self._ck(self.lib.are_cuda_clear(self.h))

    def num_primitives(self):
        return self._ck(self.lib.are_cuda_num_primitives(self.h))

    def commit(self) -> int:
        n = C.c_uint64(0)
        self._ck(self.lib.are_cuda_commit(self.h, C.byref(n)))
        return n.value

    def update_triangles(self, ids, Q, u, v):
        ids = np.ascontiguousarray(ids, np.int32)
        Q, u, v = _d(Q), _d(u), _d(v)
        self._ck(self.lib.are_cuda_update_triangles(self.h, len(ids), _ptr(ids, _ip), _ptr(Q), _ptr(u), _ptr(v)))

    def update_spheres(self, ids, c, r):
        ids = np.ascontiguousarray(ids, np.int32)
        c, r = _d(c), _d(r)
        self._ck(self.lib.are_cuda_update_spheres(self.h, len(ids), _ptr(ids, _ip), _ptr(c), _ptr(r)))

    def refit(self) -> float:
        """Refit the device-built hierarchy to the primitives moved by update_*; returns the device time in ms."""
        ms = C.c_double(0)
        self._ck(self.lib.are_cuda_refit(self.h, C.byref(ms)))
        return ms.value

    # -- per-ray harness ------------------------------------------------------------------------------
    def hit_batch(self, Q, D, t_min=0.0, precision=64, traversal=0):
        Q, D = _d(Q), _d(D)
        n = len(Q)
        prim = np.empty(n, np.int32)
        t, P, N, uv = np.empty(n), np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 2))
        self._ck(self.lib.are_cuda_hit_batch(self.h, n, _ptr(Q), _ptr(D), float(t_min), int(precision), int(traversal),
                                             _ptr(prim, _ip), _ptr(t), _ptr(P), _ptr(N), _ptr(uv)))
        return prim, t, P, N, uv

    def scatter_batch(self, mat, tex, wi, N, P, uv, rnd, precision=64):
        mat, tex = np.ascontiguousarray(mat, np.int32), np.ascontiguousarray(tex, np.int32)
        wi, N, P, uv, rnd = _d(wi), _d(N), _d(P), _d(uv), _d(rnd)
        n = len(mat)
        wo, att, emit, alive = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3)), np.empty(n, np.int32)
        self._ck(self.lib.are_cuda_scatter_batch(self.h, n, _ptr(mat, _ip), _ptr(tex, _ip), _ptr(wi), _ptr(N), _ptr(P), _ptr(uv),
                                                 _ptr(rnd), int(precision), _ptr(wo), _ptr(att), _ptr(emit), _ptr(alive, _ip)))
        return wo, att, emit, alive

    def texture_batch(self, tex, uv, P, precision=64):
        tex = np.ascontiguousarray(tex, np.int32)
        uv, P = _d(uv), _d(P)
        n = len(tex)
        rgb = np.empty((n, 3))
        self._ck(self.lib.are_cuda_texture_batch(self.h, n, _ptr(tex, _ip), _ptr(uv), _ptr(P), int(precision), _ptr(rgb)))
        return rgb

    def camera_rays(self, cam: Camera, width, height, px, py, rnd, precision=64):
        px, py = np.ascontiguousarray(px, np.int32), np.ascontiguousarray(py, np.int32)
        rnd = _d(rnd)
        n = len(px)
        Q, D = np.empty((n, 3)), np.empty((n, 3))
        self._ck(self.lib.are_cuda_camera_rays(self.h, C.byref(cam), int(width), int(height), n, _ptr(px, _ip), _ptr(py, _ip),
                                               _ptr(rnd), int(precision), _ptr(Q), _ptr(D)))
        return Q, D

    def plane_batch(self, plane4, Q, D):
        """are::Plane::intersect_ray for n (plane, ray) pairs -> (hit int32[n], P float64[n,3])."""
        plane4, Q, D = _d(plane4), _d(Q), _d(D)
        n = len(Q)
        hit, P = np.zeros(n, np.int32), np.zeros((n, 3))
        self._ck(self.lib.are_cuda_plane_batch(self.h, n, _ptr(plane4), _ptr(Q), _ptr(D), _ptr(hit, _ip), _ptr(P)))
        return hit, P

    def point_in_batch(self, Q, u, v, pts):
        """are::Triangle(Q,u,v).point_in for n points -> int32[n]."""
        Q, u, v, pts = _d(Q), _d(u), _d(v), _d(pts)
        inside = np.zeros(len(pts), np.int32)
        self._ck(self.lib.are_cuda_point_in_batch(self.h, _ptr(Q), _ptr(u), _ptr(v), len(pts), _ptr(pts), _ptr(inside, _ip)))
        return inside

    def material_reflect_batch(self, kind, plane4, origin):
        """are::Material::reflect (kind 0 Diffuse, 1 Reflective) -> (ok int32[n], new_origin float64[n,3])."""
        plane4, origin = _d(plane4), _d(origin)
        n = len(origin)
        ok, out = np.zeros(n, np.int32), np.zeros((n, 3))
        self._ck(self.lib.are_cuda_material_reflect_batch(self.h, int(kind), n, _ptr(plane4), _ptr(origin), _ptr(ok, _ip), _ptr(out)))
        return ok, out

    def philox_batch(self, seed, counters):
        c = np.ascontiguousarray(counters, np.uint32)
        out = np.empty_like(c)
        self._ck(self.lib.are_cuda_philox_batch(self.h, len(c), int(seed), _ptr(c, C.POINTER(C.c_uint32)), _ptr(out, C.POINTER(C.c_uint32))))
        return out

    # -- rendering ------------------------------------------------------------------------------------
    def render_device(self, cam: Camera, params: RenderParams, accum_ptr: int, want_stats=False, count_tests=False):
        st = RenderStats() if want_stats else None
        self._ck(self.lib.are_cuda_render_device(self.h, C.byref(cam), C.byref(params), _vp(accum_ptr),
                                                 C.byref(st) if st is not None else None, int(bool(count_tests))))
        return st

    def render(self, cam: Camera, params: RenderParams, out: np.ndarray | None = None):
        """Host-buffer call: returns (accum float32 (H,W,3) of sample SUMS, RenderStats)."""
        if out is None:
            out = np.empty((params.height, params.width, 3), np.float32)
        st = RenderStats()
        self._ck(self.lib.are_cuda_render(self.h, C.byref(cam), C.byref(params), _ptr(out, _fp), C.byref(st)))
        return out, st

    def tonemap(self, accum_ptr: int, width, height, inv_spp, encoder=0, out: np.ndarray | None = None):
        """8-bit encode on the device -> (H, W, 3) uint8.  `out`: write into this array (e.g. a view of a P6 file image)."""
        if out is None:
            out = np.empty((height, width, 3), np.uint8)
        assert out.dtype == np.uint8 and out.size == height * width * 3 and out.flags.c_contiguous
        self._ck(self.lib.are_cuda_tonemap(self.h, _vp(accum_ptr), int(width), int(height), float(inv_spp), int(encoder),
                                           _ptr(out, C.POINTER(C.c_uint8))))
        return out

    def texture_paste(self, dst: np.ndarray, src: np.ndarray, corners):
        """In-place are::Texture::paste of src (h,w,3 float64) into dst; corners = (lt, rt, lb, rb) integer pixel pairs."""
        assert dst.dtype == np.float64 and dst.flags.c_contiguous
        src = np.ascontiguousarray(src, np.float64)
        c = np.ascontiguousarray(np.asarray(corners, np.int32).reshape(8))
        self._ck(self.lib.are_cuda_texture_paste(self.h, _ptr(dst), dst.shape[1], dst.shape[0], _ptr(src), src.shape[1], src.shape[0], _ptr(c, _ip)))
        return dst

    def patch_render(self, ps, want_rgb=True, want_rgb8=True):
        """The reference's patch-as-viewport renderer (experiments/rt10.cpp) for a scenes.PatchScene:
        -> (rgb float64 (H,W,3) | None, rgb8 uint8 (H,W,3) | None, PatchStats)."""
        a = _PatchArgs(ps)
        rgb = np.empty((ps.height, ps.width, 3), np.float64) if want_rgb else None
        rgb8 = np.empty((ps.height, ps.width, 3), np.uint8) if want_rgb8 else None
        st = PatchStats()
        self._ck(self.lib.are_cuda_patch_render(self.h, C.byref(a.scene), _ptr(a.origin), _ptr(a.vp_P), _ptr(a.vp_UV), int(ps.width), int(ps.height),
                                                C.byref(a.cfg), _ptr(rgb) if want_rgb else None,
                                                _ptr(rgb8, C.POINTER(C.c_uint8)) if want_rgb8 else None, C.byref(st)))
        return rgb, rgb8, st

    def patch_trace_texture(self, ps, origin, current, tex_w, tex_h, est_area_px=0.0):
        """Object::trace_texture for triangle `current` of a PatchScene seen from `origin` -> (texture (h,w,3) float64, PatchStats)."""
        a = _PatchArgs(ps)
        o = np.ascontiguousarray(origin, np.float64)
        out = np.empty(int(ps.max_tex_res) ** 2 * 3, np.float64)
        wh = np.zeros(2, np.int32)
        st = PatchStats()
        self._ck(self.lib.are_cuda_patch_trace_texture(self.h, C.byref(a.scene), _ptr(o), int(current), int(tex_w), int(tex_h), float(est_area_px),
                                                       C.byref(a.cfg), _ptr(out), _ptr(wh, _ip), C.byref(st)))
        return out[: int(wh[0]) * int(wh[1]) * 3].reshape(int(wh[1]), int(wh[0]), 3).copy(), st

    def write_ppm(self, path, rgb8: np.ndarray):
        rgb8 = np.ascontiguousarray(rgb8, np.uint8)
        h, w = rgb8.shape[:2]
        st = self.lib.are_cuda_write_ppm(os.fsencode(path), w, h, _ptr(rgb8, C.POINTER(C.c_uint8)))
        if st < 0:
            raise AreCudaError(st, f"cannot write {path}")

    def alloc_accum(self, width, height) -> int:
        p = _vp()
        self._ck(self.lib.are_cuda_alloc_accum(self.h, int(width), int(height), C.byref(p)))
        return p.value

    def zero_accum(self, ptr, width, height):
        self._ck(self.lib.are_cuda_zero_accum(self.h, _vp(ptr), int(width), int(height)))

    def download_accum(self, ptr, width, height):
        out = np.empty((height, width, 3), np.float32)
        self._ck(self.lib.are_cuda_download_accum(self.h, _vp(ptr), int(width), int(height), _ptr(out, _fp)))
        return out

    def free_accum(self, ptr):
        self._ck(self.lib.are_cuda_free_accum(self.h, _vp(ptr)))

    def synchronize(self):
        self._ck(self.lib.are_cuda_synchronize(self.h))

    def measure_l2_peak(self):
        g, n = C.c_double(0), C.c_uint64(0)
        self._ck(self.lib.are_cuda_measure_l2_peak(self.h, C.byref(g), C.byref(n)))
        return dict(gb_per_s=g.value, buffer_bytes=n.value)

    def measure_fp32_peak(self):
        t, sm, clk = C.c_double(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.are_cuda_measure_fp32_peak(self.h, C.byref(t), C.byref(sm), C.byref(clk)))
        return dict(tflops=t.value, sm_count=sm.value, sm_clock_khz=clk.value)
