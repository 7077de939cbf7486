"""Reference-independent pins for the NEW surface (SURVEY.md §8a last row: sphere, quad, lambertian, metal, dielectric,
light, noise, camera, spp loop — none of which the reference implements, so no reference artifact can pin them).

Every case here has a CLOSED-FORM answer that does not come from the CPU twin:

  white furnace     all albedos 1, uniform environment of radiance 1, no absorption  ->  every sample returns exactly 1
  convex Lambertian sphere / quad of albedo rho in a uniform environment of radiance 1  ->  L = rho exactly (throughput of
                    a cosine-sampled Lambertian bounce is the albedo; a convex body never sees itself)
  emitter           a DiffuseLight of radiance Le seen directly  ->  Le
  Schlick / Snell / total internal reflection written out independently in numpy
  cosine lobe       moments of the cosine-weighted hemisphere: E[cos] = 2/3, E[cos^2] = 1/2, E[x] = E[y] = 0, E[x^2] = 1/4
  Perlin noise      gradient noise vanishes on the integer lattice, so every octave of the turbulence does:
                    texture = 0.5 (1 + sin(scale z))
  camera            the centre of an even-sized image looks along the forward axis; corner rays subtend the field of view

`run_*` functions take a "sink factory" (the CUDA context in the GPU tests, the CPU twin in the CPU tests) so that the
same closed forms pin both sides.
"""
import numpy as np

from aurora_rendering_engine_b200 import capi, scenes


def furnace_scene(kind, width=96, height=64):
    """Uniform white environment, all-white non-absorbing surfaces.  kind: 'mixed' (spheres, quads, triangles; lambertian,
    mirror metal, glass — BVH or generic brute force), 'lean' (an open white box + blocks: the lean kernel's scene class)."""
    s = scenes.SceneDesc("furnace_" + kind, width=width, height=height, spp=16, max_depth=400, t_min=1e-3,
                         background_bottom=(1, 1, 1), background_top=(1, 1, 1))
    white = s.solid(1, 1, 1)
    lam = s.mat(scenes.MAT_LAMBERTIAN, -1)
    if kind == "lean":
        # open-front room of white Lambertian walls with two blocks (what the scene compiler turns into boxes + the lean form)
        flat = [((5, 0, 0), (0, 5, 0), (0, 0, 5)), ((0, 0, 0), (0, 5, 0), (0, 0, 5)), ((0, 0, 0), (5, 0, 0), (0, 0, 5)),
                ((5, 5, 5), (-5, 0, 0), (0, 0, -5)), ((0, 0, 5), (5, 0, 0), (0, 5, 0))]
        for Q, u, v in flat:
            Q, u, v = np.array(Q, float), np.array(u, float), np.array(v, float)
            s.quad_as_tris(Q, Q + u, Q + u + v, Q + v, lam, white)
        # The blocks FLOAT: where a closed block stands on the floor, a ray scattered from the floor within t_min of the
        # block's side passes through that side (the hit is closer than t_min) and is trapped inside the block until
        # max_depth — the classic t_min leak of this estimator (twin and kernels alike), which would cost a furnace sample.
        scenes._box_tris(s, np.array([0, 0, 0]), np.array([1.5, 3.0, 1.5]), lam, white, 15.0, (2.4, 0.3, 2.6))
        scenes._box_tris(s, np.array([0, 0, 0]), np.array([1.5, 1.5, 1.5]), lam, white, -18.0, (1.0, 0.25, 0.6))
        s.camera = dict(pos=(2.5, 2.5, -7.0), target=(2.5, 2.5, 0), up=(0, 1, 0), vfov_deg=40.0, focus_dist=5.0, jitter=1)
        return s
    mirror = s.mat(scenes.MAT_METAL, 0.0, -1)
    glass = s.mat(scenes.MAT_DIELECTRIC, 1.5)
    s.sphere((0, -100.5, -1), 100.0, lam, white)
    s.sphere((0, 0, -1.2), 0.5, lam, white)
    s.sphere((-1.1, 0, -1.0), 0.5, glass, white)
    s.sphere((1.1, 0, -1.0), 0.5, mirror, white)
    s.quad((-2.5, -0.5, -3.0), (5, 0, 0), (0, 2.0, 0), mirror, white)
    s.quad((-2.0, 1.2, -2.5), (1.5, 0, 0.4), (0, 0.1, 1.2), lam, white)
    s.tri((0.3, 0.6, -1.6), (0.9, 0.1, 0.0), (0.2, 0.7, 0.3), lam, white)
    s.tri((-0.9, 0.7, -1.4), (0.5, 0.0, 0.2), (0.1, 0.6, 0.0), glass, white)   # a glass sheet: refract() in and out through one surface
    scenes._box_tris(s, np.array([0, 0, 0]), np.array([0.5, 0.4, 0.5]), lam, white, 30.0, (-0.3, -0.35, -0.3))
    s.camera = dict(pos=(0, 0.4, 1.5), target=(0, 0, -1), up=(0, 1, 0), vfov_deg=60.0, focus_dist=2.5, defocus_angle_deg=1.0, jitter=1)
    return s


def albedo_scene(shape, rho=(0.2, 0.5, 0.8), width=64, height=64):
    """One convex Lambertian body of albedo rho in a uniform environment of radiance 1, seen head-on."""
    s = scenes.SceneDesc("albedo_" + shape, width=width, height=height, spp=8, max_depth=8, t_min=1e-3,
                         background_bottom=(1, 1, 1), background_top=(1, 1, 1))
    t = s.solid(*rho)
    lam = s.mat(scenes.MAT_LAMBERTIAN, -1)
    if shape == "sphere":
        s.sphere((0, 0, -3), 1.0, lam, t)
    elif shape == "quad":
        s.quad((-1, -1, -3), (2, 0, 0), (0, 2, 0), lam, t)
    else:
        s.tri((-2.0, -1.5, -3), (5, 0, 0), (0, 5, 0), lam, t)
    s.camera = dict(pos=(0, 0, 0), target=(0, 0, -1), up=(0, 1, 0), vfov_deg=60.0, focus_dist=1.0, jitter=1)
    return s


def emitter_scene(Le=(3.0, 2.0, 0.5), width=48, height=48):
    s = scenes.SceneDesc("emitter", width=width, height=height, spp=4, max_depth=8, t_min=1e-3,
                         background_bottom=(0, 0, 0), background_top=(0, 0, 0))
    t = s.solid(*Le)
    s.quad((-1, -1, -3), (2, 0, 0), (0, 2, 0), s.mat(scenes.MAT_DIFFUSE_LIGHT, -1, 1.0), t)
    s.camera = dict(pos=(0, 0, 0), target=(0, 0, -1), up=(0, 1, 0), vfov_deg=60.0, focus_dist=1.0, jitter=1)
    return s


def render(sink, sc, traversal=0, **over):
    cam = capi.make_camera(**sc.camera_args())
    par = capi.make_params(**sc.params_args(traversal=traversal, **over))
    if hasattr(sink, "commit"):
        sink.clear()
        sc.feed(sink)
        sink.commit()
        img, st = sink.render(cam, par)
    else:
        img, st = sc.feed(sink).render(cam, par)
    return np.asarray(img, np.float64) / par.sample_count, st


# ---- closed forms ----------------------------------------------------------------------------------------
def dielectric_expectation(wi, n, ior, u):
    """Independent statement of the dielectric: Snell's law in vector form + Schlick's approximation + TIR.
    wi unit incoming, n unit geometric normal (either side), u the uniform number.  -> (wo, reflected?)"""
    wi, n = np.asarray(wi, float), np.asarray(n, float)
    cos_i = -np.sum(wi * n, axis=1)
    front = cos_i > 0
    nf = np.where(front[:, None], n, -n)
    cos_i = np.abs(cos_i)
    eta = np.where(front, 1.0 / ior, ior)            # n1 / n2
    sin_t2 = eta * eta * (1.0 - cos_i * cos_i)       # Snell: sin(theta_t) = eta sin(theta_i)
    tir = sin_t2 > 1.0
    r0 = ((1.0 - ior) / (1.0 + ior)) ** 2            # same for n1/n2 and n2/n1
    schlick = r0 + (1.0 - r0) * (1.0 - cos_i) ** 5
    refl = tir | (u < schlick)
    cos_t = np.sqrt(np.maximum(0.0, 1.0 - sin_t2))
    refracted = eta[:, None] * wi + (eta * cos_i - cos_t)[:, None] * nf
    mirrored = wi + 2.0 * cos_i[:, None] * nf
    return np.where(refl[:, None], mirrored, refracted), refl, np.abs(u - schlick)


def unit(v):
    v = np.asarray(v, float)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)
