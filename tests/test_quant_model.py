"""Numerical model (numpy, no GPU) of the quantised BVH nodes: lbvh.cu's k_quant_grid / k_quant_nodes arithmetic and
intersect.cuh's ray_slopes_q / bvhq_step decode, restated in fp32 — the claim the traversal rests on is that a ray
which meets the REAL box [c - h, c + h] always passes the test on its 16-bit box, rounding included, for ray origins
within 64 grid extents of the grid (dev_types.h: BvhNodeQ; profiles/r02_l1_pipe.md)."""
import numpy as np
import pytest

f32 = np.float32


def fma32(a, b, c):
    """fmaf for float32 arrays: the product of two floats is exact in double, the sum is rounded once more on the way back
    (double rounding differs from a true fma in ~1e-9 of the cases, by one ulp)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def make_grid(lo, hi):
    """k_quant_grid: root box -> (grid lo, step) per axis, as floats."""
    widest = float(np.max(hi - lo))
    glo, step = np.zeros(3, f32), np.zeros(3, f32)
    for k in range(3):
        mag = max(abs(lo[k]), abs(hi[k]))
        ext = max(max(hi[k] - lo[k], 1e-4 * widest), max(mag * 1e-4, 1e-20))
        step[k] = f32(ext / 65520.0)
        glo[k] = f32(lo[k] - 4.0 * float(step[k]))
    return glo, step


def quantise(c, h, glo, step):
    """k_quant_nodes: (centre, half-extent) floats -> (q_lo, q_hi) per axis, in double."""
    c, h = c.astype(np.float64), h.astype(np.float64)
    inv = 1.0 / step.astype(np.float64)
    a = np.clip(np.floor((c - h - glo) * inv) - 1.0, 0, 65535)
    b = np.clip(np.ceil((c + h - glo) * inv) + 1.0, 0, 65535)
    return a.astype(np.uint32), b.astype(np.uint32)


def slab_q(qlo, qhi, glo, step, o, d, tmin):
    """ray_slopes_q + bvhq_step for one box per ray, fp32."""
    big = f32(1e18)
    with np.errstate(divide="ignore"):
        inv = np.where(np.abs(d) > f32(1e-18), f32(1.0) / d, np.where(d < 0, -big, big)).astype(f32)
    A = (step[None, :] * inv).astype(f32)
    B = fma32(np.full_like(A, -8388608.0), A, ((glo[None, :] - o).astype(f32) * inv).astype(f32))
    pos = inv >= 0
    near = np.where(pos, qlo, qhi).astype(np.uint32)
    far = np.where(pos, qhi, qlo).astype(np.uint32)
    as_float = lambda q: (np.uint32(0x4B000000) | q).view(f32)  # the byte permute: 2^23 + q
    tn = fma32(as_float(near), A, B)
    tf = fma32(as_float(far), A, B)
    t_near = np.maximum(np.max(tn, axis=1), f32(tmin))
    t_far = np.min(tf, axis=1)
    return t_near <= t_far


def slab_exact(c, h, o, d, tmin):
    """The real box, in double: does the ray meet it beyond tmin?"""
    c, h, o, d = (x.astype(np.float64) for x in (c, h, o, d))
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (c - h - o) / d, (c + h - o) / d
    lo_t, hi_t = np.minimum(t0, t1), np.maximum(t0, t1)
    par = d == 0  # parallel to a slab: inside it or not
    inside = (o >= c - h) & (o <= c + h)
    lo_t = np.where(par, np.where(inside, -np.inf, np.inf), lo_t)
    hi_t = np.where(par, np.where(inside, np.inf, -np.inf), hi_t)
    return np.maximum(lo_t.max(axis=1), tmin) <= hi_t.min(axis=1)


@pytest.mark.parametrize("case", ["cube", "flat_z", "far_from_origin", "long_x", "camera_48_extents_away"])
def test_quantised_box_contains_the_real_box_for_every_ray(case):
    rng = np.random.RandomState(["cube", "flat_z", "far_from_origin", "long_x", "camera_48_extents_away"].index(case) + 17)
    n = 400_000
    root_lo, root_hi = np.array([-50.0, -50.0, -50.0]), np.array([50.0, 50.0, 50.0])
    if case == "flat_z":
        root_lo[2] = root_hi[2] = 0.0
    if case == "far_from_origin":
        root_lo += 1e4; root_hi += 1e4
    if case == "long_x":
        root_lo[0], root_hi[0] = -5e3, 5e3
    glo, step = make_grid(root_lo, root_hi)
    ext = root_hi - root_lo
    c = (root_lo + rng.uniform(0.02, 0.98, (n, 3)) * ext).astype(f32)
    h = (np.exp(rng.uniform(np.log(1e-3), np.log(5.0), (n, 3))) * np.minimum(1.0, np.maximum(ext, 1e-9) / 100.0)).astype(f32)
    if case == "flat_z":
        c[:, 2] = 0.0; h[:, 2] = 0.0
    # keep the boxes inside the root box (children of the root are)
    h = np.minimum(h, np.minimum(c - root_lo.astype(f32), root_hi.astype(f32) - c)).astype(f32)
    h = np.maximum(h, 0).astype(f32)
    qlo, qhi = quantise(c, h, glo, step)
    assert (qlo >= 1).all() and (qhi <= 65534).all() and (qlo < qhi).all()
    # rays aimed at (or just past) the box from near and far, a tenth of them with zero direction components
    reach = 48.0 if case == "camera_48_extents_away" else 2.0
    o = (c.astype(np.float64) + rng.normal(size=(n, 3)) * np.maximum(ext.max(), 1.0) * rng.uniform(0.0, reach, (n, 1))).astype(f32)
    aim = c.astype(np.float64) + rng.uniform(-1.3, 1.3, (n, 3)) * np.maximum(h, 1e-3)
    d = aim - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    d = d.astype(f32)
    z = rng.uniform(size=n) < 0.1
    d[z, rng.randint(0, 3, z.sum())] = 0.0
    tmin = 1e-3
    exact = slab_exact(c, h, o, d, tmin)
    quant = slab_q(qlo, qhi, glo, step, o, d, tmin)
    assert exact.mean() > 0.2
    missed = exact & ~quant
    assert missed.sum() == 0, f"{missed.sum()} rays meet the real box and miss its quantised box"
    extra = (quant & ~exact).sum() / max(1, (~exact).sum())
    print(f"[{case}] rays meeting the box {exact.mean():.3f}; quantised-only hits among the misses {extra:.4f}")
