"""Pin the oracle (oracle/are_oracle.c) before trusting it.

1. against tests/golden/reference_vectors.npz — outputs of the REAL reference library recorded by
   tests/golden/make_golden.py (the reference ships no tests or golden vectors of its own);
2. against the known answers the survey recorded from its own probe of the reference;
3. live against oracle/_ref/libare_ref.so on fresh random inputs when that library is present;
4. Philox against the Random123 known-answer vectors.
All integer / decision results must be identical; fp64 values must be BIT-identical (same operations, same order).
"""
import os

import numpy as np
import pytest

from aurora_rendering_engine_b200 import scenes

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLD))


def same(a, b):
    """bit-identical including NaN positions"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_sizeof_and_epsilon(g):
    assert list(g["sizeof"]) == [24, 48, 32, 120, 16]
    assert g["geometry_epsilon"][0] == 1e-12


def test_vec3_ops_golden(oracle, g):
    a, b, s = g["vec_a"], g["vec_b"], g["vec_s"]
    for op in range(8):
        assert same(oracle.vec3_binary(op, a, b, s), g[f"vec_binary_{op}"]), f"binary op {op}"
    for op in range(4):
        assert same(oracle.vec3_scalar(op, a, b), g[f"vec_scalar_{op}"]), f"scalar op {op}"
    assert np.isnan(g["vec_binary_6"][0]).all()  # normalize(0) -> NaN vector (vec3.cpp:123-129)
    assert np.isnan(g["vec_binary_5"][1]).all()  # v/0 -> NaN vector (vec3.cpp:174-179)
    assert g["vec_scalar_3"][2] == 1.0            # near_zero


def test_reflect_refract_golden(oracle, g):
    assert same(oracle.reflect(g["rr_v"], g["rr_n"]), g["reflect"])
    assert same(oracle.refract(g["rr_v"], g["rr_n"], g["rr_eta"]), g["refract"])
    assert np.isfinite(g["refract"]).all()  # the reference never reports TIR (fabs under the sqrt, vec3.cpp:191)


def test_ray_plane_material_golden(oracle, g):
    oD, oA = oracle.ray(g["ray_Q"], g["ray_D"], g["ray_t"])
    assert same(oD, g["ray_outD"]) and same(oA, g["ray_at"])
    assert same(oracle.plane_from_point_normal(g["plane_p"], g["plane_n"]), g["plane4"])
    hit, P = oracle.plane_intersect(g["plane4"], g["plane_rayQ"], g["plane_rayD"])
    assert same(hit, g["plane_hit"]) and same(P, g["plane_P"])
    assert 0 < g["plane_hit"].sum() < len(g["plane_hit"])
    ok, out = oracle.material_reflect(1, 0.9, g["mat_planes"], g["mat_origin"])
    assert same(ok, g["mat_refl_ok"]) and same(out, g["mat_refl_out"])
    assert list(g["mat_refl_ok"][:4]) == [0, 0, 0, 0]
    ok, out = oracle.material_reflect(0, 0.0, g["mat_planes"], g["mat_origin"])
    assert same(ok, g["mat_diff_ok"]) and same(out, g["mat_diff_out"]) and not g["mat_diff_ok"].any()


def test_triangle_ctor_golden(oracle, g):
    for row, st, verts in zip(g["ctor_in"], g["ctor_status"], g["ctor_verts"]):
        ost, overts = oracle.triangle_ctor(row[0:3], row[3:6], row[6:9], int(row[9]))
        assert ost == st
        if st == 0:
            assert same(overts, verts)
    # row 7: |u x v| = 1e-10 per component is "near zero" for the reference (1e-8 threshold) -> rejected as collinear
    assert list(g["ctor_status"]) == [0, 1, 1, 1, 1, 1, 1, 1, 0]


def test_triangle_hit_golden(oracle, g):
    TQ, Tu, Tv = g["cornell_TQ"], g["cornell_Tu"], g["cornell_Tv"]
    hit, P = oracle.triset_hit_matrix(TQ, Tu, Tv, g["cornell_rayQ"][:256], g["cornell_rayD"][:256])
    assert same(hit, g["cornell_hit_matrix"]) and same(P, g["cornell_hit_P"])
    n, prim, t, PP = oracle.triset_closest_hit(TQ, Tu, Tv, g["cornell_rayQ"], g["cornell_rayD"])
    assert n == g["cornell_closest_n"][0]
    assert same(prim, g["cornell_closest_prim"]) and same(t, g["cornell_closest_t"]) and same(PP, g["cornell_closest_P"])
    assert same(oracle.triset_point_in(TQ, Tu, Tv, 3, g["pointin_pts"]), g["pointin"])
    assert 0 < g["pointin"].sum() < len(g["pointin"])


def test_twin_hit_batch_agrees_with_reference_closest_hit(oracle, g):
    """The fp64 twin's scene traversal (what the GPU is compared with) picks the reference's primitive and point."""
    sc = scenes.rt_cornell()
    osc = sc.feed(oracle.scene())
    prim, t, P, N, uv = osc.hit_batch(g["cornell_rayQ"], g["cornell_rayD"], 0.0)
    assert same(prim, g["cornell_closest_prim"])
    hit = prim >= 0
    assert same(P[hit], g["cornell_closest_P"][hit])
    # the reference recovers t as (P-Q)·D; Möller–Trumbore's own t agrees to rounding
    assert np.allclose(t[hit], g["cornell_closest_t"][hit], rtol=0, atol=1e-14)


def test_texture_io_golden(oracle, g, tmp_path):
    p = tmp_path / "t.ppm"
    assert oracle.texture_save(str(p), g["tex_rgb"]) == 1
    assert p.read_bytes() == g["tex_ppm_bytes"].tobytes()
    st, back = oracle.texture_load(str(p))
    assert st == 0 and same(back, g["tex_loaded"])
    assert oracle.texture_save(str(tmp_path / "t.png"), g["tex_rgb"]) == g["tex_save_bad_suffix"][0] == 0
    assert oracle.texture_load(str(tmp_path / "nope.ppm"))[0] == g["tex_load_missing"][0] == 1
    (tmp_path / "p3.ppm").write_bytes(b"P3\n1 1\n255\n0 0 0\n")
    assert oracle.texture_load(str(tmp_path / "p3.ppm"))[0] == g["tex_load_p3"][0] == 1
    (tmp_path / "short.ppm").write_bytes(b"P6\n2 2\n255\n" + bytes(5))
    assert oracle.texture_load(str(tmp_path / "short.ppm"))[0] == g["tex_load_short"][0] == 1
    assert list(g["tex_fill_ctor"]) == [0, 1, 1]
    # truncation, no gamma (src/texture.cpp:384-386)
    payload = g["tex_ppm_bytes"][len(b"P6\n9 6\n255\n"):]
    assert same(oracle.encode_linear(g["tex_rgb"]), payload)


def test_survey_known_answers(oracle):
    """SURVEY.md §8c, recorded from the survey's own probe of the reference."""
    TQ, Tu, Tv = np.zeros((1, 3)), np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]])
    hit, P = oracle.triset_hit_matrix(TQ, Tu, Tv, [[.25, .25, 1], [.75, .75, 1], [.25, .25, -1]], [[0, 0, -2]] * 3)
    assert list(hit[:, 0]) == [1, 0, 0] and np.array_equal(P[0, 0], [0.25, 0.25, 0.0])
    oD, _ = oracle.ray([[.25, .25, 1]], [[0, 0, -2]], [0.0])
    assert np.array_equal(oD[0], [0, 0, -1])
    assert list(oracle.triset_point_in(TQ, Tu, Tv, 0, [[.2, .2, 0], [.8, .8, 0]])) == [1, 0]
    pl = oracle.plane_from_point_normal([[0, 0, 0]], [[0, 0, 2]])
    assert np.array_equal(pl[0, :3], [0, 0, 1]) and pl[0, 3] == 0.0
    ok, out = oracle.material_reflect(1, 0.9, pl, [[1, 2, 3]])
    assert ok[0] == 1 and np.array_equal(out[0], [1, 2, -3])
    ok, _ = oracle.material_reflect(0, 0.0, pl, [[1, 2, 3]])
    assert ok[0] == 0
    assert np.array_equal(oracle.reflect([[1, -1, 0]], [[0, 1, 0]])[0], [1, 1, 0])
    uv = np.array([[1, -1, 0]]) / np.sqrt(2)
    r = oracle.refract(uv, [[0, 1, 0]], [1 / 1.5])[0]
    assert abs(r[0] - 0.47140452079103162) < 1e-16 and abs(r[1] + 0.88191710368819698) < 1e-16 and r[2] == 0
    r = oracle.refract(uv, [[0, 1, 0]], [3.0])[0]
    assert np.allclose(r, [2.12132, -1.87083, 0], atol=1e-5) and np.isfinite(r).all()  # TIR case: finite, not NaN


def test_philox_known_answers(oracle):
    """Random123 kat_vectors, philox4x32 10 rounds."""
    def one(ctr, key):
        seed = key[0] | (key[1] << 32)
        return [int(x) for x in oracle.philox(seed, np.array([ctr], np.uint32))[0]]
    assert one([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert one([0xffffffff] * 4, [0xffffffff, 0xffffffff]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert one([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---- live against the real reference (build container, or wherever oracle/_ref travelled) ------------------
def test_live_reference_random(oracle, reference):
    rng = np.random.RandomState(99)
    n = 20_000
    a, b, s = rng.uniform(-5, 5, (n, 3)), rng.uniform(-5, 5, (n, 3)), rng.uniform(-3, 3, n)
    for op in range(8):
        assert same(oracle.vec3_binary(op, a, b, s), reference.vec3_binary(op, a, b, s))
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    nn = rng.normal(size=(n, 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    eta = rng.uniform(0.3, 3.0, n)
    assert same(oracle.reflect(v, nn), reference.reflect(v, nn))
    assert same(oracle.refract(v, nn, eta), reference.refract(v, nn, eta))
    sc = scenes.cornell_box(as_quads=False)
    TQ = np.stack([t[0] for t in sc.tris]); Tu = np.stack([t[1] for t in sc.tris]); Tv = np.stack([t[2] for t in sc.tris])
    Q, D = rng.uniform(50, 500, (n, 3)), rng.uniform(-1, 1, (n, 3))
    ts = reference.triset(TQ, Tu, Tv)
    rn, rp, rt, rP = ts.closest_hit(Q, D)
    on, op_, ot, oP = oracle.triset_closest_hit(TQ, Tu, Tv, Q, D)
    assert rn == on and same(rp, op_) and same(rt, ot) and same(rP, oP)
    pts = rng.uniform(0, 555, (2000, 3))
    pts[:, 1] = 0.0
    assert same(ts.point_in(6, pts), oracle.triset_point_in(TQ, Tu, Tv, 6, pts))
    ts.close()


def test_survey_million_ray_hit_count(oracle, reference):
    """1 M rays vs the 34-triangle Cornell set -> the oracle and the reference agree on every ray.
    (The survey's exact generator state is not recorded, so the count is compared live, not against 833 673.)"""
    sc = scenes.rt_cornell()
    TQ = np.stack([t[0] for t in sc.tris]); Tu = np.stack([t[1] for t in sc.tris]); Tv = np.stack([t[2] for t in sc.tris])
    rng = np.random.RandomState(12345)
    Q, D = 0.3 * rng.uniform(-1, 1, (1_000_000, 3)), rng.uniform(-1, 1, (1_000_000, 3))
    ts = reference.triset(TQ, Tu, Tv)
    rn, rp, rt, rP = ts.closest_hit(Q, D)
    on, op_, ot, oP = oracle.triset_closest_hit(TQ, Tu, Tv, Q, D)
    ts.close()
    assert rn == on and same(rp, op_) and same(rt, ot)
    assert 0.80 < rn / 1e6 < 0.87  # the survey measured 83.4 % with its generator


def test_cpu_twin_tree_equals_linear_scan(oracle):
    """The CPU twin prunes large scenes with an AABB tree; it must return exactly what its linear scan over the hittable
    list returns (same primitive — ties to the lower index —, same t, point, normal, uv).  The scan result is assembled
    from sub-scenes small enough (<= 256 primitives) to be scanned linearly."""
    import copy
    from aurora_rendering_engine_b200 import scenes
    sc = scenes.stress(n_prims=3000, width=8, height=8)
    # duplicate a few primitives so that exact ties in t exist
    sc.order = sc.order + sc.order[:40]
    osc = sc.feed(oracle.scene())
    rng = np.random.RandomState(1)
    n = 20000
    Q = rng.uniform(-6, 6, (n, 3))
    D = rng.normal(size=(n, 3))
    D[:50, 0] = 0.0
    D[50:100, 1] = 0.0
    prim, t, P, N, uv = osc.hit_batch(Q, D, t_min=1e-3)
    best_t, best_p = np.full(n, np.inf), np.full(n, -1)
    chunk = 250
    for c0 in range(0, len(sc.order), chunk):
        sub = copy.copy(sc)
        sub.order = sc.order[c0:c0 + chunk]
        p2, t2, *_ = sub.feed(oracle.scene()).hit_batch(Q, D, t_min=1e-3)
        m = (p2 >= 0) & (t2 < best_t)   # strict: an equal t later in the list does not replace an earlier one
        best_t[m], best_p[m] = t2[m], p2[m] + c0
    assert (prim >= 0).sum() > 300
    assert np.array_equal(best_p, prim)
    assert np.array_equal(best_t[prim >= 0], t[prim >= 0])


def test_rt_gather_restatement_vs_the_real_program_with_diffuse_walls(oracle, tmp_path):
    """rt.cpp's cosine gather + Russian roulette (rt.cpp:278-329, :50-55) is dead code in the scene
    the reference ships (all surfaces MAT_METAL).  oracle/_ref/rt_ref_diffuse_counted is the same program with the five
    walls made MAT_DIFFUSE by sed; it draws from mt19937, the restatement from Philox, so the pin is statistical: the
    true ray count per pixel, and the mean of ENCODED single-sample images (the program encodes one noisy estimate per
    pixel) over 16x16 blocks — noise falls with the block size, a systematic difference would not."""
    from aurora_rendering_engine_b200 import capi
    from oracle_binding import RT_REF_DIFFUSE_COUNTED, block_mean, psnr, run_rt_reference
    if not os.path.exists(RT_REF_DIFFUSE_COUNTED):
        pytest.skip("oracle/_ref/rt_ref_diffuse_counted not present")
    ref, ref_rays = run_rt_reference(RT_REF_DIFFUSE_COUNTED, tmp_path, seed=3)
    sc = scenes.rt_cornell(diffuse_walls=True)
    osc = sc.feed(oracle.scene())
    cam = capi.make_camera(**sc.camera_args())
    K, G, rays = 4, np.zeros((512, 512, 3)), 0
    for k in range(K):
        one, st = osc.render(cam, capi.make_params(**sc.params_args(sample_count=1, sample_begin=k)))
        G += oracle.encode_gamma22(one.astype(np.float32)).reshape(512, 512, 3) / 255.0
        rays += int(st.rays)
    G /= K
    assert abs(rays / K - ref_rays) < 2e-3 * ref_rays, (rays / K, ref_rays)   # 43.5 rays per pixel (34.2 in the all-mirror scene)
    assert ref_rays / 262144 > 40
    assert abs(G.mean() - ref.mean()) < 3e-3 * ref.mean(), (G.mean(), ref.mean())
    p16 = psnr(block_mean(G, 16), block_mean(ref, 16))
    assert p16 >= 48.0, p16   # measured 52-54 dB; per pixel it is 30 dB (the reference's own run-to-run noise: 28 dB)
