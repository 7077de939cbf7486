// shade.cuh — surface resolution, texture evaluation, sampling and material scatter on the device.
// One templated source: T = double for the per-ray harness (checked against the reference / oracle to ~1e-12),
// T = float inside the render kernels.  Reference arithmetic used (paths under /root/reference):
//   metal       are::reflect                          src/basic/vec3.cpp:182-184
//   dielectric  are::refract (fabs quirk kept)        src/basic/vec3.cpp:188-199
//   lambertian  cosine gather + rotateToHemisphere    experiments/rt.cpp:285-289, 50-55
//   fuzz / AO   uniform sphere, z = 1-2v              experiments/rt.cpp:226-231
//   uv checker  CheckerTexture::at                    experiments/rt.cpp:98-101
//   bary -> uv  Tri::bary2uv                          experiments/rt.cpp:139-142
//   camera      render() ray set-up                   experiments/rt.cpp:339-343,364-366
#pragma once
#include "dev_types.h"
#include "philox.cuh"
#include "vec.cuh"

namespace areb {

template <typename T> __device__ __forceinline__ T mparam(const MaterialRec &m, int i);
template <> __device__ __forceinline__ double mparam<double>(const MaterialRec &m, int i) { return m.p[i]; }
template <> __device__ __forceinline__ float mparam<float>(const MaterialRec &m, int i) { return m.pf[i]; }
template <typename T> __device__ __forceinline__ T tparam(const TextureRec &t, int i);
template <> __device__ __forceinline__ double tparam<double>(const TextureRec &t, int i) { return t.p[i]; }
template <> __device__ __forceinline__ float tparam<float>(const TextureRec &t, int i) { return t.pf[i]; }

// fp32 normalisation = one MUFU.RSQ (rsqrt(0)=inf -> NaN vector, same convention as the reference)
__device__ __forceinline__ V3<double> nrm(V3<double> a) { return normalized(a); }
__device__ __forceinline__ V3<float> nrm(V3<float> a) { return rsqrtf(len2(a)) * a; }

// A combination lx*T + ly*B + lz*N of an orthonormal frame with lx^2 + ly^2 + lz^2 = 1 is already unit length up to
// rounding: fp64 normalises anyway (literal to the oracle), fp32 keeps the 1e-7 deviation and saves a MUFU + 6 ops.
__device__ __forceinline__ V3<double> renorm(V3<double> a) { return normalized(a); }
__device__ __forceinline__ V3<float> renorm(V3<float> a) { return a; }

template <typename T> struct Pi { static constexpr T value = T(3.14159265358979323846); };

// ---- textures -----------------------------------------------------------------------------------------
// table layout (scene.cpp): [256 x float4 gradients][3 x 256 int permutations][768 double gradients]
template <typename T>
__device__ T perlin_noise(const float *tab, V3<T> p) {
	const int *perm = reinterpret_cast<const int *>(tab + 1024);
	const float4 *grad = reinterpret_cast<const float4 *>(tab);
	T fx = floor_t(p.x), fy = floor_t(p.y), fz = floor_t(p.z);
	T u = p.x - fx, v = p.y - fy, w = p.z - fz;
	int i = (int)fx, j = (int)fy, k = (int)fz;
	T uu = u * u * (T(3) - T(2) * u), vv = v * v * (T(3) - T(2) * v), ww = w * w * (T(3) - T(2) * w), acc = T(0);
#pragma unroll
	for (int di = 0; di < 2; ++di)
#pragma unroll
		for (int dj = 0; dj < 2; ++dj)
#pragma unroll
			for (int dk = 0; dk < 2; ++dk) {
				int g = perm[(i + di) & 255] ^ perm[256 + ((j + dj) & 255)] ^ perm[512 + ((k + dk) & 255)];
				const float4 gq = __ldg(grad + g);  // one 16-byte load per lattice corner
				V3<T> gv = mk<T>(T(gq.x), T(gq.y), T(gq.z));
				V3<T> wv = mk<T>(u - T(di), v - T(dj), w - T(dk));
				acc += (di ? uu : T(1) - uu) * (dj ? vv : T(1) - vv) * (dk ? ww : T(1) - ww) * dot(gv, wv);
			}
	return acc;
}
// NOISE tables are kept in double precision for the fp64 harness (second half of the block)
template <typename T>
__device__ T perlin_noise_tab(const DevScene &sc, const TextureRec &t, V3<T> p);
template <>
__device__ __forceinline__ float perlin_noise_tab<float>(const DevScene &sc, const TextureRec &t, V3<float> p) {
	return perlin_noise<float>(sc.tex_data + t.data_off, p);
}
template <>
__device__ inline double perlin_noise_tab<double>(const DevScene &sc, const TextureRec &t, V3<double> p) {
	const float *tab = sc.tex_data + t.data_off;
	const int *perm = reinterpret_cast<const int *>(tab + 1024);
	const double *gd = reinterpret_cast<const double *>(tab + 1792);
	double fx = floor(p.x), fy = floor(p.y), fz = floor(p.z);
	double u = p.x - fx, v = p.y - fy, w = p.z - fz;
	int i = (int)fx, j = (int)fy, k = (int)fz;
	double uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w), acc = 0.0;
	for (int di = 0; di < 2; ++di)
		for (int dj = 0; dj < 2; ++dj)
			for (int dk = 0; dk < 2; ++dk) {
				int g = perm[(i + di) & 255] ^ perm[256 + ((j + dj) & 255)] ^ perm[512 + ((k + dk) & 255)];
				V3<double> gv = mk<double>(gd[3 * g], gd[3 * g + 1], gd[3 * g + 2]);
				V3<double> wv = mk<double>(u - di, v - dj, w - dk);
				acc += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) * dot(gv, wv);
			}
	return acc;
}

// ---- warp-cooperative turbulence --------------------------------------------------------------------------------
// A surface hit needs sum_{o<7} 2^-o noise(2^o P) — seven octaves of eight lattice corners, ~770 instructions — and in
// the render loop only the few lanes whose ray ended on a noise-textured surface need it (5.7 of 32 on BASELINE config
// 2), so evaluated lane by lane it is the most expensive and the emptiest code of the kernel.  Here the warp shares it:
// the requesters queue (P, texture) in the warp's 32 shared-memory slots in ballot order, lane l works on queue entry
// 4 * round + l/8 and octave l%8 (the eighth lane of a group idles), i.e. ONE octave per lane with all 32 lanes busy,
// three shuffle-adds form each entry's sum and the group's first lane posts it for the requester to pick up.
// (First form, round 2: the four requesters of a round were found with find-first-set chains on the ballot and their
// points fetched by four shuffles — ~45 instructions of bookkeeping per round beside the ~105 of the octave itself;
// the queue needs ~10.  Same arithmetic, same sums.)
// Called by all 32 lanes in converged code.  queue_sb / sums_sb: 32-bit shared-memory addresses of the warp's 32 queue
// entries (float4) and 32 sums (float) — explicit ld / st.shared on a base that lives in one register: through generic
// pointers the compiler re-derives the shared window and the warp index inside the loop (S2R SR_TID.X, S2R SR_CgaCtaId,
// LEA: 29 S2R in the textured kernel, 2 % of its warp-state samples).
// Returns the sum to the lanes that set `need`; 0 elsewhere.
__device__ __forceinline__ float turbulence_coop(const DevScene &sc, bool need, int tex_id, V3<float> P, uint32_t queue_sb, uint32_t sums_sb) {
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31, slot = lane >> 3, oct = lane & 7;
	const unsigned m = __ballot_sync(full, need);
	const int n = __popc(m);
	const int my_rank = __popc(m & ((1u << lane) - 1u));  // this lane is the my_rank-th requester (if it is one)
	if (need)
		asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(queue_sb + 16u * my_rank), "f"(P.x), "f"(P.y), "f"(P.z), "f"(__int_as_float(tex_id)) : "memory");
	__syncwarp();
	for (int base = 0; base < n; base += 4) {  // warp-uniform
		const int e = base + slot;
		float v = 0.0f;
		if (e < n && oct < 7) {
			float4 q;
			asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(queue_sb + 16u * e) : "memory");
			const float up = (float)(1 << oct), w = 1.0f / up;  // exact powers of two: the same points and weights as repeated doubling / halving
			v = w * perlin_noise<float>(sc.tex_data + sc.texs[__float_as_int(q.w)].data_off, mk<float>(q.x * up, q.y * up, q.z * up));
		}
		v += __shfl_xor_sync(full, v, 1);
		v += __shfl_xor_sync(full, v, 2);
		v += __shfl_xor_sync(full, v, 4);
		if (oct == 0 && e < n) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sums_sb + 4u * e), "f"(v) : "memory");
	}
	__syncwarp();
	float result = 0.0f;
	if (need) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(result) : "r"(sums_sb + 4u * my_rank) : "memory");
	return result;
}

// turb: when non-null, the turbulence sum of a TK_NOISE texture at P already formed by turbulence_coop
template <typename T>
__device__ V3<T> tex_eval(const DevScene &sc, int id, T u, T v, V3<T> P, const T *turb = nullptr) {
	const TextureRec &t = sc.texs[id];
	switch (t.kind) {
	case TK_SOLID:
		return mk<T>(tparam<T>(t, 0), tparam<T>(t, 1), tparam<T>(t, 2));
	case TK_CHECKER_UV: {
		int xx = (int)floor_t(u * tparam<T>(t, 0)), yy = (int)floor_t(v * tparam<T>(t, 0));
		return ((xx + yy) % 2 == 0) ? mk<T>(tparam<T>(t, 1), tparam<T>(t, 2), tparam<T>(t, 3)) : mk<T>(tparam<T>(t, 4), tparam<T>(t, 5), tparam<T>(t, 6));
	}
	case TK_CHECKER_3D: {
		T inv = T(1) / tparam<T>(t, 0);
		int xi = (int)floor_t(inv * P.x), yi = (int)floor_t(inv * P.y), zi = (int)floor_t(inv * P.z);
		return ((xi + yi + zi) % 2 == 0) ? mk<T>(tparam<T>(t, 1), tparam<T>(t, 2), tparam<T>(t, 3)) : mk<T>(tparam<T>(t, 4), tparam<T>(t, 5), tparam<T>(t, 6));
	}
	case TK_NOISE: {
		T acc = T(0), weight = T(1);
		V3<T> q = P;
		if (turb) acc = *turb;
		else
			for (int i = 0; i < 7; ++i) {
				acc += weight * perlin_noise_tab<T>(sc, t, q);
				weight *= T(0.5);
				q = T(2) * q;
			}
		T g = T(0.5) * (T(1) + sin_t(tparam<T>(t, 0) * P.z + T(10) * abs_t(acc)));
		return mk<T>(g, g, g);
	}
	default: {
		T uc = min_t(T(1), max_t(T(0), u)), vc = T(1) - min_t(T(1), max_t(T(0), v));
		int i = (int)(uc * T(t.w)), j = (int)(vc * T(t.h));
		i = i > t.w - 1 ? t.w - 1 : i;
		j = j > t.h - 1 ? t.h - 1 : j;
		const float *c = sc.tex_data + t.data_off + ((long long)j * t.w + i) * 3;
		return mk<T>(T(c[0]), T(c[1]), T(c[2]));
	}
	}
}

// ---- sampling -----------------------------------------------------------------------------------------
// rotateToHemisphere's frame (rt.cpp:50-53).  frame(-n) = (-tangent, bitangent): `up` depends on |n.z| only, the cross
// products are odd / even in n and every operation commutes exactly with negation.
template <typename T>
__device__ __forceinline__ void tangent_frame(V3<T> n, V3<T> &tangent, V3<T> &bitangent) {
	V3<T> up = abs_t(n.z) < T(0.999) ? mk<T>(T(0), T(0), T(1)) : mk<T>(T(1), T(0), T(0));
	tangent = nrm(cross(n, up));
	bitangent = cross(n, tangent);
}
template <typename T>
__device__ __forceinline__ V3<T> cosine_dir_in_frame(V3<T> n, V3<T> tangent, V3<T> bitangent, T r1, T r2) {
	T sn, cs;
	sincos2pi_t(r1, &sn, &cs);
	T r2s = sqrt_t(r2);
	T lx = r2s * cs, ly = r2s * sn;
	// rt.cpp:54 rebuilds lz from lx, ly; lx^2 + ly^2 = r2 exactly, so sqrt(1 - r2) is the same quantity without
	// the sin/cos rounding amplified at grazing directions
	T lz = sqrt_t(max_t(T(0), T(1) - r2));
	return mad(lz, n, mad2(lx, tangent, ly, bitangent));
}
template <typename T>
__device__ __forceinline__ V3<T> cosine_dir(V3<T> n, T r1, T r2) {
	V3<T> tangent, bitangent;
	tangent_frame(n, tangent, bitangent);
	return cosine_dir_in_frame(n, tangent, bitangent, r1, r2);
}
template <typename T>
__device__ __forceinline__ V3<T> sphere_dir(T r0, T r1) {
	T z = T(1) - T(2) * r0, rxy = sqrt_t(max_t(T(0), T(1) - z * z)), sn, cs;
	sincos2pi_t(r1, &sn, &cs);
	return mk<T>(rxy * cs, rxy * sn, z);
}

template <typename T>
__device__ __forceinline__ V3<T> cosine_dir_framed(V3<T> nf, bool front, unsigned frame_addr, T r1, T r2);
template <>
__device__ __forceinline__ V3<double> cosine_dir_framed<double>(V3<double> nf, bool, unsigned, double r1, double r2) { return cosine_dir<double>(nf, r1, r2); }
template <>
__device__ __forceinline__ V3<float> cosine_dir_framed<float>(V3<float> nf, bool front, unsigned frame_addr, float r1, float r2) {
	float4 t, b;
	asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(frame_addr));
	asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(frame_addr + 16u));
	const V3<float> tg = front ? mk<float>(t.x, t.y, t.z) : mk<float>(-t.x, -t.y, -t.z);
	return cosine_dir_in_frame<float>(nf, tg, mk<float>(b.x, b.y, b.z), r1, r2);
}

__device__ __forceinline__ int mat_texture(const MaterialRec &m, int prim_tex) {
	int o = -1;
	if (m.kind == MK_LAMBERTIAN || m.kind == MK_LIGHT) o = (int)m.pf[0];
	else if (m.kind == MK_METAL) o = (int)m.pf[1];
	return o >= 0 ? o : prim_tex;
}

// Direction sampling shared by the general path and the SHADE_FAST path of the render loop.
// lobe: MK_DIELECTRIC (p0 = ior), MK_METAL (p0 = fuzz), MK_REFLECTIVE = perfect mirror, anything else = cosine lobe.
// wi unit incoming, Ng geometric unit normal (either side). Returns alive; wo is unit length.
// FRAME_ADDR != 0 (fp32 lean kernel): shared-memory address of the frame of +Ng, precomputed once per shading record
// by the very routine above — (tangent, 0) (bitangent, 0) as two float4.
template <typename T>
__device__ __forceinline__ bool scatter_dir(int lobe, T p0, V3<T> wi, V3<T> Ng, Rnd4<T> r, V3<T> &wo, unsigned frame_addr = 0u) {
	const bool front = dot(wi, Ng) < T(0);
	const V3<T> nf = front ? Ng : -Ng;
	if (lobe == MK_DIELECTRIC) {
		T ri = front ? T(1) / p0 : p0;
		T cos_t = min_t(dot(-wi, nf), T(1)), sin_t = sqrt_t(max_t(T(0), T(1) - cos_t * cos_t));
		T r0 = (T(1) - ri) / (T(1) + ri);
		r0 = r0 * r0;
		T x = T(1) - cos_t, x2 = x * x;
		T schlick = r0 + (T(1) - r0) * (x2 * x2 * x);
		V3<T> d = (ri * sin_t > T(1) || schlick > r.x) ? reflect(wi, nf) : refract(wi, nf, ri);
		wo = nrm(d);
		return true;
	}
	if (lobe == MK_METAL) {
		V3<T> d = nrm(reflect(wi, nf)) + p0 * sphere_dir<T>(r.x, r.y);
		if (!(dot(d, nf) > T(0))) return false;
		wo = nrm(d);
		return true;
	}
	if (lobe == MK_REFLECTIVE) {
		wo = nrm(reflect(wi, nf));
		return true;
	}
	if (frame_addr != 0u) wo = renorm(cosine_dir_framed<T>(nf, front, frame_addr, r.x, r.y));
	else wo = renorm(cosine_dir<T>(nf, r.x, r.y));
	return true;
}

// Material response, general path (any texture kind). Returns alive.
template <typename T>
__device__ bool scatter(const DevScene &sc, int mat, int tex, V3<T> wi, V3<T> Ng, V3<T> P, T u, T v, Rnd4<T> r,
	V3<T> &wo, V3<T> &att, V3<T> &emit, const T *turb = nullptr) {
	const MaterialRec &m = sc.mats[mat];
	emit = mk<T>(T(0), T(0), T(0));
	att = emit;
	wo = mk<T>(nan_t<T>(), nan_t<T>(), nan_t<T>());
	const int kind = m.kind;
	if (kind == MK_LIGHT) {
		emit = mparam<T>(m, 1) * tex_eval<T>(sc, mat_texture(m, tex), u, v, P);
		return false;
	}
	if (kind == MK_DIELECTRIC) {
		att = mk<T>(T(1), T(1), T(1));
		return scatter_dir<T>(MK_DIELECTRIC, mparam<T>(m, 0), wi, Ng, r, wo);
	}
	if (kind == MK_REFLECTIVE && r.z < mparam<T>(m, 0)) {
		att = mk<T>(mparam<T>(m, 1), mparam<T>(m, 2), mparam<T>(m, 3));
		return scatter_dir<T>(MK_REFLECTIVE, T(0), wi, Ng, r, wo);
	}
	att = tex_eval<T>(sc, mat_texture(m, tex), u, v, P, turb);
	return scatter_dir<T>(kind == MK_METAL ? MK_METAL : MK_LAMBERTIAN, kind == MK_METAL ? mparam<T>(m, 0) : T(0), wi, Ng, r, wo);
}

// ---- surface resolution -------------------------------------------------------------------------------
// Geometric unit normal (not flipped) and (u,v) of device primitive `dp` at point P, fp32 render data.
// (a,b) = planar coordinates of P in the primitive's own (u,v) frame (ignored for spheres).
// need_uv = false: the caller's texture is position-only (solid, 3-D checker, noise) — a sphere's (u, v) costs an acosf and
// an atan2f, ~100 instructions that such surfaces never look at.
__device__ __forceinline__ void surface_at(const DevScene &sc, int dp, V3<float> P, float a, float b, V3<float> &N, float &u, float &v, bool need_uv = true) {
	const HotPrim &h = sc.prim_plane[dp];
	if (dp >= sc.n_tri + sc.n_quad) {
		float inv_r = 1.0f / h.r0.w;
		N = inv_r * (P - mk<float>(h.r0.x, h.r0.y, h.r0.z));
		u = 0.0f; v = 0.0f;
		if (!need_uv) return;
		float theta = acosf(fmaxf(-1.0f, fminf(1.0f, -N.y))), phi = atan2f(-N.z, N.x) + Pi<float>::value;
		u = phi * (0.5f / Pi<float>::value);
		v = theta * (1.0f / Pi<float>::value);
	} else {
		N = mk<float>(h.r0.x, h.r0.y, h.r0.z);
		if (dp < sc.n_tri) {
			const float *t = sc.tri_uv + 6 * dp;
			float b0 = 1.0f - a - b;
			u = fmaf(t[4], b, fmaf(t[2], a, t[0] * b0));
			v = fmaf(t[5], b, fmaf(t[3], a, t[1] * b0));
		} else {
			u = a;
			v = b;
		}
	}
}

// ---- camera -------------------------------------------------------------------------------------------
template <typename T>
struct CamT {
	V3<T> pos, fwd, right, up;
	T sx, sy, lens_r, focus;
	int jitter;
};
template <typename T>
__host__ __device__ __forceinline__ CamT<T> cam_from_basis(const CamBasis &b) {
	CamT<T> c;
	c.pos = ld3<T>(b.pos); c.fwd = ld3<T>(b.fwd); c.right = ld3<T>(b.right); c.up = ld3<T>(b.up);
	c.sx = T(b.sx); c.sy = T(b.sy); c.lens_r = T(b.lens_r); c.focus = T(b.focus);
	c.jitter = b.jitter;
	return c;
}
__device__ __forceinline__ CamT<float> cam_from_f32(const CamF &b) {
	CamT<float> c;
	c.pos = mk<float>(b.pos[0], b.pos[1], b.pos[2]); c.fwd = mk<float>(b.fwd[0], b.fwd[1], b.fwd[2]);
	c.right = mk<float>(b.right[0], b.right[1], b.right[2]); c.up = mk<float>(b.up[0], b.up[1], b.up[2]);
	c.sx = b.sx; c.sy = b.sy; c.lens_r = b.lens_r; c.focus = b.focus;
	c.jitter = b.jitter;
	return c;
}
// r = (sx, sy, lens r0, lens r1); the caller substitutes sx = sy = 0.5 when jitter is off
template <typename T>
__device__ __forceinline__ void cam_ray(const CamT<T> &c, T inv_w, T inv_h, int x, int y, Rnd4<T> r, V3<T> &o, V3<T> &d) {
	T fx = (T(2) * (T(x) + r.x) * inv_w - T(1)) * c.sx, fy = (T(1) - T(2) * (T(y) + r.y) * inv_h) * c.sy;
	V3<T> dir = c.fwd + mad2(fx, c.right, fy, c.up);
	if (c.lens_r > T(0)) {
		T rr = c.lens_r * sqrt_t(r.z), sn, cs;
		sincos2pi_t(r.w, &sn, &cs);
		V3<T> off = mad2(rr * cs, c.right, rr * sn, c.up);
		o = c.pos + off;
		d = nrm(c.focus * dir - off);
	} else {
		o = c.pos;
		d = nrm(dir);
	}
}
// fp32 (the render kernels, the wavefront's generate stage and the precision=32 camera harness): the same ray with the
// affine maps folded — fx = (x + sx) kx - c.sx with kx = 2 c.sx / W (warp-uniform, hoisted out of the loop), and the
// direction as two chained multiply-adds on fwd — 10 instructions fewer per camera ray, which the render loop executes
// on every trip for the ~5 lanes that start a new path.
template <>
__device__ __forceinline__ void cam_ray<float>(const CamT<float> &c, float inv_w, float inv_h, int x, int y, Rnd4<float> r, V3<float> &o, V3<float> &d) {
	const float kx = 2.0f * inv_w * c.sx, ky = -2.0f * inv_h * c.sy;
	const float fx = fmaf((float)x + r.x, kx, -c.sx), fy = fmaf((float)y + r.y, ky, c.sy);
	const V3<float> dir = mad(fx, c.right, mad(fy, c.up, c.fwd));
	if (c.lens_r > 0.0f) {
		float rr = c.lens_r * sqrt_t(r.z), sn, cs;
		sincos2pi_t(r.w, &sn, &cs);
		const V3<float> off = mad2(rr * cs, c.right, rr * sn, c.up);
		o = c.pos + off;
		d = nrm(c.focus * dir - off);
	} else {
		o = c.pos;
		d = nrm(dir);
	}
}

}  // namespace areb
