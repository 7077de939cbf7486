// render.cu — the sm_100a render kernels.
//
//   k_render_path   megakernel: camera ray generation + pixel x sample x bounce loop with in-place path
//                   regeneration (a lane whose path ends starts its pixel's next sample in the same loop trip, so
//                   warps stay full until a pixel runs out of samples), closest hit (brute force over a
//                   shared-memory copy of the hot primitives, or BVH2 from L2), material scatter, texture
//                   evaluation, accumulation.  Replaces render()+trace() of the reference's only sampling
//                   renderer (/root/reference/experiments/rt.cpp:251-374) as a general path tracer.
//   k_render_rtao   the reference's own shading (rt.cpp:221-334: primary + AO rays + one mirror bounce + cosine
//                   gather with Russian roulette), same arithmetic, Philox instead of mt19937.
//   k_hit32 / k_scatter32 / k_texture32 / k_camera32   per-ray harness over the SAME device routines.
//   k_tonemap       reference encoders (rt.cpp:383-386 gamma 2.2 truncate; src/texture.cpp:384-386 linear truncate).
//   k_fp32_peak     FFMA issue-rate micro-kernel (roofline denominator).
#include <stdio.h>

#include "dev_types.h"
#include "intersect.cuh"
#include "kernels.h"
#include "philox.cuh"
#include "shade.cuh"
#include "vec.cuh"

namespace areb {

typedef V3<float> F3;

#define RENDER_THREADS 128
#define BRUTE_MAX_PRIMS 1024  // 48 KB of shared memory

size_t brute_smem_limit_prims() { return BRUTE_MAX_PRIMS; }

__device__ __forceinline__ F3 background(const RenderArgs &A, F3 d) {
	float a = 0.5f * (d.y + 1.0f);
	return mk<float>((1.0f - a) * A.bg_bottom[0] + a * A.bg_top[0], (1.0f - a) * A.bg_bottom[1] + a * A.bg_top[1],
		(1.0f - a) * A.bg_bottom[2] + a * A.bg_top[2]);
}
__device__ __forceinline__ bool finite3(F3 v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }

// pixel owned by this thread: a warp covers an 8x4 tile, a block a 16x8 tile (coherent primary rays and texture reads)
__device__ __forceinline__ void thread_pixel(int W, int &x, int &y) {
	const int tiles_x = (W + 15) >> 4;
	const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	x = bx * 16 + (warp & 1) * 8 + (lane & 7);
	y = by * 8 + (warp >> 1) * 4 + (lane >> 3);
}

template <bool BVH, bool COUNT>
__global__ void __launch_bounds__(RENDER_THREADS) k_render_path(const __grid_constant__ RenderArgs A) {
	extern __shared__ float4 s_raw[];
	const HotPrim *s_prims = reinterpret_cast<const HotPrim *>(s_raw);
	if (!BVH) {
		const float4 *src = reinterpret_cast<const float4 *>(A.sc.brute);
		const int n4 = A.sc.n_hot * 3;
		for (int i = threadIdx.x; i < n4; i += RENDER_THREADS) s_raw[i] = __ldg(src + i);
		__syncthreads();
	}
	int x, y;
	thread_pixel(A.W, x, y);
	const bool inside = x < A.W && y < A.H;
	const uint32_t pixel = (uint32_t)(y * A.W + x);
	const CamT<float> cam = cam_from_basis<float>(A.cam);
	const float inv_w = 1.0f / (float)A.W, inv_h = 1.0f / (float)A.H;
	const HotRange br = A.sc.brute_range;
	const HotIds *ids = BVH ? A.sc.bvh_ids : A.sc.brute_ids;

	F3 sum = mk<float>(0.f, 0.f, 0.f), o = sum, d = sum, thr = sum;
	int s = 0, bounce = 0;
	unsigned int rays = 0;
	TravCounters tc = { 0, 0, 0, 0 };
	const int s_count = inside ? A.s_count : 0;

	while (true) {
		if (bounce == 0) {
			if (s >= s_count) break;
			Rnd4<float> r = rnd4<float>(A.key, pixel, (uint32_t)(A.s_begin + s), 0u, 0u);
			if (!cam.jitter) { r.x = 0.5f; r.y = 0.5f; }
			cam_ray<float>(cam, inv_w, inv_h, x, y, r, o, d);
			thr = mk<float>(1.f, 1.f, 1.f);
			bounce = 1;
		}
		Hit h;
		h.t = INFINITY; h.idx = -1; h.a = 0.f; h.b = 0.f;
		if (BVH) intersect_bvh<COUNT>(A.sc, o, d, A.tmin, h, &tc);
		else intersect_range<lds4>(s_prims, br.first, br.nq, br.nt, br.ns, o, d, A.tmin, h);
		++rays;
		if (h.idx < 0) {
			F3 c = thr * background(A, d);
			if (finite3(c)) sum = sum + c;
			++s; bounce = 0;
			continue;
		}
		const F3 P = o + h.t * d;
		const Resolved rs = resolve_hit(A.sc, ids, h, P);
		const PrimInfo pi = A.sc.info[rs.dev_prim];
		F3 N, wo, att, emit;
		float u, v;
		surface_at(A.sc, rs.dev_prim, P, rs.a, rs.b, N, u, v);
		const Rnd4<float> r = rnd4<float>(A.key, pixel, (uint32_t)(A.s_begin + s), (uint32_t)bounce, 0u);
		const bool alive = scatter<float>(A.sc, pi.mat, pi.tex, d, N, P, u, v, r, wo, att, emit);
		if (!alive) {
			F3 c = thr * emit;
			if (finite3(c)) sum = sum + c;
			++s; bounce = 0;
			continue;
		}
		thr = thr * att;
		o = P;
		d = wo;
		if (++bounce > A.max_depth) { ++s; bounce = 0; }  // truncated path contributes nothing (RTIOW depth cut-off)
	}
	if (inside) {
		float *acc = A.accum + (size_t)pixel * 3;
		acc[0] += sum.x; acc[1] += sum.y; acc[2] += sum.z;
	}
	// counters: one atomic per warp
	unsigned long long r64 = rays;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) r64 += __shfl_down_sync(0xffffffffu, r64, off);
	if ((threadIdx.x & 31) == 0 && r64) atomicAdd(A.counters + CNT_RAYS, r64);
	if (BVH && COUNT) {
		unsigned long long c[4] = { tc.nodes, tc.quads, tc.tris, tc.spheres };
#pragma unroll
		for (int k = 0; k < 4; ++k) {
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) c[k] += __shfl_down_sync(0xffffffffu, c[k], off);
			if ((threadIdx.x & 31) == 0 && c[k]) atomicAdd(A.counters + CNT_NODES + k, c[k]);
		}
	}
}

int launch_render_path(const RenderArgs &a, bool use_bvh, bool count_tests, cudaStream_t s) {
	const int tiles = ((a.W + 15) / 16) * ((a.H + 7) / 8);
	if (tiles <= 0) return -1;
	if (use_bvh) {
		if (count_tests) k_render_path<true, true><<<tiles, RENDER_THREADS, 0, s>>>(a);
		else k_render_path<true, false><<<tiles, RENDER_THREADS, 0, s>>>(a);
	} else {
		if (!a.sc.brute || a.sc.n_hot > BRUTE_MAX_PRIMS) return -1;
		size_t smem = (size_t)a.sc.n_hot * sizeof(HotPrim);
		k_render_path<false, false><<<tiles, RENDER_THREADS, smem, s>>>(a);
	}
	return 1;
}

// =========================================================================================================
// rt.cpp shading (config 0)
// =========================================================================================================
#define RT_EPS 1e-5f  // rt.cpp:15

struct RtSurf {
	F3 p, n, albedo;
	int mat_kind;
	float refl;
	F3 tint;
};

__device__ __forceinline__ bool rt_closest(const RenderArgs &A, const HotPrim *s_prims, F3 o, F3 d, RtSurf &sf, unsigned int &rays) {
	Hit h;
	h.t = INFINITY; h.idx = -1; h.a = 0.f; h.b = 0.f;
	const HotRange br = A.sc.brute_range;
	intersect_range<lds4>(s_prims, br.first, br.nq, br.nt, br.ns, o, d, RT_EPS, h);
	++rays;
	if (h.idx < 0) return false;
	const F3 P = o + h.t * d;
	const Resolved rs = resolve_hit(A.sc, A.sc.brute_ids, h, P);
	const PrimInfo pi = A.sc.info[rs.dev_prim];
	float u, v;
	surface_at(A.sc, rs.dev_prim, P, rs.a, rs.b, sf.n, u, v);
	sf.p = P;
	sf.albedo = tex_eval<float>(A.sc, pi.tex, u, v, P);
	const MaterialRec &m = A.sc.mats[pi.mat];
	sf.mat_kind = m.kind;
	sf.refl = m.pf[0];
	sf.tint = mk<float>(m.pf[1], m.pf[2], m.pf[3]);
	return true;
}
__device__ __forceinline__ bool rt_occluded(const RenderArgs &A, const HotPrim *s_prims, F3 o, F3 d, unsigned int &rays) {
	Hit h;
	h.t = INFINITY; h.idx = -1; h.a = 0.f; h.b = 0.f;
	const HotRange br = A.sc.brute_range;
	intersect_range<lds4>(s_prims, br.first, br.nq, br.nt, br.ns, o, d, RT_EPS, h);
	++rays;
	return h.idx >= 0;
}
// rt.cpp:221-248
__device__ float rt_ao(const RenderArgs &A, const HotPrim *s_prims, F3 p, F3 n, uint32_t pixel, uint32_t sample, uint32_t slot_base, unsigned int &rays) {
	const int N = A.ao_samples;
	int unocc = 0;
	F3 axis = cross(mk<float>(0.f, 0.f, 1.f), n);
	const float sa = length(axis), ca = n.z;
	const bool rot = sa > RT_EPS;
	float sang = 0.f, cang = 1.f;
	if (rot) {
		axis = (1.0f / sa) * axis;
		float ang = acosf(ca);
		sincosf(ang, &sang, &cang);
	}
	const F3 org = p + RT_EPS * n;
	for (int i = 0; i < N; ++i) {
		Rnd4<float> r = rnd4<float>(A.key, pixel, sample, slot_base + (uint32_t)i, 1u);
		F3 hd = sphere_dir<float>(r.y, r.x);  // theta = 2*pi*u, z = cos(acos(1-2v)) = 1-2v
		if (hd.z < 0.f) hd.z = -hd.z;
		F3 dd = hd;
		if (rot) dd = cang * hd + sang * cross(axis, hd) + (dot(axis, hd) * (1.0f - cang)) * axis;
		dd = nrm(dd);
		if (!rt_occluded(A, s_prims, org, dd, rays)) ++unocc;
	}
	return 0.25f + 0.75f * ((float)unocc / (float)N);
}
__device__ __forceinline__ F3 clamp01(F3 c) {
	return mk<float>(fmaxf(0.f, fminf(1.f, c.x)), fmaxf(0.f, fminf(1.f, c.y)), fmaxf(0.f, fminf(1.f, c.z)));
}
// rt.cpp:50-55
__device__ __forceinline__ F3 rt_rotate(F3 n, float u, float v) {
	F3 up = fabsf(n.z) < 0.999f ? mk<float>(0.f, 0.f, 1.f) : mk<float>(1.f, 0.f, 0.f);
	F3 tangent = nrm(cross(n, up));
	F3 bitangent = cross(n, tangent);
	return u * tangent + v * bitangent + sqrtf(fmaxf(0.f, 1.f - u * u - v * v)) * n;
}
// rt.cpp:278-329 — cosine gather with Russian roulette, only for non-mirror surfaces
__device__ F3 rt_gather(const RenderArgs &A, const HotPrim *s_prims, const RtSurf &sf, uint32_t pixel, uint32_t sample, int depth, unsigned int &rays) {
	const uint32_t N = (uint32_t)A.ao_samples;
	F3 accum = mk<float>(0.f, 0.f, 0.f);
	for (uint32_t k = 0; k < N; ++k) {
		const uint32_t slot = ((uint32_t)depth * N + k) * 4u;
		Rnd4<float> r = rnd4<float>(A.key, pixel, sample, slot, 2u);
		float sn, cs, r2s = sqrtf(r.y);
		sincospif(2.0f * r.x, &sn, &cs);
		F3 dir = rt_rotate(sf.n, r2s * cs, r2s * sn);
		F3 org = sf.p + RT_EPS * sf.n, thr = sf.albedo;
		for (int b = 0; b < 3; ++b) {
			RtSurf bs;
			if (!rt_closest(A, s_prims, org, dir, bs, rays)) break;
			thr = thr * bs.albedo;
			r = rnd4<float>(A.key, pixel, sample, slot + 1u + (uint32_t)b, 2u);
			float pr = fmaxf(thr.x, fmaxf(thr.y, thr.z));
			if (r.x > pr) break;
			thr = (1.0f / pr) * thr;
			F3 nd;
			if (bs.mat_kind == MK_REFLECTIVE) {
				F3 view = nrm(-dir);
				nd = view - (2.0f * dot(view, bs.n)) * bs.n;
			} else {
				float nsn, ncs, nr2s = sqrtf(r.z);
				sincospif(2.0f * r.y, &nsn, &ncs);
				nd = rt_rotate(bs.n, nr2s * ncs, nr2s * nsn);
			}
			org = bs.p + RT_EPS * bs.n;
			dir = nd;
		}
		accum = accum + thr;
	}
	return (1.0f / (float)N) * accum;
}

__global__ void __launch_bounds__(RENDER_THREADS) k_render_rtao(const __grid_constant__ RenderArgs A) {
	extern __shared__ float4 s_raw[];
	const HotPrim *s_prims = reinterpret_cast<const HotPrim *>(s_raw);
	{
		const float4 *src = reinterpret_cast<const float4 *>(A.sc.brute);
		const int n4 = A.sc.n_hot * 3;
		for (int i = threadIdx.x; i < n4; i += RENDER_THREADS) s_raw[i] = __ldg(src + i);
		__syncthreads();
	}
	int x, y;
	thread_pixel(A.W, x, y);
	const bool inside = x < A.W && y < A.H;
	const uint32_t pixel = (uint32_t)(y * A.W + x);
	const CamT<float> cam = cam_from_basis<float>(A.cam);
	const float inv_w = 1.0f / (float)A.W, inv_h = 1.0f / (float)A.H;
	const F3 bg = mk<float>(A.bg_bottom[0], A.bg_bottom[1], A.bg_bottom[2]);
	F3 sum = mk<float>(0.f, 0.f, 0.f);
	unsigned int rays = 0;
	const int s_count = inside ? A.s_count : 0;
	for (int s = 0; s < s_count; ++s) {
		const uint32_t sample = (uint32_t)(A.s_begin + s);
		Rnd4<float> c; c.x = 0.5f; c.y = 0.5f; c.z = 0.f; c.w = 0.f;  // rt.cpp:364-366 pixel centres
		F3 o, d;
		cam_ray<float>(cam, inv_w, inv_h, x, y, c, o, d);
		F3 col = bg;
		RtSurf s0;
		if (rt_closest(A, s_prims, o, d, s0, rays)) {
			const float ao0 = rt_ao(A, s_prims, s0.p, s0.n, pixel, sample, 0u, rays);
			if (s0.mat_kind == MK_REFLECTIVE) {  // rt.cpp:267-275
				F3 view = nrm(o - s0.p);
				F3 refl = view - (2.0f * dot(view, s0.n)) * s0.n;
				F3 reflected = bg;
				RtSurf s1;
				if (rt_closest(A, s_prims, s0.p + RT_EPS * s0.n, refl, s1, rays)) {
					const float ao1 = rt_ao(A, s_prims, s1.p, s1.n, pixel, sample, (uint32_t)A.ao_samples, rays);
					F3 a1 = s1.albedo;
					if (s1.mat_kind != MK_REFLECTIVE) a1 = a1 + rt_gather(A, s_prims, s1, pixel, sample, 1, rays);
					reflected = clamp01(ao1 * a1);
				}
				col = clamp01(ao0 * ((1.0f - s0.refl) * s0.albedo + (s0.refl * reflected) * s0.tint));
			} else {
				F3 a0 = s0.albedo + rt_gather(A, s_prims, s0, pixel, sample, 0, rays);
				col = clamp01(ao0 * a0);
			}
		}
		sum = sum + col;
	}
	if (inside) {
		float *acc = A.accum + (size_t)pixel * 3;
		acc[0] += sum.x; acc[1] += sum.y; acc[2] += sum.z;
	}
	unsigned long long r64 = rays;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) r64 += __shfl_down_sync(0xffffffffu, r64, off);
	if ((threadIdx.x & 31) == 0 && r64) atomicAdd(A.counters + CNT_RAYS, r64);
}

int launch_render_rtao(const RenderArgs &a, cudaStream_t s) {
	const int tiles = ((a.W + 15) / 16) * ((a.H + 7) / 8);
	if (tiles <= 0 || !a.sc.brute || a.sc.n_hot > BRUTE_MAX_PRIMS) return -1;
	size_t smem = (size_t)a.sc.n_hot * sizeof(HotPrim);
	k_render_rtao<<<tiles, RENDER_THREADS, smem, s>>>(a);
	return 1;
}

// =========================================================================================================
// fp32 per-ray harness
// =========================================================================================================
__global__ void k_hit32(DevScene sc, int n, const double *__restrict__ Q, const double *__restrict__ D, float tmin, int use_bvh,
	int *__restrict__ prim, double *__restrict__ tout, double *__restrict__ P, double *__restrict__ N, double *__restrict__ uv) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	F3 o = ld3<float>(Q + 3 * i), d = nrm(ld3<float>(D + 3 * i));
	Hit h;
	h.t = INFINITY; h.idx = -1; h.a = 0.f; h.b = 0.f;
	TravCounters tc;
	if (use_bvh) intersect_bvh<false>(sc, o, d, tmin, h, &tc);
	else intersect_range<ldg4>(sc.brute, sc.brute_range.first, sc.brute_range.nq, sc.brute_range.nt, sc.brute_range.ns, o, d, tmin, h);
	const float nan = nan_t<float>();
	F3 x = mk<float>(nan, nan, nan), nn = x;
	float cu = nan, cv = nan;
	int uid = -1;
	if (h.idx >= 0) {
		x = o + h.t * d;
		Resolved rs = resolve_hit(sc, use_bvh ? sc.bvh_ids : sc.brute_ids, h, x);
		uid = sc.info[rs.dev_prim].user_id;
		surface_at(sc, rs.dev_prim, x, rs.a, rs.b, nn, cu, cv);
	}
	if (prim) prim[i] = uid;
	if (tout) tout[i] = h.idx >= 0 ? (double)h.t : (double)nan;
	if (P) st3(P + 3 * i, x);
	if (N) st3(N + 3 * i, nn);
	if (uv) { uv[2 * i] = cu; uv[2 * i + 1] = cv; }
}
__global__ void k_scatter32(DevScene sc, int n, const int *mat, const int *tex, const double *wi, const double *N, const double *P,
	const double *uv, const double *rnd, double *wo, double *att, double *emit, int *alive) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Rnd4<float> r;
	r.x = (float)rnd[4 * i]; r.y = (float)rnd[4 * i + 1]; r.z = (float)rnd[4 * i + 2]; r.w = (float)rnd[4 * i + 3];
	F3 o, a, e;
	bool ok = scatter<float>(sc, mat[i], tex[i], ld3<float>(wi + 3 * i), ld3<float>(N + 3 * i), ld3<float>(P + 3 * i), (float)uv[2 * i], (float)uv[2 * i + 1], r, o, a, e);
	st3(wo + 3 * i, o); st3(att + 3 * i, a); st3(emit + 3 * i, e);
	alive[i] = ok ? 1 : 0;
}
__global__ void k_texture32(DevScene sc, int n, const int *tex, const double *uv, const double *P, double *rgb) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	st3(rgb + 3 * i, tex_eval<float>(sc, tex[i], (float)uv[2 * i], (float)uv[2 * i + 1], ld3<float>(P + 3 * i)));
}
__global__ void k_camera32(CamBasis cb, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	CamT<float> c = cam_from_basis<float>(cb);
	Rnd4<float> r;
	r.x = (float)rnd[4 * i]; r.y = (float)rnd[4 * i + 1]; r.z = (float)rnd[4 * i + 2]; r.w = (float)rnd[4 * i + 3];
	F3 o, d;
	cam_ray<float>(c, 1.0f / (float)W, 1.0f / (float)H, px[i], py[i], r, o, d);
	st3(Q + 3 * i, o); st3(D + 3 * i, d);
}

static inline int blocks(int n) { return (n + 127) / 128; }

void launch_hit32(const DevScene &sc, int n, const double *Q, const double *D, double tmin, bool use_bvh, int *prim, double *t, double *P, double *N, double *uv, cudaStream_t s) {
	if (n > 0) k_hit32<<<blocks(n), 128, 0, s>>>(sc, n, Q, D, (float)tmin, use_bvh ? 1 : 0, prim, t, P, N, uv);
}
void launch_scatter32(const DevScene &sc, int n, const int *mat, const int *tex, const double *wi, const double *N, const double *P, const double *uv,
	const double *rnd, double *wo, double *att, double *emit, int *alive, cudaStream_t s) {
	if (n > 0) k_scatter32<<<blocks(n), 128, 0, s>>>(sc, n, mat, tex, wi, N, P, uv, rnd, wo, att, emit, alive);
}
void launch_texture32(const DevScene &sc, int n, const int *tex, const double *uv, const double *P, double *rgb, cudaStream_t s) {
	if (n > 0) k_texture32<<<blocks(n), 128, 0, s>>>(sc, n, tex, uv, P, rgb);
}
void launch_camera32(const CamBasis &cb, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D, cudaStream_t s) {
	if (n > 0) k_camera32<<<blocks(n), 128, 0, s>>>(cb, W, H, n, px, py, rnd, Q, D);
}

// =========================================================================================================
// output encoders
// =========================================================================================================
__global__ void k_tonemap(const float *__restrict__ accum, long n, double inv_spp, int encoder, uint8_t *__restrict__ out) {
	long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (encoder == 1) {  // src/texture.cpp:384-386 in fp64: (unsigned char)clamp(c*255, 0, 255)
		double c = (double)accum[i] * inv_spp;
		double v = c * 255.0;
		v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
		out[i] = (uint8_t)v;
	} else if (encoder == 2) {
		float c = accum[i] * (float)inv_spp;
		float g = c > 0.f ? sqrtf(c) : 0.f;
		g = g < 0.f ? 0.f : (g > 0.999f ? 0.999f : g);
		out[i] = (uint8_t)(256.0f * g);
	} else {  // rt.cpp:72-76,383-386 in fp32
		float c = accum[i] * (float)inv_spp;
		c = fmaxf(0.0f, fminf(1.0f, c));
		out[i] = (uint8_t)(powf(c, 1 / 2.2f) * 255);
	}
}
void launch_tonemap(const float *accum, int W, int H, double inv_spp, int encoder, uint8_t *out, cudaStream_t s) {
	long n = (long)W * H * 3;
	if (n > 0) k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(accum, n, inv_spp, encoder, out);
}

// =========================================================================================================
// FP32 FMA issue peak
// =========================================================================================================
__global__ void __launch_bounds__(256) k_fp32_peak(float *sink, int iters) {
	float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
	const float m = 0.999f, c = 1e-3f;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
			a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
		}
	}
	float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
	if (r == 123.456f) sink[0] = r;
}
double launch_fp32_peak(float *sink, int sm_count, int iters, cudaStream_t s) {
	const int grid = sm_count * 8;
	k_fp32_peak<<<grid, 256, 0, s>>>(sink, iters);
	return (double)grid * 256.0 * (double)iters * 16.0 * 8.0 * 2.0;
}

}  // namespace areb
