// rtao.cu — the reference's own shading loop (experiments/rt.cpp, BASELINE config 0) on the device.
//
// This kernel does NOT use the plane-form primitives of the path tracer: it evaluates rt.cpp's arithmetic itself,
// in fp32, operation for operation —
//     Tri::intersect      Möller–Trumbore, EPS = 1e-5, strict [0,1] barycentrics      rt.cpp:123-138
//     intersect           closest hit by linear scan, first-wins on ties               rt.cpp:209-218
//     computeAO           N uniform-sphere rays flipped to +z, Rodrigues z -> n         rt.cpp:221-248
//     trace               p from barycentrics, bary2uv, mirror bounce at depth 0,       rt.cpp:251-334
//                         cosine gather with Russian roulette for non-mirror surfaces
//     render              pinhole pixel-centre rays                                      rt.cpp:339-343,364-366
// and the file is compiled with -fmad=false and IEEE division / square root, so every +,-,*,/ and sqrt rounds as in
// the reference's x86-64 build.  That matters: rt.cpp's mirror ray starts EPS above the surface and points INTO it
// (rt.cpp:268-270 reflects the view vector, not the incoming direction), so it re-hits its own triangle at
// t ~ EPS/cos — a decision that sits on the t > EPS threshold and flips with the last bit.  Only sinf/cosf/acosf
// (AO and gather directions) may differ from glibc by an ulp; the random source is Philox instead of
// mt19937(random_device); stream layout: counter = (pixel, sample, slot, stream) with stream 1 = AO
// ray `slot`, stream 2 = gather sample/bounce `slot`.
#include "dev_types.h"
#include "kernels.h"
#include "philox.cuh"

namespace areb {

namespace {

struct F3 {
	float x, y, z;
};
__device__ __forceinline__ F3 F(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 add(F3 a, F3 b) { return F(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 sub(F3 a, F3 b) { return F(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 sc(F3 a, float s) { return F(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 mul3(F3 a, F3 b) { return F(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return F(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float len(F3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ F3 norm(F3 a) { return sc(a, 1.0f / len(a)); }  // rt.cpp:46-48
__device__ __forceinline__ F3 clamp01(F3 c) { return F(fmaxf(0.f, fminf(1.f, c.x)), fmaxf(0.f, fminf(1.f, c.y)), fmaxf(0.f, fminf(1.f, c.z))); }

#define RT_EPS 1e-5f  // rt.cpp:15
#define RT_PI 3.14159265358979323846  // M_PI (double) as rt.cpp uses it

struct RtTri {
	F3 v0, e1, e2, n;
};
__device__ __forceinline__ RtTri load_tri(const float4 *s, int i) {
	float4 a = s[3 * i], b = s[3 * i + 1], c = s[3 * i + 2];
	RtTri t;
	t.v0 = F(a.x, a.y, a.z); t.e1 = F(a.w, b.x, b.y); t.e2 = F(b.z, b.w, c.x); t.n = F(c.y, c.z, c.w);
	return t;
}
struct RtHit {
	int tri;
	float t, u, v;
};
// rt.cpp:209-218 over rt.cpp:123-138
__device__ __forceinline__ RtHit rt_intersect(const float4 *s_tris, int ntri, F3 o, F3 d, unsigned int &rays) {
	RtHit r;
	r.tri = -1; r.t = 1e30f; r.u = 0.f; r.v = 0.f;
	for (int i = 0; i < ntri; ++i) {
		const RtTri T = load_tri(s_tris, i);
		F3 h = cross(d, T.e2), s = sub(o, T.v0);
		float a = dot(T.e1, h);
		if (fabsf(a) < RT_EPS) continue;
		float f = 1.f / a;
		float u = f * dot(s, h);
		if (u < 0 || u > 1) continue;
		F3 q = cross(s, T.e1);
		float v = f * dot(d, q);
		if (v < 0 || u + v > 1) continue;
		float t = f * dot(T.e2, q);
		if (t > RT_EPS && t < r.t) { r.tri = i; r.t = t; r.u = u; r.v = v; }
	}
	++rays;
	return r;
}

struct RtSurf {
	F3 p, n, albedo, tint;
	float refl;
	bool mirror;
};
__device__ __forceinline__ RtSurf rt_surface(const DevScene &S, const float4 *s_tris, const RtHit &h) {
	const RtTri T = load_tri(s_tris, h.tri);
	RtSurf sf;
	sf.n = T.n;
	F3 v1 = add(T.v0, T.e1), v2 = add(T.v0, T.e2);
	float b0 = 1 - h.u - h.v;
	sf.p = add(add(sc(T.v0, b0), sc(v1, h.u)), sc(v2, h.v));  // rt.cpp:258
	const float *t6 = S.tri_uv + 6 * h.tri;                   // rt.cpp:139-142
	float uu = t6[0] * b0 + t6[2] * h.u + t6[4] * h.v;
	float vv = t6[1] * b0 + t6[3] * h.u + t6[5] * h.v;
	const PrimInfo pi = S.info[h.tri];
	const TextureRec &tx = S.texs[pi.tex];
	if (tx.kind == TK_CHECKER_UV) {                           // rt.cpp:98-101
		int xx = (int)floorf(uu * tx.pf[0]), yy = (int)floorf(vv * tx.pf[0]);
		sf.albedo = ((xx + yy) % 2 == 0) ? F(tx.pf[1], tx.pf[2], tx.pf[3]) : F(tx.pf[4], tx.pf[5], tx.pf[6]);
	} else sf.albedo = F(tx.pf[0], tx.pf[1], tx.pf[2]);
	const MaterialRec &m = S.mats[pi.mat];
	sf.mirror = m.kind == MK_REFLECTIVE;
	sf.refl = m.pf[0];
	sf.tint = F(m.pf[1], m.pf[2], m.pf[3]);
	return sf;
}
// rt.cpp:221-248
__device__ float rt_ao(const RenderArgs &A, const float4 *s_tris, int ntri, F3 p, F3 n, uint32_t pixel, uint32_t sample, uint32_t slot_base, unsigned int &rays) {
	const int N = A.ao_samples;
	int unocc = 0;
	for (int i = 0; i < N; ++i) {
		Rnd4<float> r = rnd4<float>(A.key, pixel, sample, slot_base + (uint32_t)i, 1u);
		float theta = (float)(2 * RT_PI * (double)r.x);  // rt.cpp:227: double product, narrowed
		float phi = acosf(1 - 2 * r.y);
		float x = sinf(phi) * cosf(theta), y = sinf(phi) * sinf(theta), z = cosf(phi);
		if (z < 0) z = -z;
		F3 hemi = F(x, y, z), axis = cross(F(0, 0, 1), n);
		float sa = len(axis), ca = dot(F(0, 0, 1), n);
		F3 d = hemi;
		if (sa > RT_EPS) {
			axis = norm(axis);
			float ang = acosf(ca);
			d = add(add(sc(d, cosf(ang)), sc(cross(axis, d), sinf(ang))), sc(axis, dot(axis, d) * (1 - cosf(ang))));
		}
		d = norm(d);
		if (rt_intersect(s_tris, ntri, add(p, sc(n, RT_EPS)), d, rays).tri == -1) unocc++;
	}
	return 0.25f + 0.75f * (unocc / (float)N);
}
// rt.cpp:50-55
__device__ __forceinline__ F3 rt_rotate(F3 normal, float u, float v) {
	F3 up = fabsf(normal.z) < 0.999f ? F(0, 0, 1) : F(1, 0, 0);
	F3 tangent = norm(cross(normal, up));
	F3 bitangent = cross(normal, tangent);
	return add(add(sc(tangent, u), sc(bitangent, v)), sc(normal, sqrtf(fmaxf(0.f, 1 - u * u - v * v))));
}
// rt.cpp:278-329
__device__ F3 rt_gather(const RenderArgs &A, const float4 *s_tris, int ntri, const RtSurf &sf, uint32_t pixel, uint32_t sample, int depth, unsigned int &rays) {
	const uint32_t N = (uint32_t)A.ao_samples;
	F3 accum = F(0, 0, 0);
	for (uint32_t k = 0; k < N; ++k) {
		const uint32_t slot = ((uint32_t)depth * N + k) * 4u;
		Rnd4<float> r = rnd4<float>(A.key, pixel, sample, slot, 2u);
		float phi = (float)(2 * RT_PI * (double)r.x), r2s = sqrtf(r.y);
		F3 dir = rt_rotate(sf.n, r2s * cosf(phi), r2s * sinf(phi));
		F3 org = add(sf.p, sc(sf.n, RT_EPS)), thr = sf.albedo;
		int b = 0;
		while (b < 3) {
			RtHit bh = rt_intersect(s_tris, ntri, org, dir, rays);
			if (bh.tri == -1) break;
			RtSurf bs = rt_surface(A.sc, s_tris, bh);
			thr = mul3(thr, bs.albedo);
			r = rnd4<float>(A.key, pixel, sample, slot + 1u + (uint32_t)b, 2u);
			float pr = fmaxf(thr.x, fmaxf(thr.y, thr.z));
			if (r.x > pr) break;
			thr = sc(thr, 1 / pr);
			F3 nd;
			if (bs.mirror) {
				F3 view = norm(sub(F(0, 0, 0), dir));
				nd = sub(view, sc(bs.n, 2 * dot(view, bs.n)));
			} else {
				float nphi = (float)(2 * RT_PI * (double)r.y), nr2s = sqrtf(r.z);
				nd = rt_rotate(bs.n, nr2s * cosf(nphi), nr2s * sinf(nphi));
			}
			org = add(bs.p, sc(bs.n, RT_EPS));
			dir = nd;
			b++;
		}
		accum = add(accum, thr);
	}
	return sc(accum, 1.0f / (float)N);
}

}  // namespace

#define RTAO_THREADS 128

__global__ void __launch_bounds__(RTAO_THREADS) k_render_rtao(const __grid_constant__ RenderArgs A) {
	extern __shared__ float4 s_tris[];
	const int ntri = A.sc.n_tri;
	{
		const float4 *src = reinterpret_cast<const float4 *>(A.sc.rt_tris);
		for (int i = threadIdx.x; i < ntri * 3; i += RTAO_THREADS) s_tris[i] = __ldg(src + i);
		__syncthreads();
	}
	// a warp covers an 8x4 pixel tile, a block a 16x8 tile
	const int tiles_x = (A.W + 15) >> 4;
	const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int x = bx * 16 + (warp & 1) * 8 + (lane & 7), y = by * 8 + (warp >> 1) * 4 + (lane >> 3);
	const bool inside = x < A.W && y < A.H;
	const uint32_t pixel = (uint32_t)(y * A.W + x);
	const RtCam &c = A.rtcam;
	const F3 pos = F(c.pos[0], c.pos[1], c.pos[2]), fwd = F(c.fwd[0], c.fwd[1], c.fwd[2]), right = F(c.right[0], c.right[1], c.right[2]),
			 up = F(c.up[0], c.up[1], c.up[2]);
	const F3 bg = F(0.06f, 0.09f, 0.14f);  // rt.cpp:255
	F3 sum = F(0, 0, 0);
	unsigned int rays = 0;
	const int s_count = inside ? A.s_count : 0;
	for (int s = 0; s < s_count; ++s) {
		const uint32_t sample = (uint32_t)(A.s_begin + s);
		float fx = (2 * (x + 0.5f) / A.W - 1) * c.aspect * c.scale;  // rt.cpp:364-366
		float fy = (1 - 2 * (y + 0.5f) / A.H) * c.scale;
		F3 dir = norm(add(add(fwd, sc(right, fx)), sc(up, fy)));
		F3 col = bg;
		RtHit h0 = rt_intersect(s_tris, ntri, pos, dir, rays);
		if (h0.tri != -1) {
			const RtSurf s0 = rt_surface(A.sc, s_tris, h0);
			const float ao0 = rt_ao(A, s_tris, ntri, s0.p, s0.n, pixel, sample, 0u, rays);
			if (s0.mirror) {  // rt.cpp:267-275
				F3 view = norm(sub(pos, s0.p));
				F3 refl = sub(view, sc(s0.n, 2 * dot(view, s0.n)));
				F3 reflected = bg;
				RtHit h1 = rt_intersect(s_tris, ntri, add(s0.p, sc(s0.n, RT_EPS)), refl, rays);
				if (h1.tri != -1) {
					const RtSurf s1 = rt_surface(A.sc, s_tris, h1);
					const float ao1 = rt_ao(A, s_tris, ntri, s1.p, s1.n, pixel, sample, (uint32_t)A.ao_samples, rays);
					F3 a1 = s1.albedo;
					if (!s1.mirror) a1 = add(a1, rt_gather(A, s_tris, ntri, s1, pixel, sample, 1, rays));
					reflected = clamp01(sc(a1, ao1));
				}
				F3 ret = add(sc(s0.albedo, 1 - s0.refl), mul3(sc(reflected, s0.refl), s0.tint));
				col = clamp01(sc(ret, ao0));
			} else {
				F3 a0 = add(s0.albedo, rt_gather(A, s_tris, ntri, s0, pixel, sample, 0, rays));
				col = clamp01(sc(a0, ao0));
			}
		}
		col = clamp01(col);  // rt.cpp:368
		sum = add(sum, col);
	}
	if (inside) {
		float *acc = A.accum + (size_t)pixel * 3;
		acc[0] += sum.x; acc[1] += sum.y; acc[2] += sum.z;
	}
	unsigned long long r64 = rays;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) r64 += __shfl_down_sync(0xffffffffu, r64, off);
	if (lane == 0 && r64) atomicAdd(A.counters + CNT_RAYS, r64);
}

int launch_render_rtao(const RenderArgs &a, cudaStream_t s) {
	const int tiles = ((a.W + 15) / 16) * ((a.H + 7) / 8);
	const size_t smem = (size_t)a.sc.n_tri * 48;
	if (tiles <= 0 || !a.sc.rt_tris || a.sc.n_tri <= 0 || smem > 48 * 1024) return -1;
	k_render_rtao<<<tiles, RTAO_THREADS, smem, s>>>(a);
	return 1;
}

}  // namespace areb
