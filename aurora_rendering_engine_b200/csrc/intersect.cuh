// intersect.cuh — fp32 ray/primitive tests and scene traversal used by the render kernels (and, so that the
// harness checks the code that actually renders, by the precision=32 per-ray kernels).
//
// Ray directions are unit length (are::Ray's convention, /root/reference/src/basic/ray.cpp:5), so t is the
// geometric distance.  Triangles and parallelograms are held in the 48-byte plane form of dev_types.h: the
// same hit the reference's Möller–Trumbore routine finds (src/object/triangle.cpp:82-121 — two-sided, no
// culling), evaluated as plane distance + two planar coordinates, 16 FP32 instructions + one MUFU.RCP instead
// of ~27 + RCP.  A hot quad may stand for a fused pair of coplanar triangles forming a parallelogram; the
// owning triangle is recovered after the loop from alpha >= beta.
#pragma once
#include "dev_types.h"
#include "vec.cuh"

namespace areb {

struct Hit {
	float t;   // closest distance so far (INFINITY = none)
	int idx;   // index into the hot array traversed (-1 = miss)
	float a, b;  // planar coordinates in the hot primitive's frame
};

__device__ __forceinline__ float4 ldg4(const f4 *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 lds4(const f4 *p) { return *reinterpret_cast<const float4 *>(p); }

// One plane-form test. QUAD: accept alpha,beta in [0,1]; else triangle: alpha,beta >= 0, alpha+beta <= 1.
template <bool QUAD>
__device__ __forceinline__ void test_plane(float4 r0, float4 r1, float4 r2, V3<float> o, V3<float> d, float tmin, int idx, Hit &h) {
	float denom = r0.x * d.x + r0.y * d.y + r0.z * d.z;
	float num = r0.w - (r0.x * o.x + r0.y * o.y + r0.z * o.z);
	float t = __fdividef(num, denom);  // denom == 0 -> inf/NaN, rejected by the window test below
	float px = fmaf(t, d.x, o.x), py = fmaf(t, d.y, o.y), pz = fmaf(t, d.z, o.z);
	float a = r1.x * px + r1.y * py + r1.z * pz - r1.w;
	float b = r2.x * px + r2.y * py + r2.z * pz - r2.w;
	bool ok = (t > tmin) & (t < h.t) & (a >= 0.0f) & (b >= 0.0f);
	if (QUAD) ok = ok & (a <= 1.0f) & (b <= 1.0f);
	else ok = ok & (a + b <= 1.0f);
	if (ok) { h.t = t; h.idx = idx; h.a = a; h.b = b; }
}

// Sphere: r0 = (c, r), r1.x = r*r. oc = c - o; h = d·oc; c = |oc|^2 - r^2; disc = h^2 - c (|d| = 1)
__device__ __forceinline__ void test_sphere(float4 r0, float4 r1, V3<float> o, V3<float> d, float tmin, int idx, Hit &h) {
	float ox = r0.x - o.x, oy = r0.y - o.y, oz = r0.z - o.z;
	float hh = d.x * ox + d.y * oy + d.z * oz;
	float c = ox * ox + oy * oy + oz * oz - r1.x;
	float disc = hh * hh - c;
	float sq = sqrtf(fmaxf(disc, 0.0f));
	float t0 = hh - sq, t1 = hh + sq;
	float t = (t0 > tmin) ? t0 : t1;  // nearest root inside the window (t0 < t1 always)
	bool ok = (disc >= 0.0f) & (t > tmin) & (t < h.t);
	if (ok) { h.t = t; h.idx = idx; h.a = 0.0f; h.b = 0.0f; }
}

// Test a type-sorted run of hot primitives. LD = ldg4 (global / L2) or lds4 (shared-memory copy).
template <float4 (*LD)(const f4 *)>
__device__ __forceinline__ void intersect_range(const HotPrim *prims, int first, int nq, int nt, int ns, V3<float> o, V3<float> d, float tmin, Hit &h) {
	int i = first;
	int end = first + nq;
#pragma unroll 4
	for (; i < end; ++i) test_plane<true>(LD(&prims[i].r0), LD(&prims[i].r1), LD(&prims[i].r2), o, d, tmin, i, h);
	end += nt;
#pragma unroll 4
	for (; i < end; ++i) test_plane<false>(LD(&prims[i].r0), LD(&prims[i].r1), LD(&prims[i].r2), o, d, tmin, i, h);
	end += ns;
#pragma unroll 4
	for (; i < end; ++i) test_sphere(LD(&prims[i].r0), LD(&prims[i].r1), o, d, tmin, i, h);
}

// ---- BVH2 traversal, per-thread stack -----------------------------------------------------------------
#define ARE_BVH_STACK 48

struct TravCounters {
	unsigned long long nodes, quads, tris, spheres;
};

template <bool COUNT>
__device__ __forceinline__ void intersect_bvh(const DevScene &sc, V3<float> o, V3<float> d, float tmin, Hit &h, TravCounters *cnt) {
	if (sc.n_nodes == 0) {
		int m = sc.root_leaf_meta;
		intersect_range<ldg4>(sc.bvh_prims, 0, m & 255, (m >> 8) & 255, (m >> 16) & 255, o, d, tmin, h);
		if (COUNT) { cnt->quads += m & 255; cnt->tris += (m >> 8) & 255; cnt->spheres += (m >> 16) & 255; }
		return;
	}
	// safe reciprocals: a zero component becomes a huge finite slope so the slab maths never sees 0*inf
	const float big = 1e30f;
	float idx = fabsf(d.x) > 1e-30f ? 1.0f / d.x : (d.x < 0 ? -big : big);
	float idy = fabsf(d.y) > 1e-30f ? 1.0f / d.y : (d.y < 0 ? -big : big);
	float idz = fabsf(d.z) > 1e-30f ? 1.0f / d.z : (d.z < 0 ? -big : big);
	float oxi = o.x * idx, oyi = o.y * idy, ozi = o.z * idz;
	int stack[ARE_BVH_STACK];
	int sp = 0;
	int node = 0;
	while (true) {
		const BvhNode *n = sc.nodes + node;
		float4 b0 = ldg4(&n->b0), b1 = ldg4(&n->b1), b2 = ldg4(&n->b2);
		int4 cm = __ldg(reinterpret_cast<const int4 *>(&n->child[0]));
		if (COUNT) cnt->nodes++;
		float c0lox = fmaf(b0.x, idx, -oxi), c0hix = fmaf(b0.y, idx, -oxi), c0loy = fmaf(b0.z, idy, -oyi), c0hiy = fmaf(b0.w, idy, -oyi);
		float c0loz = fmaf(b2.x, idz, -ozi), c0hiz = fmaf(b2.y, idz, -ozi);
		float c1lox = fmaf(b1.x, idx, -oxi), c1hix = fmaf(b1.y, idx, -oxi), c1loy = fmaf(b1.z, idy, -oyi), c1hiy = fmaf(b1.w, idy, -oyi);
		float c1loz = fmaf(b2.z, idz, -ozi), c1hiz = fmaf(b2.w, idz, -ozi);
		float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin));
		float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
		float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin));
		float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
		bool hit0 = t0n <= t0f, hit1 = t1n <= t1f;
		int next = -1;  // inner node to descend into
		// leaves are intersected immediately; inner children are ordered near-first
		if (hit0 && cm.x < 0) {
			int m = cm.z;
			intersect_range<ldg4>(sc.bvh_prims, ~cm.x, m & 255, (m >> 8) & 255, (m >> 16) & 255, o, d, tmin, h);
			if (COUNT) { cnt->quads += m & 255; cnt->tris += (m >> 8) & 255; cnt->spheres += (m >> 16) & 255; }
			hit0 = false;
		}
		if (hit1 && cm.y < 0) {
			if (t1n <= h.t) {
				int m = cm.w;
				intersect_range<ldg4>(sc.bvh_prims, ~cm.y, m & 255, (m >> 8) & 255, (m >> 16) & 255, o, d, tmin, h);
				if (COUNT) { cnt->quads += m & 255; cnt->tris += (m >> 8) & 255; cnt->spheres += (m >> 16) & 255; }
			}
			hit1 = false;
		}
		if (hit0 && hit1) {
			bool near0 = t0n <= t1n;
			next = near0 ? cm.x : cm.y;
			if (sp < ARE_BVH_STACK) stack[sp++] = near0 ? cm.y : cm.x;
		} else if (hit0) next = cm.x;
		else if (hit1) next = cm.y;
		if (next < 0) {
			if (sp == 0) break;
			next = stack[--sp];
		}
		node = next;
	}
}

// ---- map a hot hit back to the user primitive ---------------------------------------------------------
struct Resolved {
	int dev_prim;   // device primitive index (device order: triangles, quads, spheres)
	float a, b;     // planar coordinates in that primitive's own frame
};
__device__ __forceinline__ Resolved resolve_hit(const DevScene &sc, const HotIds *ids, const Hit &h, V3<float> P) {
	Resolved r;
	HotIds id = ids[h.idx];
	r.a = h.a;
	r.b = h.b;
	r.dev_prim = id.a;
	if (id.b >= 0) {  // fused pair: pick the triangle, recompute its own barycentrics from its plane form
		r.dev_prim = (h.a >= h.b) ? id.a : id.b;
		const HotPrim &tp = sc.prim_plane[r.dev_prim];
		r.a = tp.r1.x * P.x + tp.r1.y * P.y + tp.r1.z * P.z - tp.r1.w;
		r.b = tp.r2.x * P.x + tp.r2.y * P.y + tp.r2.z * P.z - tp.r2.w;
	}
	return r;
}

}  // namespace areb
