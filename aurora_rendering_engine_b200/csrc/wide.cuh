// wide.cuh — traversal of the compressed 8-wide BVH (WideNode, dev_types.h).
//
// Cursor (kept by the caller so a traversal can be suspended and resumed, like the BVH2 one):
//   ng = (first child node of the current group, hit mask << 24 | imask)   inner children still to visit
//   tg = (first primitive slot of the current node, 24-bit mask)            leaf primitives still to test
//   stack of node groups with unvisited inner children
// A node visit decodes eight quantised child boxes (three FMAs per bound on a per-ray scaled grid) and produces
// both masks at once; inner children are visited highest bit first, and the builder's octant slot order makes
// bit = 24 + (slot XOR ray octant) a front-to-back order without storing any distance.
#pragma once
#include "dev_types.h"
#include "intersect.cuh"
#include "vec.cuh"

namespace areb {

#define ARE_WIDE_STACK 32

struct WideRay {
	float idx, idy, idz;   // reciprocal direction (huge finite slope for a zero component)
	float ox, oy, oz;      // origin
	unsigned octinv4;      // (7 - octant) replicated in four bytes; octant bit k = direction component k negative
	bool nx, ny, nz;
};
__device__ __forceinline__ WideRay wide_ray(V3<float> o, V3<float> d) {
	const RaySlopes rs = ray_slopes(o, d);
	WideRay r;
	r.idx = rs.idx; r.idy = rs.idy; r.idz = rs.idz;
	r.ox = o.x; r.oy = o.y; r.oz = o.z;
	r.nx = rs.idx < 0.0f; r.ny = rs.idy < 0.0f; r.nz = rs.idz < 0.0f;
	const unsigned oct = (r.nx ? 1u : 0u) | (r.ny ? 2u : 0u) | (r.nz ? 4u : 0u);
	r.octinv4 = (7u - oct) * 0x01010101u;
	return r;
}
__device__ __forceinline__ float byte_f(unsigned w, int j) { return (float)((w >> (8 * j)) & 0xffu); }

__device__ __forceinline__ void wide_start(uint2 &ng, uint2 &tg, int &sp) {
	ng = make_uint2(0u, 0x80000000u);  // "group" holding only the root: base 0, hit bit 31, imask 0
	tg = make_uint2(0u, 0u);
	sp = 0;
}

// One node step for a lane whose tg is empty.  Returns false when the traversal is complete.
template <bool COUNT>
__device__ __forceinline__ bool wide_node_step(const DevScene &sc, const WideRay &r, float tmin, float tmax, uint2 &ng, uint2 &tg, int &sp, uint2 *stack,
	TravCounters *cnt) {
	if (ng.y <= 0x00ffffffu) {
		if (sp == 0) return false;
		ng = stack[--sp];
	}
	const unsigned hits = ng.y;
	const int bit = 31 - __clz(hits);
	ng.y &= ~(1u << bit);
	if (ng.y > 0x00ffffffu && sp < ARE_WIDE_STACK) stack[sp++] = ng;
	const unsigned slot = (unsigned)(bit - 24) ^ (r.octinv4 & 7u);
	const unsigned rel = __popc(hits & 0xffu & ((1u << slot) - 1u));
	const WideNode *n = sc.wnodes + (ng.x + rel);
	const float4 n0 = ldg4(&n->n0);
	const uint4 n1 = __ldg(reinterpret_cast<const uint4 *>(&n->n1));
	const uint4 n2 = __ldg(reinterpret_cast<const uint4 *>(&n->n2));
	const uint4 n3 = __ldg(reinterpret_cast<const uint4 *>(&n->n3));
	const uint4 n4 = __ldg(reinterpret_cast<const uint4 *>(&n->n4));
	if (COUNT) cnt->nodes += 4;  // eight child boxes = four "box pairs" in the units the BVH2 counter uses
	const unsigned ew = __float_as_uint(n0.w);
	// per-node grid: bound = origin + q * 2^e  =>  t = q * (2^e / d) + (origin - o) / d
	const float sx = __uint_as_float((ew & 0xffu) << 23) * r.idx, sy = __uint_as_float(((ew >> 8) & 0xffu) << 23) * r.idy,
				sz = __uint_as_float(((ew >> 16) & 0xffu) << 23) * r.idz;
	const float bx = (n0.x - r.ox) * r.idx, by = (n0.y - r.oy) * r.idy, bz = (n0.z - r.oz) * r.idz;
	unsigned hitmask = 0u;
#pragma unroll
	for (int half = 0; half < 2; ++half) {
		const unsigned meta4 = half ? n1.w : n1.z;
		const unsigned inner4 = ((meta4 & (meta4 << 1)) & 0x10101010u) >> 4;  // 1 per byte whose low five bits are >= 24
		const unsigned bit4 = (meta4 ^ (r.octinv4 & (inner4 * 0xffu))) & 0x1f1f1f1fu;
		const unsigned present4 = (meta4 >> 5) & 0x01010101u;
		const unsigned lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
		const unsigned hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
		const unsigned nearx = r.nx ? hix : lox, farx = r.nx ? lox : hix;
		const unsigned neary = r.ny ? hiy : loy, fary = r.ny ? loy : hiy;
		const unsigned nearz = r.nz ? hiz : loz, farz = r.nz ? loz : hiz;
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const float tn = fmaxf(fmaxf(fmaf(byte_f(nearx, j), sx, bx), fmaf(byte_f(neary, j), sy, by)), fmaxf(fmaf(byte_f(nearz, j), sz, bz), tmin));
			const float tf = fminf(fminf(fmaf(byte_f(farx, j), sx, bx), fmaf(byte_f(fary, j), sy, by)), fminf(fmaf(byte_f(farz, j), sz, bz), tmax));
			if (tn <= tf) hitmask |= ((present4 >> (8 * j)) & 1u) << ((bit4 >> (8 * j)) & 31u);
		}
	}
	ng = make_uint2(n1.x, (hitmask & 0xff000000u) | (ew >> 24));
	tg = make_uint2(n1.y, hitmask & 0x00ffffffu);
	return true;
}

// Test one pending leaf primitive of tg.
template <bool COUNT>
__device__ __forceinline__ void wide_leaf_step(const DevScene &sc, V3<float> o, V3<float> d, float tmin, uint2 &tg, Hit &h, TravCounters *cnt) {
	const int bit = 31 - __clz(tg.y);
	tg.y &= ~(1u << bit);
	const int slot = (int)tg.x + bit;
	const int kind = sc.wide_kinds[slot];
	const HotPrim *p = sc.wide_prims + slot;
	const float4 r0 = ldg4(&p->r0), r1 = ldg4(&p->r1);
	if (kind == 3) {
		test_sphere(r0, r1, o, d, tmin, slot, h);
		if (COUNT) cnt->spheres++;
	} else {
		const float4 r2 = ldg4(&p->r2);
		if (kind == 1) { test_plane<true>(r0, r1, r2, o, d, tmin, slot, h); if (COUNT) cnt->quads++; }
		else if (kind == 2) { test_plane<false>(r0, r1, r2, o, d, tmin, slot, h); if (COUNT) cnt->tris++; }
		else { test_box(r0, r1, r2, ldg4(&p[1].r0), o, d, tmin, slot, h); if (COUNT) cnt->boxes++; }
	}
}

// Whole traversal in one go (per-ray harness).
template <bool COUNT>
__device__ __forceinline__ void intersect_wide(const DevScene &sc, V3<float> o, V3<float> d, float tmin, Hit &h, TravCounters *cnt) {
	const WideRay r = wide_ray(o, d);
	uint2 stack[ARE_WIDE_STACK];
	uint2 ng, tg;
	int sp;
	wide_start(ng, tg, sp);
	while (true) {
		if (tg.y != 0u) wide_leaf_step<COUNT>(sc, o, d, tmin, tg, h, cnt);
		else if (!wide_node_step<COUNT>(sc, r, tmin, h.t, ng, tg, sp, stack, cnt)) break;
	}
}

}  // namespace areb
