// intersect.cuh — fp32 ray/primitive tests and scene traversal used by the render kernels (and, so that the
// harness checks the code that actually renders, by the precision=32 per-ray kernels).
//
// Ray directions are unit length (are::Ray's convention, /root/reference/src/basic/ray.cpp:5), so t is the
// geometric distance.  Triangles and parallelograms are held in the 48-byte plane form of dev_types.h: the
// same hit the reference's Möller–Trumbore routine finds (src/object/triangle.cpp:82-121 — two-sided, no
// culling), evaluated as plane distance + two planar coordinates: 15 FFMA/FMUL + one MUFU.RCP + 4 compares +
// 2 selects per test (the reference's formulation needs ~27 + RCP + 7 + 4).  A hot quad may stand for a fused
// pair of coplanar triangles forming a parallelogram; the owning triangle is recovered after the loop.
//
// The loop keeps only (t, index) of the best candidate; planar coordinates are recomputed once for the winner.
#pragma once
#include "dev_types.h"
#include "vec.cuh"

namespace areb {

struct Hit {
	float t;    // closest distance so far (INFINITY = none)
	int idx;    // index into the hot array traversed (-1 = miss)
	int orig;   // in: hot index of the primitive the ray starts ON (-1: none, e.g. camera rays)
};

__device__ __forceinline__ float4 ldg4(const f4 *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// 32 bytes in one request (sm_100: LDG.E.256): p is 32-byte aligned.  A hierarchy node is then two (BVH2) or four (BVH4)
// requests instead of four / seven — what the divergent node fetches of a traversing warp load the L1 data pipe with is
// requests x distinct lines, not bytes.
#ifndef ARE_LDG256
#define ARE_LDG256 1
#endif
#ifndef ARE_STACK_CACHE
#define ARE_STACK_CACHE 0  // top entries of the traversal stack held in registers (CachedStack below): 0, 1 or 2
#endif
__device__ __forceinline__ void ldg8(const f4 *p, float4 &a, float4 &b) {
#if ARE_LDG256
	asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
	a = ldg4(p); b = ldg4(p + 1);
#endif
}
__device__ __forceinline__ float4 lds4(const f4 *p) { return *reinterpret_cast<const float4 *>(p); }

// single MUFU.RCP; rcp(0) = inf, so a ray parallel to the plane gives t = +-inf / NaN and fails the window test
__device__ __forceinline__ float rcp_fast(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {
	float r;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// planar coordinates of point p in a plane-form record
__device__ __forceinline__ void plane_coords(float4 r1, float4 r2, V3<float> p, float &a, float &b) {
	a = fmaf(r1.x, p.x, fmaf(r1.y, p.y, fmaf(r1.z, p.z, -r1.w)));
	b = fmaf(r2.x, p.x, fmaf(r2.y, p.y, fmaf(r2.z, p.z, -r2.w)));
}

// One plane-form test. QUAD: accept alpha,beta in [0,1]; else triangle: alpha,beta >= 0, alpha+beta <= 1.
// Range checks run on the float bit patterns: x in [0,1]  <=>  bits(x) <= 0x3f800000 as unsigned (negative
// values have the sign bit set, NaN is above 0x7f800000), so two coordinates cost one UMAX + one ISETP.
// Acceptance of a plane-form candidate (distance t, planar coordinates a, b): shared with the scene-specialised
// ("baked") kernel, which computes t, a and b with the zero components of its constant vectors left out.
template <bool QUAD>
__device__ __forceinline__ void plane_accept(float t, float a, float b, float tmin, int idx, Hit &h) {
	bool ok = (t > tmin) & (t < h.t);
	if (QUAD) {
		ok = ok & (max(__float_as_uint(a), __float_as_uint(b)) <= 0x3f800000u);
	} else {
		float c = 1.0f - (a + b);
		ok = ok & ((int)(__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c)) >= 0);
	}
	if (ok) { h.t = t; h.idx = idx; }
}
template <bool QUAD>
__device__ __forceinline__ void test_plane(float4 r0, float4 r1, float4 r2, V3<float> o, V3<float> d, float tmin, int idx, Hit &h) {
	float denom = fmaf(r0.x, d.x, fmaf(r0.y, d.y, r0.z * d.z));
	float num = fmaf(-r0.x, o.x, fmaf(-r0.y, o.y, fmaf(-r0.z, o.z, r0.w)));
	float t = num * rcp_fast(denom);
	V3<float> p = mk<float>(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
	float a, b;
	plane_coords(r1, r2, p, a, b);
	plane_accept<QUAD>(t, a, b, tmin, idx, h);
}

// Sphere: r0 = (c, r), r1.x = r*r.  Discriminant from the perpendicular offset l = oc - (oc·d)d, i.e.
// disc = r^2 - |l|^2 (Haines et al., "Precision improvements for ray/sphere intersection"): no h^2 - |oc|^2
// cancellation, so small spheres far from the origin and huge spheres under the origin both keep fp32 accuracy.
__device__ __forceinline__ void test_sphere(float4 r0, float4 r1, V3<float> o, V3<float> d, float tmin, int idx, Hit &h) {
	float ox = r0.x - o.x, oy = r0.y - o.y, oz = r0.z - o.z;
	float hh = fmaf(d.x, ox, fmaf(d.y, oy, d.z * oz));
	float lx = fmaf(-hh, d.x, ox), ly = fmaf(-hh, d.y, oy), lz = fmaf(-hh, d.z, oz);
	float disc = r1.x - fmaf(lx, lx, fmaf(ly, ly, lz * lz));
	float sq = sqrt_fast(fmaxf(disc, 0.0f));
	float t0 = hh - sq, t1 = hh + sq;
	float t = (t0 > tmin) ? t0 : t1;  // nearest root inside the window (t0 <= t1)
	// A ray that starts ON this sphere has the exact roots 0 and 2(oc·d): leaving it (oc·d < 0) it cannot come back,
	// entering it the far side is at 2(oc·d).  Evaluating that case through r^2 - |l|^2 would decide on rounding noise
	// (|oc|^2 carries 0.06 absolute error for a radius-1000 ground sphere) and lets grazing rays leak INSIDE the
	// sphere, where a two-sided Lambertian surface traps them until max_depth.
	const bool self = idx == h.orig;
	t = self ? 2.0f * hh : t;
	bool ok = ((disc >= 0.0f) | self) & (t > tmin) & (t < h.t);
	if (ok) { h.t = t; h.idx = idx; }
}

// ---- parallelepiped ("box") -----------------------------------------------------------------------------
// Three slab pairs.  Per axis: s = n·d, e = c - n·o, m = e/s, k = h/|s|: the ray is inside the slab for t in
// [m-k, m+k] (already ordered, no min/max), entering through face "-" when s > 0.  A line meets a convex body's
// surface at most twice, so the only possible hits on the (two-sided) faces are the entry t_in = max(m-k) and the
// exit t_out = min(m+k); each is accepted if it lies in the (tmin, best) window and its face is present.
// The 1e-30 folded into the dot product keeps s away from an exact zero (a direction lying exactly in a face plane
// occurs with probability 2^-24 per cosine sample), so 1/s stays finite and the slab ordering below stays valid.
//
// An open box is stored with its one absent face on axis 2, "-" side (scene compiler): a ray enters through that
// hole iff the entry is decided by axis 2 while moving along +n2, and leaves through it iff the exit is decided by
// axis 2 while moving along -n2.
// OPEN: 0 = the record says (r3.w), 1 = known closed (the open-face logic is compiled out), 2 = known open
// The part of the test after the six dot products (s_i = n_i·d + 1e-30, e_i = c_i - n_i·o), shared with the baked kernel.
template <int OPEN>
__device__ __forceinline__ void test_box_se(float s0, float s1, float s2, float e0, float e1, float e2, float h0, float h1, float h2, bool open_flag,
	float tmin, int idx, Hit &h) {
	const float i0 = rcp_fast(s0), i1 = rcp_fast(s1), i2 = rcp_fast(s2);
	const float k0 = h0 * fabsf(i0), k1 = h1 * fabsf(i1), k2 = h2 * fabsf(i2);
	// slab interval of axis i: e_i/s_i -/+ h_i/|s_i|, each end ONE fused multiply-add of (e_i, 1/s_i, -/+ k_i)
	const float n0 = fmaf(e0, i0, -k0), n1 = fmaf(e1, i1, -k1), n2 = fmaf(e2, i2, -k2);
	const float f0 = fmaf(e0, i0, k0), f1 = fmaf(e1, i1, k1), f2 = fmaf(e2, i2, k2);
	const float t_in = fmaxf(fmaxf(n0, n1), n2), t_out = fminf(fminf(f0, f1), f2);
	if (OPEN == 2 || (OPEN == 0 && open_flag)) {  // warp-uniform: all lanes test the same primitive
		const bool in_ok = (t_in > tmin) & !((n2 == t_in) & (s2 > 0.0f));
		const bool out_ok = (t_out > tmin) & !((f2 == t_out) & (s2 < 0.0f));
		const float t = in_ok ? t_in : t_out;
		const bool ok = (t_in <= t_out) & (in_ok | out_ok) & (t < h.t);
		if (ok) { h.t = t; h.idx = idx; }
	} else {  // closed: the entry if it lies beyond tmin, else the exit; one window test on the chosen distance
		const float t = t_in > tmin ? t_in : t_out;
		const bool ok = (t_in <= t_out) & (t > tmin) & (t < h.t);
		if (ok) { h.t = t; h.idx = idx; }
	}
}
template <int OPEN = 0>
__device__ __forceinline__ void test_box(float4 r0, float4 r1, float4 r2, float4 r3, V3<float> o, V3<float> d, float tmin, int idx, Hit &h) {
	const float s0 = fmaf(r0.x, d.x, fmaf(r0.y, d.y, fmaf(r0.z, d.z, 1e-30f)));
	const float s1 = fmaf(r1.x, d.x, fmaf(r1.y, d.y, fmaf(r1.z, d.z, 1e-30f)));
	const float s2 = fmaf(r2.x, d.x, fmaf(r2.y, d.y, fmaf(r2.z, d.z, 1e-30f)));
	const float e0 = fmaf(-r0.x, o.x, fmaf(-r0.y, o.y, fmaf(-r0.z, o.z, r0.w)));
	const float e1 = fmaf(-r1.x, o.x, fmaf(-r1.y, o.y, fmaf(-r1.z, o.z, r1.w)));
	const float e2 = fmaf(-r2.x, o.x, fmaf(-r2.y, o.y, fmaf(-r2.z, o.z, r2.w)));
	test_box_se<OPEN>(s0, s1, s2, e0, e1, e2, r3.x, r3.y, r3.z, r3.w != 0.0f, tmin, idx, h);
}
// Which face (2*axis + side) of a box the hit point P lies on: the axis whose normalised slab coordinate
// q_i = (n_i·P - c_i)/h_i is closest to +-1, i.e. largest in magnitude; rih = (1/h_0, 1/h_1, 1/h_2).
__device__ __forceinline__ int box_hit_face(float4 r0, float4 r1, float4 r2, float4 rih, V3<float> P) {
	const float q0 = fmaf(r0.x, P.x, fmaf(r0.y, P.y, fmaf(r0.z, P.z, -r0.w))) * rih.x;
	const float q1 = fmaf(r1.x, P.x, fmaf(r1.y, P.y, fmaf(r1.z, P.z, -r1.w))) * rih.y;
	const float q2 = fmaf(r2.x, P.x, fmaf(r2.y, P.y, fmaf(r2.z, P.z, -r2.w))) * rih.z;
	const float a0 = fabsf(q0), a1 = fabsf(q1), a2 = fabsf(q2);
	// selects only (same ties as the nested comparison: axis 0 before 1 before 2)
	const bool p01 = a0 >= a1;
	const float q01 = p01 ? q0 : q1, a01 = p01 ? a0 : a1;
	const bool p2 = a01 >= a2;
	const float q = p2 ? q01 : q2;
	const int ax2 = p2 ? (p01 ? 0 : 2) : 4;  // 2 * axis
	return ax2 + (q > 0.0f ? 1 : 0);
}

// Test a type-sorted run of hot primitives. LD = ldg4 (global / L2) or lds4 (shared-memory copy).
template <float4 (*LD)(const f4 *)>
__device__ __forceinline__ void intersect_range(const HotPrim *prims, int first, int nq, int nt, int ns, int nb, V3<float> o, V3<float> d, float tmin, Hit &h) {
	int i = first;
	int end = first + 2 * nb;
#pragma unroll 2
	for (; i < end; i += 2) test_box(LD(&prims[i].r0), LD(&prims[i].r1), LD(&prims[i].r2), LD(&prims[i + 1].r0), o, d, tmin, i, h);
	end += nq;
#pragma unroll 4
	for (; i < end; ++i) test_plane<true>(LD(&prims[i].r0), LD(&prims[i].r1), LD(&prims[i].r2), o, d, tmin, i, h);
	end += nt;
#pragma unroll 4
	for (; i < end; ++i) test_plane<false>(LD(&prims[i].r0), LD(&prims[i].r1), LD(&prims[i].r2), o, d, tmin, i, h);
	end += ns;
#pragma unroll 4
	for (; i < end; ++i) test_sphere(LD(&prims[i].r0), LD(&prims[i].r1), o, d, tmin, i, h);
}

// The lean form (scene.h): at most LEAN_MAX boxes, quad tests and triangle tests, no spheres, boxes from slot 0.
// A guarded full unroll: the counts are kernel parameters, so every guard is a uniform branch (UISETP + BRA.U) and every
// shared-memory offset a compile-time constant; the tests run in list order, exactly as intersect_range runs them.
// (Measured alternative: one `switch` per kind into count-specialised runs — the indirect BRX jumps cost more than the
// twelve guards, 7762 vs 7999 Msamples/s on the Cornell box.)
// `sb` = 32-bit shared-memory address of slot 0: explicit ld.shared keeps the base in one register (through generic
// pointers the compiler re-derives the shared window — S2UR CgaCtaId, three ALU ops — before every group of loads).
__device__ __forceinline__ float4 lds4a(uint32_t addr) {
	float4 v;
	asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ int lds1a(uint32_t addr) {
	int v;
	asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void intersect_lean(uint32_t sb, int nb, int n_open, int nq, int nt, V3<float> o, V3<float> d, float tmin, Hit &h) {
#pragma unroll
	for (int i = 0; i < LEAN_MAX; ++i)
		if (i < nb) {
			const float4 r0 = lds4a(sb + 96 * i), r1 = lds4a(sb + 96 * i + 16), r2 = lds4a(sb + 96 * i + 32), r3 = lds4a(sb + 96 * i + 48);
#ifdef LEAN_NO_OPEN_SPLIT  // A/B: the record's own flag, predicated open-face logic in every box test
			test_box<0>(r0, r1, r2, r3, o, d, tmin, 2 * i, h);
#else
			if (i < n_open) test_box<2>(r0, r1, r2, r3, o, d, tmin, 2 * i, h);  // open boxes come first in the list
			else test_box<1>(r0, r1, r2, r3, o, d, tmin, 2 * i, h);
#endif
		}
	const uint32_t qb = sb + 96 * nb;
#pragma unroll
	for (int j = 0; j < LEAN_MAX; ++j)
		if (j < nq) test_plane<true>(lds4a(qb + 48 * j), lds4a(qb + 48 * j + 16), lds4a(qb + 48 * j + 32), o, d, tmin, 2 * nb + j, h);
	const uint32_t tb = qb + 48 * nq;
#pragma unroll
	for (int j = 0; j < LEAN_MAX; ++j)
		if (j < nt) test_plane<false>(lds4a(tb + 48 * j), lds4a(tb + 48 * j + 16), lds4a(tb + 48 * j + 32), o, d, tmin, 2 * nb + nq + j, h);
}

// ---- BVH2 traversal, per-thread stack -----------------------------------------------------------------

struct TravCounters {
	unsigned long long nodes, quads, tris, spheres, boxes;
};

// Ray constants of the slab test: reciprocal direction (a zero component becomes a huge finite slope so the slab
// maths never sees 0*inf) and origin * reciprocal.
struct RaySlopes {
	float idx, idy, idz, oxi, oyi, ozi;
	float ax, ay, az;  // |reciprocal direction|
};
__device__ __forceinline__ RaySlopes ray_slopes(V3<float> o, V3<float> d) {
	const float big = 1e30f;
	RaySlopes r;
	r.idx = fabsf(d.x) > 1e-30f ? 1.0f / d.x : (d.x < 0 ? -big : big);
	r.idy = fabsf(d.y) > 1e-30f ? 1.0f / d.y : (d.y < 0 ? -big : big);
	r.idz = fabsf(d.z) > 1e-30f ? 1.0f / d.z : (d.z < 0 ? -big : big);
	r.oxi = o.x * r.idx; r.oyi = o.y * r.idy; r.ozi = o.z * r.idz;
	r.ax = fabsf(r.idx); r.ay = fabsf(r.idy); r.az = fabsf(r.idz);
	return r;
}

// BVH leaves hold ONE hot primitive each; a negative child reference encodes it completely:
//   ref = ~(slot | kind << 29),  kind: 0 box (two slots), 1 quad / fused pair, 2 triangle, 3 sphere.
// No per-leaf loops, no counts to fetch: a leaf test is one switch on the kind.
__device__ __forceinline__ int leaf_slot(int ref) { return (~ref) & 0x1fffffff; }
template <bool COUNT>
__device__ __forceinline__ void test_leaf(const DevScene &sc, int ref, V3<float> o, V3<float> d, float tmin, Hit &h, TravCounters *cnt) {
	const unsigned code = (unsigned)~ref;
	const int slot = (int)(code & 0x1fffffffu), kind = (int)(code >> 29);
	const HotPrim *p = sc.bvh_prims + slot;
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

// ---- traversal step -------------------------------------------------------------------------------------
// The continuation logic is what a BVH2 step spends most of its (half-rate) ALU-pipe instructions on (~30 of them in the
// first formulation of this round: separate "node" and "pending leaf" words, a stack index with overflow and emptiness
// tests; RTIOW +18 % without them, profiles/r01_scenes.md), so it is kept to the minimum: ONE cursor word (>= 0 inner node, < 0 leaf reference, TRAV_DONE finished), a stack POINTER instead of
// an index (no address arithmetic per push / pop), a TRAV_DONE sentinel at the bottom of the stack (a pop never checks
// for emptiness) and no overflow test (are_cuda_commit refuses hierarchies deeper than the stack).  A child that is
// missed gets distance +inf, so "nearer child" also covers the one-hit case:
//   near / far by two selects, push far if both were hit, pop if neither was.
#define TRAV_DONE ((int)0x80000000)  // = ~(slot 2^29 - 1 | kind 3 << 29): the one leaf reference that cannot occur (hot arrays are far smaller)
#define TRAV_DONE_V ((int)0x80000000)  // = TRAV_DONE (below)
// Where the postponed children live.  PtrStack: a per-thread array in local memory walked by a pointer (the default).
// ShortStack (north_star item 2, built with -DARE_SHORT_STACK=n): the first n entries of every thread sit in shared memory
// (entry e of thread t at s[e][t]: conflict-free) and only a traversal that postpones more than n children at once
// touches the local-memory overflow array.  Measured against PtrStack in profiles/r02_short_stack.md.
struct PtrStack {
	int *top;
	__device__ __forceinline__ void reset(int *base) { base[0] = TRAV_DONE_V; top = base + 1; }
	__device__ __forceinline__ void push(int v) { *top++ = v; }
	__device__ __forceinline__ int pop() { return *--top; }
};
// PtrStack with its top entry (ARE_STACK_CACHE == 1) or top two entries (== 2) held in registers.  The local-memory
// stack of a traversing warp is addressed lane by lane at different depths, so every push / pop request costs the L1 data
// pipe one wavefront per distinct depth — and that pipe, not latency, is what bounds the 1 M-primitive scene (91–96 %
// busy, profiles/r02_l1_pipe.md).  A push directly followed by a pop (both children hit, the nearer subtree then missed)
// never reaches memory.
#define TRAV_EMPTY ((int)0x80000001)  // = ~(slot 2^29 - 2 | kind 3 << 29): cannot occur as a reference
struct CachedStack {
	int *top;
	int c0;
#if ARE_STACK_CACHE >= 2
	int c1;
#endif
	__device__ __forceinline__ void reset(int *base) {
		base[0] = TRAV_DONE_V; top = base + 1; c0 = TRAV_EMPTY;
#if ARE_STACK_CACHE >= 2
		c1 = TRAV_EMPTY;
#endif
	}
#if ARE_STACK_CACHE >= 2
	// c0 = top of stack, c1 = the entry under it; memory holds the rest
	__device__ __forceinline__ void push(int v) {
		if (c1 != TRAV_EMPTY) *top++ = c1;
		c1 = c0; c0 = v;
	}
	__device__ __forceinline__ int pop() {
		int r = c0;
		c0 = c1; c1 = TRAV_EMPTY;
		if (r == TRAV_EMPTY) r = *--top;
		return r;
	}
#else
	__device__ __forceinline__ void push(int v) {
		if (c0 != TRAV_EMPTY) *top++ = c0;
		c0 = v;
	}
	__device__ __forceinline__ int pop() {
		int r = c0;
		c0 = TRAV_EMPTY;
		if (r == TRAV_EMPTY) r = *--top;
		return r;
	}
#endif
};
template <int N, int THREADS>
struct ShortStack {
	int *deep;      // local-memory overflow, entries N, N + 1, ...
	int *sm;        // &s[0][thread]
	int sp;
	__device__ __forceinline__ void reset(int *base) { deep = base; sp = 1; sm[0] = TRAV_DONE_V; }
	__device__ __forceinline__ void push(int v) {
		if (sp < N) sm[sp * THREADS] = v;
		else deep[sp - N] = v;
		++sp;
	}
	__device__ __forceinline__ int pop() {
		--sp;
		return sp < N ? sm[sp * THREADS] : deep[sp - N];
	}
};
template <bool COUNT, class STK>
__device__ __forceinline__ void bvh_step(const DevScene &sc, float tmin, const RaySlopes &rs, int &cur, STK &stk, const Hit &h, TravCounters *cnt) {
	const BvhNode *n = sc.nodes + cur;
	float4 b0, b1, b2, cm;  // cm = (child[0], child[1], meta[0], meta[1]) as bits
	ldg8(&n->b0, b0, b1);
	ldg8(&n->b2, b2, cm);
	const int2 ch = make_int2(__float_as_int(cm.x), __float_as_int(cm.y));
	if (COUNT) cnt->nodes++;
	const float c0x = fmaf(b0.x, rs.idx, -rs.oxi), c0y = fmaf(b0.z, rs.idy, -rs.oyi), c0z = fmaf(b2.x, rs.idz, -rs.ozi);
	const float c1x = fmaf(b1.x, rs.idx, -rs.oxi), c1y = fmaf(b1.z, rs.idy, -rs.oyi), c1z = fmaf(b2.z, rs.idz, -rs.ozi);
	const float t0n = fmaxf(fmaxf(fmaf(-b0.y, rs.ax, c0x), fmaf(-b0.w, rs.ay, c0y)), fmaxf(fmaf(-b2.y, rs.az, c0z), tmin));
	const float t0f = fminf(fminf(fmaf(b0.y, rs.ax, c0x), fmaf(b0.w, rs.ay, c0y)), fminf(fmaf(b2.y, rs.az, c0z), h.t));
	const float t1n = fmaxf(fmaxf(fmaf(-b1.y, rs.ax, c1x), fmaf(-b1.w, rs.ay, c1y)), fmaxf(fmaf(-b2.w, rs.az, c1z), tmin));
	const float t1f = fminf(fminf(fmaf(b1.y, rs.ax, c1x), fmaf(b1.w, rs.ay, c1y)), fminf(fmaf(b2.w, rs.az, c1z), h.t));
	const bool hit0 = t0n <= t0f, hit1 = t1n <= t1f;
	const float d0 = hit0 ? t0n : INFINITY, d1 = hit1 ? t1n : INFINITY;
	const bool near0 = d0 <= d1;
	int next = near0 ? ch.x : ch.y;
	const int farc = near0 ? ch.y : ch.x;
	if (hit0 & hit1) stk.push(farc);  // (measured: prefetch.global.L1 of the postponed child's node here: -1.3 % / -3.5 %)
	if (!(hit0 | hit1)) next = stk.pop();
	cur = next;
}
// ---- quantised nodes (dev_types.h: BvhNodeQ) ----------------------------------------------------------------
// A 16-bit plane index q becomes the float 2^23 + q by ONE byte permute (0x4b000000 | q), and the slab distance
//   t = (lo + q step - o) / d = (2^23 + q) * A + B',   A = step / d,  B' = (lo - o) / d - 2^23 A
// by ONE multiply-add — the de-quantisation costs nothing beyond the permute.  Which 16 bits of a (lo, hi) word are the
// near plane is a per-ray permute selector (the sign of d), so there are no per-axis min / max either.
// Rounding: B' is rounded at the magnitude of max(2^23 A, |lo - o| / d), i.e. to half a grid step for ray origins within
// 128 grid extents of the grid — the builder pads every box by a whole step — and beyond that to the same ulp of the
// distance that the fp32 step's c / d - o / d carries: never worse than the fp32 nodes by more than the padding covers.
struct RaySlopesQ {
	float ax, ay, az, bx, by, bz;
	unsigned sx, sy, sz;  // permute selector of the near plane: 0x7610 (low half) when d >= 0, 0x7632 (high half) otherwise
};
__device__ __forceinline__ RaySlopesQ ray_slopes_q(const QGrid *g, V3<float> o, V3<float> d) {
	const float big = 1e18f;
	const float4 g0 = __ldg(reinterpret_cast<const float4 *>(g)), g1 = __ldg(reinterpret_cast<const float4 *>(g) + 1);  // lo.xyz step.x | step.yz
	const float ix = fabsf(d.x) > 1e-18f ? 1.0f / d.x : (d.x < 0 ? -big : big);
	const float iy = fabsf(d.y) > 1e-18f ? 1.0f / d.y : (d.y < 0 ? -big : big);
	const float iz = fabsf(d.z) > 1e-18f ? 1.0f / d.z : (d.z < 0 ? -big : big);
	RaySlopesQ r;
	r.ax = g0.w * ix; r.ay = g1.x * iy; r.az = g1.y * iz;
	r.bx = fmaf(-8388608.0f, r.ax, (g0.x - o.x) * ix);
	r.by = fmaf(-8388608.0f, r.ay, (g0.y - o.y) * iy);
	r.bz = fmaf(-8388608.0f, r.az, (g0.z - o.z) * iz);
	r.sx = ix >= 0.0f ? 0x7610u : 0x7632u;
	r.sy = iy >= 0.0f ? 0x7610u : 0x7632u;
	r.sz = iz >= 0.0f ? 0x7610u : 0x7632u;
	return r;
}
#define ARE_QPLANE(w, sel) __uint_as_float(__byte_perm((w), 0x4b000000u, (sel)))
template <bool COUNT, class STK>
__device__ __forceinline__ void bvhq_step(const DevScene &sc, float tmin, const RaySlopesQ &rs, int &cur, STK &stk, const Hit &h, TravCounters *cnt) {
	float4 w0, w1;  // bits: (c0.x, c0.y, c0.z, c1.x) (c1.y, c1.z, child[0], child[1])
	ldg8(reinterpret_cast<const f4 *>(sc.nodes_q + cur), w0, w1);
	const unsigned q0x = __float_as_uint(w0.x), q0y = __float_as_uint(w0.y), q0z = __float_as_uint(w0.z);
	const unsigned q1x = __float_as_uint(w0.w), q1y = __float_as_uint(w1.x), q1z = __float_as_uint(w1.y);
	const int2 ch = make_int2(__float_as_int(w1.z), __float_as_int(w1.w));
	if (COUNT) cnt->nodes++;
	const unsigned fx = rs.sx ^ 0x22u, fy = rs.sy ^ 0x22u, fz = rs.sz ^ 0x22u;
	const float t0n = fmaxf(fmaxf(fmaf(ARE_QPLANE(q0x, rs.sx), rs.ax, rs.bx), fmaf(ARE_QPLANE(q0y, rs.sy), rs.ay, rs.by)), fmaxf(fmaf(ARE_QPLANE(q0z, rs.sz), rs.az, rs.bz), tmin));
	const float t0f = fminf(fminf(fmaf(ARE_QPLANE(q0x, fx), rs.ax, rs.bx), fmaf(ARE_QPLANE(q0y, fy), rs.ay, rs.by)), fminf(fmaf(ARE_QPLANE(q0z, fz), rs.az, rs.bz), h.t));
	const float t1n = fmaxf(fmaxf(fmaf(ARE_QPLANE(q1x, rs.sx), rs.ax, rs.bx), fmaf(ARE_QPLANE(q1y, rs.sy), rs.ay, rs.by)), fmaxf(fmaf(ARE_QPLANE(q1z, rs.sz), rs.az, rs.bz), tmin));
	const float t1f = fminf(fminf(fmaf(ARE_QPLANE(q1x, fx), rs.ax, rs.bx), fmaf(ARE_QPLANE(q1y, fy), rs.ay, rs.by)), fminf(fmaf(ARE_QPLANE(q1z, fz), rs.az, rs.bz), h.t));
	const bool hit0 = t0n <= t0f, hit1 = t1n <= t1f;
	const float d0 = hit0 ? t0n : INFINITY, d1 = hit1 ? t1n : INFINITY;
	const bool near0 = d0 <= d1;
	int next = near0 ? ch.x : ch.y;
	const int farc = near0 ? ch.y : ch.x;
	if (hit0 & hit1) stk.push(farc);
	if (!(hit0 | hit1)) next = stk.pop();
	cur = next;
}
// The node phase of a traversal MODE (render_path.cuh): 1 = BVH2 (fp32 nodes), 3 = BVH4, 4 = BVH2 over quantised nodes.
template <int MODE> struct Trav {
	typedef RaySlopes Slopes;
	static __device__ __forceinline__ Slopes slopes(const DevScene &, V3<float> o, V3<float> d) { return ray_slopes(o, d); }
	template <bool COUNT, class STK>
	static __device__ __forceinline__ void step(const DevScene &sc, float tmin, const Slopes &rs, int &cur, STK &stk, const Hit &h, TravCounters *cnt) {
		bvh_step<COUNT>(sc, tmin, rs, cur, stk, h, cnt);
	}
};
template <> struct Trav<4> {
	typedef RaySlopesQ Slopes;
	static __device__ __forceinline__ Slopes slopes(const DevScene &sc, V3<float> o, V3<float> d) { return ray_slopes_q(sc.qgrid, o, d); }
	template <bool COUNT, class STK>
	static __device__ __forceinline__ void step(const DevScene &sc, float tmin, const Slopes &rs, int &cur, STK &stk, const Hit &h, TravCounters *cnt) {
		bvhq_step<COUNT>(sc, tmin, rs, cur, stk, h, cnt);
	}
};

// One BVH4 visit: four slab tests on one 128-byte node.  The nearest hit child becomes the cursor, the other hit
// children are postponed (slot order).  Nearest = minimum over keys (bits of the entry distance with the slot number
// in the two low mantissa bits): distances are positive, so their bit patterns order like the floats.
template <bool COUNT, class STK>
__device__ __forceinline__ void bvh4_step(const DevScene &sc, float tmin, const RaySlopes &rs, int &cur, STK &stk, const Hit &h, TravCounters *cnt) {
	const Bvh4Node *n = sc.nodes4 + cur;
	float4 cx, hx, cy, hy, cz, hz;
	ldg8(&n->cx, cx, hx);
	ldg8(&n->cy, cy, hy);
	ldg8(&n->cz, cz, hz);
	const int4 ch = __ldg(reinterpret_cast<const int4 *>(&n->child[0]));
	if (COUNT) cnt->nodes += 2;  // node_visits counts child-box PAIRS
#define ARE_SLAB4(k, C)                                                                                                      \
	const float ax##k = fmaf(cx.C, rs.idx, -rs.oxi), ay##k = fmaf(cy.C, rs.idy, -rs.oyi), az##k = fmaf(cz.C, rs.idz, -rs.ozi); \
	const float tn##k = fmaxf(fmaxf(fmaf(-hx.C, rs.ax, ax##k), fmaf(-hy.C, rs.ay, ay##k)), fmaxf(fmaf(-hz.C, rs.az, az##k), tmin)); \
	const float tf##k = fminf(fminf(fmaf(hx.C, rs.ax, ax##k), fmaf(hy.C, rs.ay, ay##k)), fminf(fmaf(hz.C, rs.az, az##k), h.t)); \
	const bool hit##k = tn##k <= tf##k;                                                                                      \
	const unsigned key##k = hit##k ? ((__float_as_uint(tn##k) & ~3u) | k##u) : 0xffffffffu;
	ARE_SLAB4(0, x) ARE_SLAB4(1, y) ARE_SLAB4(2, z) ARE_SLAB4(3, w)
#undef ARE_SLAB4
	const unsigned kmin = min(min(key0, key1), min(key2, key3));
	if (kmin == 0xffffffffu) { cur = stk.pop(); return; }
	const unsigned nearest = kmin & 3u;
	if (hit3 & (nearest != 3u)) stk.push(ch.w);
	if (hit2 & (nearest != 2u)) stk.push(ch.z);
	if (hit1 & (nearest != 1u)) stk.push(ch.y);
	if (hit0 & (nearest != 0u)) stk.push(ch.x);
	cur = nearest == 0u ? ch.x : (nearest == 1u ? ch.y : (nearest == 2u ? ch.z : ch.w));
}
template <> struct Trav<3> {
	typedef RaySlopes Slopes;
	static __device__ __forceinline__ Slopes slopes(const DevScene &, V3<float> o, V3<float> d) { return ray_slopes(o, d); }
	template <bool COUNT, class STK>
	static __device__ __forceinline__ void step(const DevScene &sc, float tmin, const Slopes &rs, int &cur, STK &stk, const Hit &h, TravCounters *cnt) {
		bvh4_step<COUNT>(sc, tmin, rs, cur, stk, h, cnt);
	}
};
// leaf phase of the single-cursor form: test the leaf under the cursor, then pop
template <bool COUNT, class STK>
__device__ __forceinline__ void bvh_leaf(const DevScene &sc, V3<float> o, V3<float> d, float tmin, int &cur, STK &stk, Hit &h, TravCounters *cnt) {
	test_leaf<COUNT>(sc, cur, o, d, tmin, h, cnt);
	cur = stk.pop();
}

// Whole traversal in one go (per-ray harness).
template <bool COUNT>
__device__ __forceinline__ void intersect_bvh(const DevScene &sc, V3<float> o, V3<float> d, float tmin, Hit &h, TravCounters *cnt) {
	if (sc.n_nodes == 0) {
		if (sc.root_leaf_meta != 0) test_leaf<COUNT>(sc, sc.root_leaf_meta, o, d, tmin, h, cnt);  // a one-primitive scene
		return;
	}
	const RaySlopes rs = ray_slopes(o, d);
	int stack[ARE_BVH_STACK];
	PtrStack stk;
	stk.reset(stack);
	int cur = 0;
	while (cur != TRAV_DONE) {
		if (cur >= 0) bvh_step<COUNT>(sc, tmin, rs, cur, stk, h, cnt);
		else bvh_leaf<COUNT>(sc, o, d, tmin, cur, stk, h, cnt);
	}
}

template <bool COUNT>
__device__ __forceinline__ void intersect_bvh4(const DevScene &sc, V3<float> o, V3<float> d, float tmin, Hit &h, TravCounters *cnt) {
	if (sc.n_nodes == 0 || !sc.nodes4) { intersect_bvh<COUNT>(sc, o, d, tmin, h, cnt); return; }
	const RaySlopes rs = ray_slopes(o, d);
	int stack[ARE_BVH4_STACK];
	PtrStack stk;
	stk.reset(stack);
	int cur = 0;
	while (cur != TRAV_DONE) {
		if (cur >= 0) bvh4_step<COUNT>(sc, tmin, rs, cur, stk, h, cnt);
		else bvh_leaf<COUNT>(sc, o, d, tmin, cur, stk, h, cnt);
	}
}

// ---- map a hot hit back to the user primitive ---------------------------------------------------------
struct Resolved {
	int dev_prim;   // device primitive index (device order: triangles, quads, spheres)
	float a, b;     // planar coordinates in that primitive's own frame (0 for spheres)
};
// The user primitives behind hot slot idx. `hot` = the array (shared or global) the index refers to — needed again
// only for boxes, whose face is found from the hit point.
template <float4 (*LD)(const f4 *)>
__device__ __forceinline__ HotIds hit_ids(const DevScene &sc, const HotPrim *hot, const HotIds *ids, int idx, V3<float> P) {
	HotIds id = ids[idx];
	if (id.a < 0) {  // box: continue with the parallelogram (fused pair or quad) behind the face that was hit
		const int face = box_hit_face(LD(&hot[idx].r0), LD(&hot[idx].r1), LD(&hot[idx].r2), LD(&hot[idx + 1].r1), P);
		id = sc.box_faces[6 * (-1 - id.a) + face];
	}
	return id;
}
// Exact owner + its planar coordinates: for a fused pair of coplanar triangles, P belongs to triangle id.a (the
// lower user id, which also owns the shared diagonal) iff its barycentrics there are all inside, else to the other.
__device__ __forceinline__ Resolved resolve_exact(const DevScene &sc, HotIds id, V3<float> P) {
	Resolved r;
	r.dev_prim = id.a;
	r.a = 0.0f;
	r.b = 0.0f;
	if (id.a < sc.n_tri + sc.n_quad) {
		const HotPrim &ta = sc.prim_plane[id.a];
		plane_coords(*reinterpret_cast<const float4 *>(&ta.r1), *reinterpret_cast<const float4 *>(&ta.r2), P, r.a, r.b);
		if (id.b != -1) {
			const bool in_a = (r.a >= 0.0f) & (r.b >= 0.0f) & (r.a + r.b <= 1.0f);
			if (!in_a) {
				r.dev_prim = id.b >= 0 ? id.b : -2 - id.b;
				const HotPrim &tb = sc.prim_plane[r.dev_prim];
				plane_coords(*reinterpret_cast<const float4 *>(&tb.r1), *reinterpret_cast<const float4 *>(&tb.r2), P, r.a, r.b);
			}
		}
	}
	return r;
}

}  // namespace areb
