// render_path.cuh — the path-tracing megakernel's body (device code only; render.cu wraps it in the precompiled
// kernels, bake.cpp compiles it again with NVRTC around a scene-specific intersect_baked()).
#pragma once
#include "dev_types.h"
#include "intersect.cuh"
#include "philox.cuh"
#include "render_args.h"
#include "shade.cuh"
#include "vec.cuh"
#include "wide.cuh"

namespace areb {

typedef V3<float> F3;

#ifndef RENDER_THREADS
#define RENDER_THREADS 128  // 4 warps = a 16x8 pixel tile per CTA
#endif
#define BRUTE_MAX_PRIMS 512  // 24 KB of shared memory; larger scenes traverse the BVH

__device__ __forceinline__ F3 background(const RenderArgs &A, F3 d) {
	const float a = 0.5f * (d.y + 1.0f), b = 1.0f - a;
	return mk<float>(fmaf(b, A.bg_bottom[0], a * A.bg_top[0]), fmaf(b, A.bg_bottom[1], a * A.bg_top[1]), fmaf(b, A.bg_bottom[2], a * A.bg_top[2]));
}
__device__ __forceinline__ bool finite3(F3 v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }

// A warp owns an 8x4 pixel tile, a block (4 warps) a 16x8 tile: coherent primary rays and texture reads.
__device__ __forceinline__ void warp_tile_origin(int W, int &x0, int &y0) {
	const int tiles_x = (W + 15) >> 4;  // 16x8 super-tiles of four warp tiles, so neighbouring warps share texels / BVH nodes in L1
	const int gw = blockIdx.x * (RENDER_THREADS / 32) + (threadIdx.x >> 5);
	const int st = gw >> 2, sub = gw & 3;
	x0 = (st % tiles_x) * 16 + (sub & 1) * 8;
	y0 = (st / tiles_x) * 8 + (sub >> 1) * 4;
}

// Path-tracing megakernel.
//
// Work distribution: the warp's tile x sample range is a pool of 32*s_count (pixel, sample) tasks, task k = pixel
// (k & 31) of the tile, sample (k >> 5).  A lane whose path ends pulls the next task with a ballot/popc ticket in
// the same loop trip (path regeneration), so all 32 lanes stay busy until the pool is empty; finished samples are
// added to the tile's accumulators in shared memory (one RED.ADD.F32 x3 per sample).  Each trip = trace one ray
// segment for every lane, classify, then ONE Philox call per lane that feeds either the camera (new path) or the
// material scatter (continuing path).
#ifndef RENDER_MIN_BLOCKS
#define RENDER_MIN_BLOCKS 5  // generic brute-force kernel: 96 registers at 4 / 5 CTAs per SM, 72 at 7; textured scene, 4 / 5 / 7: 12 791 / 12 737 / 11 013 Msamples/s
#endif
#ifndef RENDER_MIN_BLOCKS_BVH2
#define RENDER_MIN_BLOCKS_BVH2 7  // single-cursor BVH2 traversal, L1-resident hierarchies: 72 registers; measured 6 / 7 / 8 CTAs/SM on RTIOW: 4207 / 4359 / 4211 Msamples/s
#endif
#ifndef RENDER_MIN_BLOCKS_LEAN
#define RENDER_MIN_BLOCKS_LEAN 8
#endif
#ifndef RENDER_MIN_BLOCKS_BIG
#define RENDER_MIN_BLOCKS_BIG 12  // hierarchies that live in L2 (not L1) are latency-bound: 48 resident warps/SM at 40 registers (with spills)
#endif                            // beat 24 at 80 — measured on the 1 M-primitive scene: 467 -> 577 Msamples/s; RTIOW (L1-resident) loses 5 %
#ifndef RENDER_MIN_BLOCKS_Q
#define RENDER_MIN_BLOCKS_Q 12    // big hierarchies over quantised nodes
#endif
#ifndef RENDER_MIN_BLOCKS_BVH4
#define RENDER_MIN_BLOCKS_BVH4 6       // 4-wide nodes: 24 box floats + 4 references in flight per step
#endif
#ifndef RENDER_MIN_BLOCKS_BVH4_BIG
#define RENDER_MIN_BLOCKS_BVH4_BIG 8
#endif
#ifndef TRAV_LEAF_EVERY
#define TRAV_LEAF_EVERY 1  // node phases per leaf phase inside a slice (must divide TRAV_STEPS_PER_VOTE)
#endif
#ifndef BVH_BIG_NODES
#define BVH_BIG_NODES 8192        // beyond this: the high-occupancy build over quantised nodes.  Crossover measured on the random-cloud scene at
                                  // 3 k / 6 k / 12 k / 16 k / 24 k / 50 k / 200 k / 1 M primitives, quantised against the L1-resident fp32 kernel:
                                  // -8 % / -2 % / +7 % / +10 % / (against the big fp32 kernel:) +18 % / +30 % / +53 % / +66 %
#endif
#ifndef TRAV_MIN_LANES
#define TRAV_MIN_LANES 6   // BVH slices end when fewer lanes than this are still traversing and others are waiting
#endif
// ... of the traversal over quantised nodes (MODE 4: big hierarchies, long rays).  1 M primitives at 16 / 4 spp per launch,
// (steps per vote, min lanes): (4, 6) 1 110 / 988, (8, 6) 1 153 / 1 016, (16, 6) 1 173 / 1 031, (8, 20) 1 166 / 1 050,
// (16, 16) 1 175 / 1 047, (16, 20) 1 187 / 1 068, (16, 24) 1 175 / 1 071, (24, 20) 1 202 / 1 074, (32, 20) 1 200 / 1 073
#ifndef TRAV_MIN_LANES_Q
#define TRAV_MIN_LANES_Q 20
#endif
#ifndef TRAV_STEPS_PER_VOTE_Q
#define TRAV_STEPS_PER_VOTE_Q 16
#endif
#ifndef TRAV_LEAF_EVERY_Q
#define TRAV_LEAF_EVERY_Q 2   // node phases per leaf phase (must divide TRAV_STEPS_PER_VOTE_Q); 1 / 2 / 4 / 8 on the random cloud at 1 M primitives:
                              // 1 190 / 1 235 / 1 214 / 1 230, at 50 k: 3 657 / 3 777 / 3 771 / 3 828, at 12 k: 6 187 / 6 397 / 6 258 / 6 437
#endif
#ifndef TRAV_STEPS_PER_VOTE
#define TRAV_STEPS_PER_VOTE 4  // measured 1 / 2 / 4 / 8: RTIOW 3498 / 3582 / 3596 / 3417, 1 M primitives 584 / 591 / 602 / 607 Msamples/s
#endif
// (Postponing leaf tests until several lanes hold one was measured in the first formulation of the traversal: testing a
// leaf as soon as a lane has it was best on RTIOW and on the 1 M-primitive scene, and needs no votes.)
// MODE: 0 = brute force from shared memory, 1 = BVH2, 2 = compressed 8-wide BVH, 3 = uncompressed 4-wide BVH (the BVH2's
// traversal loop with bvh4_step as its node phase), 4 = BVH2 over the quantised 32-byte nodes (bvhq_step; big hierarchies)
// LEAN (brute force only): the scene compiler's lean form (scene.h: at most LEAN_MAX boxes / quad tests / triangle
// tests, no spheres, every surface shaded from its ShadeRec alone).  The tests are a guarded full unroll with
// compile-time shared-memory offsets instead of four counted loops, and a hit goes straight to its shading record
// (six per box, one per face) held in shared memory: no id chain, no owner resolution, no general material path.
// BAKED (brute force; lean or generic shading): the closest-hit tests are not read from shared memory but are straight-line code generated from the
// committed scene and compiled at commit time (bake.cpp): intersect_baked() with every primitive constant an immediate
// and the zero components of its normals left out.  Everything else is the lean kernel.
// NOISE = false: the scene has no noise texture, the cooperative turbulence stage is compiled out (its live values cost
// the BVH kernels registers: RTIOW lost 4 % with the stage merely present).
template <int MODE, bool COUNT, bool BIG = false, bool LEAN = false, bool BAKED = false, bool NOISE = true>
__device__ __forceinline__ void render_path_body(const RenderArgs &A) {
	constexpr bool BVH = MODE != 0, WIDE = MODE == 2;
	static_assert(!LEAN || MODE == 0, "the lean form is a brute-force list");
	static_assert(!BAKED || MODE == 0, "a baked kernel tests a brute-force list");
	extern __shared__ float4 s_raw[];
	__shared__ float s_acc[RENDER_THREADS / 32][96];
	__shared__ float4 s_turb_q[RENDER_THREADS / 32][(!LEAN && NOISE) ? 32 : 1];  // turbulence_coop's queue and sums, per warp
	__shared__ float s_turb_sum[RENDER_THREADS / 32][(!LEAN && NOISE) ? 32 : 1];
	const HotPrim *s_prims = reinterpret_cast<const HotPrim *>(s_raw);
	// LEAN shared-memory layout behind the hot slots: [ShadeRec per record][frame per record: (tangent,0) (bitangent,0)][record base per slot]
	const uint32_t sb_prims = (uint32_t)__cvta_generic_to_shared(s_raw);
	const uint32_t sb_shade = sb_prims + 48u * A.sc.n_hot, sb_frame = sb_shade + 32u * A.sc.n_lean_shade, sb_sbase = sb_frame + 32u * A.sc.n_lean_shade;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = lane; i < 96; i += 32) s_acc[warp][i] = 0.0f;
	// The warp's shared-memory areas as 32-bit shared addresses held in registers (the empty asm makes them opaque: left as
	// expressions of threadIdx.x they are re-derived inside the loop, see shade.cuh: turbulence_coop)
	uint32_t acc_sb = (uint32_t)__cvta_generic_to_shared(&s_acc[warp][0]);
	uint32_t turb_q_sb = (uint32_t)__cvta_generic_to_shared(&s_turb_q[warp][0]), turb_sum_sb = (uint32_t)__cvta_generic_to_shared(&s_turb_sum[warp][0]);
	if (!LEAN) asm volatile("" : "+r"(acc_sb), "+r"(turb_q_sb), "+r"(turb_sum_sb));  // (the lean / baked kernel never re-derived them, and loses 1.6 % with the register pinned)
	if (!BVH) {
		const float4 *src = reinterpret_cast<const float4 *>(A.sc.brute);
		const int n4 = A.sc.n_hot * 3;
		for (int i = threadIdx.x; i < n4; i += RENDER_THREADS) s_raw[i] = __ldg(src + i);
		if (LEAN) {
			const float4 *ssrc = reinterpret_cast<const float4 *>(A.sc.lean_shade);
			float4 *sdst = s_raw + n4, *fdst = sdst + 2 * A.sc.n_lean_shade;
			for (int i = threadIdx.x; i < 2 * A.sc.n_lean_shade; i += RENDER_THREADS) sdst[i] = __ldg(ssrc + i);
			for (int i = threadIdx.x; i < A.sc.n_lean_shade; i += RENDER_THREADS) {  // the cosine lobe's frame, once per record instead of once per bounce
				const float4 n = __ldg(ssrc + 2 * i);
				F3 tg, bt;
				tangent_frame(mk<float>(n.x, n.y, n.z), tg, bt);
				fdst[2 * i] = make_float4(tg.x, tg.y, tg.z, 0.f);
				fdst[2 * i + 1] = make_float4(bt.x, bt.y, bt.z, 0.f);
			}
			int *bdst = reinterpret_cast<int *>(fdst + 2 * A.sc.n_lean_shade);
			for (int i = threadIdx.x; i < A.sc.n_hot; i += RENDER_THREADS) bdst[i] = __ldg(A.sc.lean_sbase + i);
		}
	}
	__syncthreads();
	int x0, y0;
	warp_tile_origin(A.W, x0, y0);
	const CamT<float> cam = cam_from_f32(A.camf);
	const float inv_w = 1.0f / (float)A.W, inv_h = 1.0f / (float)A.H;
	const HotRange br = A.sc.brute_range;
	const HotIds *ids = WIDE ? A.sc.wide_ids : (BVH ? A.sc.bvh_ids : A.sc.brute_ids);
	const HotPrim *tree_prims = WIDE ? A.sc.wide_prims : A.sc.bvh_prims;
	const unsigned full = 0xffffffffu;
	const unsigned lt_mask = (1u << lane) - 1u;
	const int total = (x0 < A.W && y0 < A.H) ? 32 * A.s_count : 0;
	const bool tile_inside = x0 + 8 <= A.W && y0 + 4 <= A.H;  // warp-uniform: no per-task border test

	int next = 0;          // next unissued task (warp-uniform)
	int task = -1;         // this lane's task, -1 = needs one
	bool exhausted = false;
	bool ray_ok = false;
	int bounce = 0;
	F3 o = mk<float>(0.f, 0.f, 0.f), d = mk<float>(0.f, 0.f, 1.f), thr = o;
	// surface interaction carried from the classify stage to the scatter stage
	F3 sP = o, sN = o;
	float sp0 = 0.f;
	int sdev = 0, sbits = 0;
	int orig = -1;  // hot slot of the primitive the current ray starts on
	Hit h;          // best hit of the ray in flight (BVH: survives across trips while its traversal is suspended)
	h.t = INFINITY; h.idx = -1; h.orig = -1;
	bool trav = false;  // BVH: traversal in progress
	int node = 0, sp = 0;  // BVH2: `node` is the traversal cursor (intersect.cuh: bvh_step); 8-wide: sp indexes wstack
	int stack[(MODE == 1 || MODE == 4) ? ARE_BVH_STACK : (MODE == 3 ? ARE_BVH4_STACK : 1)];
#ifdef ARE_SHORT_STACK
	__shared__ int s_short[(MODE == 1 || MODE == 3) ? ARE_SHORT_STACK : 1][RENDER_THREADS];
	ShortStack<ARE_SHORT_STACK, RENDER_THREADS> stk;
	stk.sm = &s_short[0][threadIdx.x]; stk.deep = stack; stk.sp = 0;
#else
#if ARE_STACK_CACHE
	CachedStack stk;
	stk.reset(stack);
#else
	PtrStack stk;  // BVH2 stack pointer
	stk.top = stack;
#endif
#endif
	uint2 ng = make_uint2(0u, 0u), tg = ng;  // wide-BVH cursor
	uint2 wstack[WIDE ? ARE_WIDE_STACK : 1];
	unsigned int rays = 0;
	TravCounters tc = { 0, 0, 0, 0, 0 };

	while (true) {
		// ---- A. trace + classify ----------------------------------------------------------------------
		// Brute force: every lane with a ray finishes it in this trip.  BVH: traversals run in SLICES — all lanes step
		// through their hierarchy until fewer than TRAV_MIN_LANES are still traversing while finished lanes wait; the
		// unfinished ones keep their cursor (node, stack, best hit) and resume next trip, the finished ones go on to
		// shading and regeneration.  Long rays no longer hold 31 idle lanes hostage.
		bool finished = ray_ok;
		if (WIDE) {
			if (ray_ok && !trav) {  // a fresh ray
				h.t = INFINITY; h.idx = -1; h.orig = orig;
				++rays;
				wide_start(ng, tg, sp);
				trav = true;
			}
			const int n_rays = __popc(__ballot_sync(full, ray_ok));
			if (__any_sync(full, trav)) {
				const WideRay wr = wide_ray(o, d);
				while (true) {
					// node phase: every traversing lane with no leaf primitive waiting decodes one wide node
					if (trav && tg.y == 0u) trav = wide_node_step<COUNT>(A.sc, wr, A.tmin, h.t, ng, tg, sp, wstack, &tc);
					// leaf phase: a wide node step is ~6x a primitive test, so every lane drains the primitives its node
					// produced before the next node phase — all traversing lanes then take part in every node step
					while (__any_sync(full, trav && tg.y != 0u)) {
						if (trav && tg.y != 0u) wide_leaf_step<COUNT>(A.sc, o, d, A.tmin, tg, h, &tc);
					}
					const int n_trav = __popc(__ballot_sync(full, trav));
					if (n_trav == 0 || (n_trav < TRAV_MIN_LANES && n_trav < n_rays)) break;
				}
			}
			finished = ray_ok && !trav;
		} else if (BVH) {
			// single-cursor form (intersect.cuh): `node` is the cursor, `top` the stack pointer, trav <=> cursor != TRAV_DONE
			if (ray_ok && !trav) {  // a fresh ray
				h.t = INFINITY; h.idx = -1; h.orig = orig;
				++rays;
				if (A.sc.n_nodes == 0) {  // zero or one primitive
					if (A.sc.root_leaf_meta != 0) test_leaf<COUNT>(A.sc, A.sc.root_leaf_meta, o, d, A.tmin, h, &tc);
				} else { node = 0; stk.reset(stack); trav = true; }
			}
			const int n_rays = __popc(__ballot_sync(full, ray_ok));
			if (__any_sync(full, trav)) {
				const typename Trav<MODE>::Slopes rs = Trav<MODE>::slopes(A.sc, o, d);
				if (!trav) node = TRAV_DONE;
				while (true) {
#pragma unroll 1
					for (int rep = 0; rep < (MODE == 4 ? TRAV_STEPS_PER_VOTE_Q / TRAV_LEAF_EVERY_Q : TRAV_STEPS_PER_VOTE / TRAV_LEAF_EVERY); ++rep) {  // several steps between the warp votes that decide the end of the slice
						// The warp executes the leaf phase whenever ANY lane holds a leaf — with ~19 lanes traversing that is most
						// iterations, for one or two lanes each time.  Running it once per TRAV_LEAF_EVERY node phases lets lanes that
						// reach a leaf wait a step or two and cuts the leaf code's share of the issue slots accordingly.
#pragma unroll
						for (int r = 0; r < (MODE == 4 ? TRAV_LEAF_EVERY_Q : TRAV_LEAF_EVERY); ++r)
							if (node >= 0) Trav<MODE>::template step<COUNT>(A.sc, A.tmin, rs, node, stk, h, &tc);  // node phase
						if (node < 0 && node != TRAV_DONE) bvh_leaf<COUNT>(A.sc, o, d, A.tmin, node, stk, h, &tc);  // leaf phase
					}
					const int n_trav = __popc(__ballot_sync(full, node != TRAV_DONE));
					if (n_trav == 0 || (n_trav < (MODE == 4 ? TRAV_MIN_LANES_Q : TRAV_MIN_LANES) && n_trav < n_rays)) break;
				}
				trav = node != TRAV_DONE;
			}
			finished = ray_ok && !trav;
		} else if (ray_ok) {
			h.t = INFINITY; h.idx = -1; h.orig = orig;
#ifdef ARE_BAKED
			if (BAKED) intersect_baked(o, d, A.tmin, h);
			else
#endif
			if (LEAN) intersect_lean(sb_prims, br.nb, A.sc.lean_n_open, br.nq, br.nt, o, d, A.tmin, h);
			else intersect_range<lds4>(s_prims, br.first, br.nq, br.nt, br.ns, br.nb, o, d, A.tmin, h);
			++rays;
		}
		if (finished) {
			F3 contrib = mk<float>(0.f, 0.f, 0.f);
			bool done;
			if (h.idx < 0) {
				if (!A.bg_black) contrib = thr * background(A, d);
				done = true;
			} else if (LEAN) {
				sP = o + h.t * d;
				int rec = lds1a(sb_sbase + 4u * h.idx);
				if (h.idx < 2 * br.nb) {
					const uint32_t pb = sb_prims + 48u * h.idx;
					rec += box_hit_face(lds4a(pb), lds4a(pb + 16u), lds4a(pb + 32u), lds4a(pb + 64u), sP);
				}
				const float4 s0 = lds4a(sb_shade + 32u * rec), s1 = lds4a(sb_shade + 32u * rec + 16u);
				sdev = rec;
				sbits = __float_as_int(s0.w);
				sN = mk<float>(s0.x, s0.y, s0.z);
				sp0 = s1.w;
				const F3 scol = mk<float>(s1.x, s1.y, s1.z);
				if ((sbits & 255) == MK_LIGHT) {
					contrib = thr * scol;
					done = true;
				} else {
					done = bounce >= A.max_depth;
					thr = thr * scol;
				}
			} else {
				sP = o + h.t * d;
				orig = h.idx;
				const HotIds id = BVH ? hit_ids<ldg4>(A.sc, tree_prims, ids, h.idx, sP) : hit_ids<lds4>(A.sc, s_prims, ids, h.idx, sP);
				// which half of a fused pair was hit only matters when the halves shade differently (id.b >= 0)
				sdev = id.b >= 0 ? resolve_exact(A.sc, id, sP).dev_prim : id.a;
				const float4 s0 = __ldg(reinterpret_cast<const float4 *>(&A.sc.shade[sdev].r0));
				const float4 s1 = __ldg(reinterpret_cast<const float4 *>(&A.sc.shade[sdev].r1));
				sbits = __float_as_int(s0.w);
				sN = mk<float>(s0.x, s0.y, s0.z);
				F3 scol = mk<float>(s1.x, s1.y, s1.z);
				sp0 = s1.w;
				if (sdev >= A.sc.n_tri + A.sc.n_quad) {  // sphere: outward normal from (centre, radius)
					const float4 c = __ldg(reinterpret_cast<const float4 *>(&A.sc.prim_plane[sdev].r0));
					sN = (1.0f / c.w) * (sP - mk<float>(c.x, c.y, c.z));
				}
				const int kind = sbits & 255;
				if (kind == MK_LIGHT) {
					if (!((sbits >> 8) & 1)) {  // textured light: general path
						const Resolved rs = resolve_exact(A.sc, id, sP);
						const PrimInfo pi = A.sc.info[rs.dev_prim];
						float u, v;
						surface_at(A.sc, rs.dev_prim, sP, rs.a, rs.b, sN, u, v);
						const MaterialRec &m = A.sc.mats[pi.mat];
						scol = m.pf[1] * tex_eval<float>(A.sc, mat_texture(m, pi.tex), u, v, sP);
					}
					contrib = thr * scol;
					done = true;
				} else {
					done = bounce >= A.max_depth;  // truncated path contributes nothing (RTIOW depth cut-off)
					if ((sbits >> 8) & 1) thr = thr * scol;  // SHADE_FAST: the attenuation is the solid colour, apply it now
				}
			}
			if (done) {
				const float csum = contrib.x + contrib.y + contrib.z;  // radiance is non-negative: 0 adds nothing, NaN / inf are dropped
				// (measured: dropping the finiteness half of this test in the lean kernel — its contributions are finite by
				// construction — removes six instructions and LOSES 1.8 %: the compiler schedules the loop differently)
				if (csum > 0.0f && csum < INFINITY) {
					if (LEAN) {
						float *acc = &s_acc[warp][(task & 31) * 3];
						atomicAdd(acc, contrib.x); atomicAdd(acc + 1, contrib.y); atomicAdd(acc + 2, contrib.z);
					} else {
						const uint32_t acc = acc_sb + 12u * (uint32_t)(task & 31);
						asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(acc), "f"(contrib.x) : "memory");
						asm volatile("red.shared.add.f32 [%0 + 4], %1;" ::"r"(acc), "f"(contrib.y) : "memory");
						asm volatile("red.shared.add.f32 [%0 + 8], %1;" ::"r"(acc), "f"(contrib.z) : "memory");
					}
				}
				task = -1;
			}
		}
		// ---- B. lanes without a task take tickets -----------------------------------------------------
		{
			const bool need = task < 0 && !exhausted;
			const unsigned m = __ballot_sync(full, need);
			if (m == 0u && __all_sync(full, exhausted)) break;
			if (need) {
				const int k = next + __popc(m & lt_mask);
				if (k < total) {
					if (tile_inside || (x0 + (k & 7) < A.W && y0 + ((k >> 3) & 3) < A.H)) {  // tiles on the image border own pixels outside it
						task = k;
						bounce = 0;
					}
				} else exhausted = true;
			}
			next += __popc(m);
		}
		// ---- C. one Philox draw per lane: camera ray for a new path, scatter for a continuing one -------
		if (!trav) ray_ok = false;  // a suspended traversal keeps its ray
		// C0. Noise-textured surfaces: the lanes about to scatter on one hand their turbulence sums to the whole warp
		// (shade.cuh: turbulence_coop) instead of each walking seven octaves alone in a nearly empty warp.
		float turb = 0.0f;
		bool have_turb = false;
		if (!LEAN && NOISE && A.sc.has_noise) {
			const bool general = task >= 0 && !trav && bounce > 0 && !((sbits >> 8) & 1);
			if (__any_sync(full, general)) {
				// the shading record says which texture colours the hit (dev_types.h: ShadeRec bits)
				const int mk_ = sbits & 255;
				const int noise_tex = general && ((sbits >> SHADE_TEXKIND_SHIFT) & 7) == TK_NOISE && mk_ != MK_DIELECTRIC && mk_ != MK_LIGHT
					? ((sbits >> SHADE_TEXID_SHIFT) & SHADE_TEXID_MASK) : -1;
				if (__any_sync(full, noise_tex >= 0)) {
					turb = turbulence_coop(A.sc, noise_tex >= 0, noise_tex, sP, turb_q_sb, turb_sum_sb);
					have_turb = noise_tex >= 0;
				}
			}
		}
		if (task >= 0 && !trav) {
			// task -> (pixel of the tile, sample): recomputed here instead of living in four registers across the trip
			const int px = x0 + (task & 7), py = y0 + ((task >> 3) & 3);
			const uint32_t pixel = (uint32_t)(py * A.W + px), sample = (uint32_t)(A.s_begin + (task >> 5));
			Rnd4<float> r = rnd4<float>(A.key, pixel, sample, (uint32_t)bounce, 0u);
			if (bounce == 0) {
				if (!cam.jitter) { r.x = 0.5f; r.y = 0.5f; }
				cam_ray<float>(cam, inv_w, inv_h, px, py, r, o, d);
				orig = -1;
				thr = mk<float>(1.f, 1.f, 1.f);
				bounce = 1;
				ray_ok = true;
			} else {
				F3 wo;
				bool alive;
				if (LEAN) alive = scatter_dir<float>(sbits & 255, sp0, d, sN, r, wo, sb_frame + 32u * sdev);
				else if ((sbits >> 8) & 1) alive = scatter_dir<float>(sbits & 255, sp0, d, sN, r, wo);  // SHADE_FAST: solid colour, simple lobe
				else {  // general path: textures, Reflective's lobe choice
					const Resolved rs = resolve_exact(A.sc, HotIds{ sdev, -1 }, sP);
					const PrimInfo pi = A.sc.info[rs.dev_prim];
					float u, v;
					F3 att, emit;
					const int tk = (sbits >> SHADE_TEXKIND_SHIFT) & 7;
					surface_at(A.sc, rs.dev_prim, sP, rs.a, rs.b, sN, u, v, tk == TK_CHECKER_UV || tk == TK_IMAGE);
					alive = scatter<float>(A.sc, pi.mat, pi.tex, d, sN, sP, u, v, r, wo, att, emit, have_turb ? &turb : nullptr);
					thr = thr * att;
				}
				if (alive) {
					o = sP;
					d = wo;
					++bounce;
					ray_ok = true;
				} else task = -1;  // absorbed (e.g. fuzzed metal reflection below the surface): sample contributes 0
			}
		}
	}
	__syncwarp();
	{
		const int hx = x0 + (lane & 7), hy = y0 + (lane >> 3);
		if (hx < A.W && hy < A.H && A.s_count > 0) {
			float *acc = A.accum + ((size_t)hy * A.W + hx) * 3;
			acc[0] += s_acc[warp][lane * 3]; acc[1] += s_acc[warp][lane * 3 + 1]; acc[2] += s_acc[warp][lane * 3 + 2];
		}
	}
	// counters: one atomic per warp
	unsigned long long r64 = rays;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) r64 += __shfl_down_sync(0xffffffffu, r64, off);
	if (lane == 0 && r64) atomicAdd(A.counters + CNT_RAYS, r64);
	if (BVH && COUNT) {
		unsigned long long c[5] = { tc.nodes, tc.quads, tc.tris, tc.spheres, tc.boxes };
#pragma unroll
		for (int k = 0; k < 5; ++k) {
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) c[k] += __shfl_down_sync(0xffffffffu, c[k], off);
			if (lane == 0 && c[k]) atomicAdd(A.counters + CNT_NODES + k, c[k]);
		}
	}
}

}  // namespace areb
