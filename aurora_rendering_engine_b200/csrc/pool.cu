// pool.cu — the path integrator with an ON-CHIP ray pool per warp (north_star item 2/3: "compacted ray queues", here in
// shared memory instead of HBM).
//
// What costs the BVH megakernel its lanes is the variance of ray lengths inside a warp: 19 of 32 lanes work on average
// (RTIOW), and the wavefront measurement (profiles/r02_wavefront.md) showed that compacting queues BETWEEN kernels does
// not change that (20.7 lanes in a pure traversal kernel) while costing 160 B of HBM traffic per ray.  This kernel
// decouples traversal from shading inside the warp instead:
//
//   a warp owns an 8x4 pixel tile (as in the megakernel) and a pool of POOL_RAYS ray slots in shared memory
//   (48 B of path state + 8 B of hit record each);
//   fill      free slots are filled with camera rays of the tile's next (pixel, sample) tasks;
//   traverse  lanes FETCH rays from the pool by ticket (ballot / popc) until it is drained: a lane whose ray ends early
//             takes the next one at once, so the traversal loop runs with nearly all lanes busy until the pool's tail;
//   shade     all hits are shaded 32 at a time (path_shade.cuh: the megakernel's classify + scatter stages), finished
//             samples go to the tile accumulators, continuing rays are compacted in place.
//
// Same estimator, same Philox counters, same device routines as k_render_path<BVH2>: identical ray counts, images equal
// up to summation order (tests/test_gpu_pool.py).
#include <stdio.h>

#include <algorithm>

#include "kernels.h"
#include "path_shade.cuh"

namespace areb {

#ifndef POOL_RAYS
#define POOL_RAYS 64
#endif
#ifndef POOL_MIN_BLOCKS
#define POOL_MIN_BLOCKS 8
#endif
#ifndef POOL_MIN_BLOCKS_BIG
#define POOL_MIN_BLOCKS_BIG 12
#endif

namespace {

struct PoolSlots {  // per warp, structure of arrays: one LDS.128 / STS.128 per field
	float4 a[POOL_RAYS];   // o.xyz, d.x
	float4 b[POOL_RAYS];   // d.yz, thr.xy
	float4 c[POOL_RAYS];   // thr.z, task (pixel of the tile | sample << 5), bounce, slot the ray starts on
	float2 hit[POOL_RAYS]; // t, hot slot (-1: miss)
};

template <bool COUNT, bool BIG>
__global__ void __launch_bounds__(RENDER_THREADS, BIG ? POOL_MIN_BLOCKS_BIG : POOL_MIN_BLOCKS) k_render_pool(const __grid_constant__ RenderArgs A) {
	__shared__ PoolSlots s_pool[RENDER_THREADS / 32];
	__shared__ float s_acc[RENDER_THREADS / 32][96];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	PoolSlots &P = s_pool[warp];
	for (int i = lane; i < 96; i += 32) s_acc[warp][i] = 0.0f;
	__syncwarp();
	int x0, y0;
	warp_tile_origin(A.W, x0, y0);
	const CamT<float> cam = cam_from_f32(A.camf);
	const float inv_w = 1.0f / (float)A.W, inv_h = 1.0f / (float)A.H;
	const unsigned full = 0xffffffffu, lt_mask = (1u << lane) - 1u;
	const int total = (x0 < A.W && y0 < A.H) ? 32 * A.s_count : 0;
	const bool tile_inside = x0 + 8 <= A.W && y0 + 4 <= A.H;
	int next_task = 0, n_pool = 0;  // warp-uniform
	unsigned int rays = 0;
	TravCounters tc = { 0, 0, 0, 0, 0 };
	int stack[ARE_BVH_STACK];

	while (true) {
		// ---- fill: camera rays of the next tasks into the free slots -----------------------------------------
		while (n_pool < POOL_RAYS && next_task < total) {
			const int k = next_task + lane;
			const bool take = lane < POOL_RAYS - n_pool && k < total;
			const bool inside = take && (tile_inside || (x0 + (k & 7) < A.W && y0 + ((k >> 3) & 3) < A.H));
			const unsigned m = __ballot_sync(full, inside);
			if (inside) {
				const int px = x0 + (k & 7), py = y0 + ((k >> 3) & 3);
				const uint32_t pixel = (uint32_t)(py * A.W + px), sample = (uint32_t)(A.s_begin + (k >> 5));
				Rnd4<float> rn = rnd4<float>(A.key, pixel, sample, 0u, 0u);
				if (!cam.jitter) { rn.x = 0.5f; rn.y = 0.5f; }
				F3 o, d;
				cam_ray<float>(cam, inv_w, inv_h, px, py, rn, o, d);
				const int s = n_pool + __popc(m & lt_mask);
				P.a[s] = make_float4(o.x, o.y, o.z, d.x);
				P.b[s] = make_float4(d.y, d.z, 1.0f, 1.0f);
				P.c[s] = make_float4(1.0f, __int_as_float(k), __int_as_float(1), __int_as_float(-1));
			}
			next_task += min(32, min(POOL_RAYS - n_pool, total - next_task));
			n_pool += __popc(m);
		}
		if (n_pool == 0) break;
		__syncwarp();
		// ---- traverse: lanes fetch rays by ticket until the pool is drained -----------------------------------
		{
			int fetched = 0;         // warp-uniform: rays handed out so far
			int my = -1;             // slot of the ray this lane is traversing
			int node = TRAV_DONE;
			F3 o = mk<float>(0.f, 0.f, 0.f), d = mk<float>(0.f, 0.f, 1.f);
			RaySlopes rs = ray_slopes(o, d);
			Hit h;
			h.t = INFINITY; h.idx = -1; h.orig = -1;
			PtrStack stk;
			stk.top = stack;
			while (true) {
				const bool need = node == TRAV_DONE;
				if (need && my >= 0) {  // the ray just finished: record its hit
					P.hit[my] = make_float2(h.t, __int_as_float(h.idx));
					my = -1;
				}
				const unsigned m = __ballot_sync(full, need);
				if (fetched >= n_pool && m == full) break;  // nothing left to hand out and nobody traversing
				if (need) {
					const int k = fetched + __popc(m & lt_mask);
					if (k < n_pool) {
						my = k;
						const float4 a = P.a[k], b = P.b[k], c = P.c[k];
						o = mk<float>(a.x, a.y, a.z); d = mk<float>(a.w, b.x, b.y);
						h.t = INFINITY; h.idx = -1; h.orig = __float_as_int(c.w);
						++rays;
						if (A.sc.n_nodes == 0) {  // zero or one primitive: no hierarchy
							if (A.sc.root_leaf_meta != 0) test_leaf<COUNT>(A.sc, A.sc.root_leaf_meta, o, d, A.tmin, h, &tc);
							P.hit[my] = make_float2(h.t, __int_as_float(h.idx));
							my = -1;
						} else {
							rs = ray_slopes(o, d);
							node = 0;
							stk.reset(stack);
						}
					}
				}
				fetched = min(n_pool, fetched + __popc(m));
#pragma unroll 1
				for (int rep = 0; rep < TRAV_STEPS_PER_VOTE; ++rep) {
					if (node >= 0) bvh_step<COUNT>(A.sc, A.tmin, rs, node, stk, h, &tc);
					if (node < 0 && node != TRAV_DONE) bvh_leaf<COUNT>(A.sc, o, d, A.tmin, node, stk, h, &tc);
				}
			}
		}
		__syncwarp();
		// ---- shade: 32 hits at a time; continuing rays are compacted in place ---------------------------------
		int n_new = 0;
		for (int base = 0; base < n_pool; base += 32) {
			const int s = base + lane;
			const bool valid = s < n_pool;
			bool alive = false;
			PathRay pr;
			int task = 0;
			if (valid) {
				const float4 a = P.a[s], b = P.b[s], c = P.c[s];
				const float2 hr = P.hit[s];
				pr.o = mk<float>(a.x, a.y, a.z); pr.d = mk<float>(a.w, b.x, b.y); pr.thr = mk<float>(b.z, b.w, c.x);
				task = __float_as_int(c.y); pr.bounce = __float_as_int(c.z); pr.orig = __float_as_int(c.w);
				const int px = x0 + (task & 7), py = y0 + ((task >> 3) & 3);
				F3 contrib;
				bool done;
				shade_one(A, pr, hr.x, __float_as_int(hr.y), (uint32_t)(py * A.W + px), (uint32_t)(A.s_begin + (task >> 5)), done, alive, contrib);
				if (done) {
					const float csum = contrib.x + contrib.y + contrib.z;
					if (csum > 0.0f && csum < INFINITY) {
						float *acc = &s_acc[warp][(task & 31) * 3];
						atomicAdd(acc, contrib.x); atomicAdd(acc + 1, contrib.y); atomicAdd(acc + 2, contrib.z);
					}
				}
			}
			__syncwarp();  // every lane has read its slot before any slot is overwritten
			const unsigned m = __ballot_sync(full, alive);
			if (alive) {
				const int t = n_new + __popc(m & lt_mask);  // <= s: never a slot that is still to be read
				P.a[t] = make_float4(pr.o.x, pr.o.y, pr.o.z, pr.d.x);
				P.b[t] = make_float4(pr.d.y, pr.d.z, pr.thr.x, pr.thr.y);
				P.c[t] = make_float4(pr.thr.z, __int_as_float(task), __int_as_float(pr.bounce), __int_as_float(pr.orig));
			}
			n_new += __popc(m);
			__syncwarp();
		}
		n_pool = n_new;
	}
	__syncwarp();
	{
		const int hx = x0 + (lane & 7), hy = y0 + (lane >> 3);
		if (hx < A.W && hy < A.H && A.s_count > 0) {
			float *acc = A.accum + ((size_t)hy * A.W + hx) * 3;
			acc[0] += s_acc[warp][lane * 3]; acc[1] += s_acc[warp][lane * 3 + 1]; acc[2] += s_acc[warp][lane * 3 + 2];
		}
	}
	unsigned long long r64 = rays;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) r64 += __shfl_down_sync(0xffffffffu, r64, off);
	if (lane == 0 && r64) atomicAdd(A.counters + CNT_RAYS, r64);
	if (COUNT) {
		unsigned long long cn[5] = { tc.nodes, tc.quads, tc.tris, tc.spheres, tc.boxes };
#pragma unroll
		for (int k = 0; k < 5; ++k) {
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) cn[k] += __shfl_down_sync(0xffffffffu, cn[k], off);
			if (lane == 0 && cn[k]) atomicAdd(A.counters + CNT_NODES + k, cn[k]);
		}
	}
}

}  // namespace

int launch_render_pool(const RenderArgs &a, bool count_tests, cudaStream_t s) {
	const int warps = ((a.W + 15) / 16) * ((a.H + 7) / 8) * 4;
	const int tiles = (warps + RENDER_THREADS / 32 - 1) / (RENDER_THREADS / 32);
	if (tiles <= 0) return -1;
	if (!a.sc.nodes && a.sc.n_nodes > 0) return -1;
	const bool big = render_path_is_big(a);
	if (count_tests) {
		if (big) k_render_pool<true, true><<<tiles, RENDER_THREADS, 0, s>>>(a);
		else k_render_pool<true, false><<<tiles, RENDER_THREADS, 0, s>>>(a);
	} else {
		if (big) k_render_pool<false, true><<<tiles, RENDER_THREADS, 0, s>>>(a);
		else k_render_pool<false, false><<<tiles, RENDER_THREADS, 0, s>>>(a);
	}
	return 1;
}

}  // namespace areb
