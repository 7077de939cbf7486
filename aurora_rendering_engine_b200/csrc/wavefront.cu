// wavefront.cu — the wavefront formulation of the path integrator (north_star item 3, SURVEY.md §2.2 k_wf_*), built to
// MEASURE the megakernel-versus-wavefront choice on both sides (profiles/r02_wavefront.md).
//
// Same estimator, same Philox counters (pixel, global sample, bounce), same device routines (cam_ray, the BVH2 traversal,
// the shading records, scatter_dir / scatter) as k_render_path — only the schedule differs:
//
//   k_wf_generate   one thread per (pixel, sample) of the batch: camera ray into the ray queue
//   per bounce:
//   k_wf_extend     one thread per queued ray: closest hit through the BVH2 -> (t, hot slot) per ray.  Traversal only:
//                   every lane of a warp is traversing, none is shading or idle
//   k_wf_shade      one thread per queued ray: miss / light -> contribution added to the accumulator; otherwise the
//                   material scatters and the continuing ray is appended to the NEXT queue — warp-aggregated
//                   (__ballot_sync + one atomicAdd per warp), so the next extend launch runs on a compacted queue
//
// A queued ray is 48 bytes (origin, direction, throughput, pixel, sample | bounce, slot it starts on) + 8 bytes of hit
// record: every path segment costs 48 + 8 written and 48 + 8 read by the next stage, twice (extend reads the ray, shade
// reads ray + hit and writes the next ray): >= 160 bytes of HBM traffic per ray where the megakernel moves none.
#include <stdio.h>

#include <algorithm>

#include "kernels.h"
#include "path_shade.cuh"

namespace areb {

namespace {

struct WfRay {  // three float4
	float4 a;   // o.xyz, d.x
	float4 b;   // d.yz, thr.xy
	float4 c;   // thr.z, pixel, sample << 8 | bounce, slot the ray starts on (-1: none)
};

__device__ __forceinline__ void wf_store(float4 *q, size_t cap, size_t i, F3 o, F3 d, F3 thr, int pixel, int sample, int bounce, int orig) {
	q[i] = make_float4(o.x, o.y, o.z, d.x);
	q[cap + i] = make_float4(d.y, d.z, thr.x, thr.y);
	q[2 * cap + i] = make_float4(thr.z, __int_as_float(pixel), __int_as_float(sample << 8 | bounce), __int_as_float(orig));
}

__global__ void __launch_bounds__(256) k_wf_generate(const __grid_constant__ RenderArgs A, int s_begin, int s_count, float4 *queue, size_t cap, unsigned *n_out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	// a warp covers an 8x4 pixel tile of one sample, as in the megakernel: coherent primary rays
	const int tiles_x = (A.W + 7) >> 3, tiles_y = (A.H + 3) >> 2;
	const size_t per_sample = (size_t)tiles_x * tiles_y * 32;
	if (i >= per_sample * (size_t)s_count) return;
	const int s = (int)(i / per_sample);
	const size_t r = i - (size_t)s * per_sample;
	const int tile = (int)(r >> 5), lane = (int)(r & 31);
	const int px = (tile % tiles_x) * 8 + (lane & 7), py = (tile / tiles_x) * 4 + (lane >> 3);
	const bool inside = px < A.W && py < A.H;
	const unsigned m = __ballot_sync(0xffffffffu, inside);
	unsigned base = 0;
	if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(n_out, (unsigned)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (!inside) return;
	const CamT<float> cam = cam_from_f32(A.camf);
	const uint32_t pixel = (uint32_t)(py * A.W + px), sample = (uint32_t)(s_begin + s);
	Rnd4<float> rn = rnd4<float>(A.key, pixel, sample, 0u, 0u);
	if (!cam.jitter) { rn.x = 0.5f; rn.y = 0.5f; }
	F3 o, d;
	cam_ray<float>(cam, 1.0f / (float)A.W, 1.0f / (float)A.H, px, py, rn, o, d);
	wf_store(queue, cap, base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u)), o, d, mk<float>(1.f, 1.f, 1.f), (int)pixel, (int)sample, 1, -1);
}

template <bool COUNT>
__global__ void __launch_bounds__(256) k_wf_extend(const __grid_constant__ RenderArgs A, const float4 *__restrict__ queue, size_t cap, const unsigned *__restrict__ n_in,
	float2 *__restrict__ hits) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned n = *n_in;
	TravCounters tc = { 0, 0, 0, 0, 0 };
	if (i < n) {
		const float4 a = queue[i], b = queue[cap + i], c = queue[2 * cap + i];
		Hit h;
		h.t = INFINITY; h.idx = -1; h.orig = __float_as_int(c.w);
		intersect_bvh<COUNT>(A.sc, mk<float>(a.x, a.y, a.z), mk<float>(a.w, b.x, b.y), A.tmin, h, &tc);
		hits[i] = make_float2(h.t, __int_as_float(h.idx));
	}
	if (i == 0) atomicAdd(A.counters + CNT_RAYS, (unsigned long long)n);
	if (COUNT) {
		unsigned long long cn[5] = { tc.nodes, tc.quads, tc.tris, tc.spheres, tc.boxes };
#pragma unroll
		for (int k = 0; k < 5; ++k) {
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) cn[k] += __shfl_down_sync(0xffffffffu, cn[k], off);
			if ((threadIdx.x & 31) == 0 && cn[k]) atomicAdd(A.counters + CNT_NODES + k, cn[k]);
		}
	}
}

__global__ void __launch_bounds__(256) k_wf_shade(const __grid_constant__ RenderArgs A, const float4 *__restrict__ queue, size_t cap, const unsigned *__restrict__ n_in,
	const float2 *__restrict__ hits, float4 *__restrict__ next, unsigned *n_out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned n = *n_in;
	bool alive = false;
	F3 o = mk<float>(0.f, 0.f, 0.f), d = o, thr = o;
	int pixel = 0, sample = 0, bounce = 0, orig = -1;
	if (i < n) {
		const float4 a = queue[i], b = queue[cap + i], c = queue[2 * cap + i];
		const float2 hr = hits[i];
		o = mk<float>(a.x, a.y, a.z); d = mk<float>(a.w, b.x, b.y); thr = mk<float>(b.z, b.w, c.x);
		pixel = __float_as_int(c.y);
		sample = __float_as_int(c.z) >> 8; bounce = __float_as_int(c.z) & 255;
		PathRay pr;
		pr.o = o; pr.d = d; pr.thr = thr; pr.bounce = bounce; pr.orig = orig;
		F3 contrib;
		bool done;
		shade_one(A, pr, hr.x, __float_as_int(hr.y), (uint32_t)pixel, (uint32_t)sample, done, alive, contrib);
		o = pr.o; d = pr.d; thr = pr.thr; bounce = pr.bounce; orig = pr.orig;
		if (done) {
			const float csum = contrib.x + contrib.y + contrib.z;
			if (csum > 0.0f && csum < INFINITY) {
				float *acc = A.accum + (size_t)pixel * 3;
				atomicAdd(acc, contrib.x); atomicAdd(acc + 1, contrib.y); atomicAdd(acc + 2, contrib.z);
			}
		}
	}
	// compaction: the continuing rays of a warp take consecutive slots of the next queue
	const unsigned m = __ballot_sync(0xffffffffu, alive);
	unsigned base = 0;
	if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(n_out, (unsigned)__popc(m));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (alive) wf_store(next, cap, base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u)), o, d, thr, pixel, sample, bounce, orig);
}

}  // namespace

size_t wavefront_workspace_bytes(int W, int H, int batch_spp) {
	const size_t cap = (size_t)W * H * batch_spp;
	return 2 * 3 * cap * sizeof(float4) + cap * sizeof(float2) + 256;
}

// Renders samples [s_begin, s_begin + s_count) in batches of batch_spp.  workspace: wavefront_workspace_bytes(W, H, batch_spp).
// Returns kernels launched (< 0: error).
int launch_render_wavefront(const RenderArgs &a, bool count_tests, int batch_spp, void *workspace, cudaStream_t s) {
	if (!a.sc.nodes && a.sc.n_nodes > 0) return -1;
	const size_t cap = (size_t)a.W * a.H * batch_spp;
	float4 *q[2] = { static_cast<float4 *>(workspace), static_cast<float4 *>(workspace) + 3 * cap };
	float2 *hits = reinterpret_cast<float2 *>(static_cast<float4 *>(workspace) + 6 * cap);
	unsigned *cnt = reinterpret_cast<unsigned *>(hits + cap);  // two queue counters in the 256 spare bytes
	int launches = 0;
	const int tiles_x = (a.W + 7) >> 3, tiles_y = (a.H + 3) >> 2;
	for (int b0 = 0; b0 < a.s_count; b0 += batch_spp) {
		const int bs = std::min(batch_spp, a.s_count - b0);
		cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned), s);
		const size_t gen_threads = (size_t)tiles_x * tiles_y * 32 * bs;
		k_wf_generate<<<(unsigned)((gen_threads + 255) / 256), 256, 0, s>>>(a, a.s_begin + b0, bs, q[0], cap, cnt);
		++launches;
		const unsigned grid = (unsigned)(((size_t)a.W * a.H * bs + 255) / 256);
		for (int bounce = 1; bounce <= a.max_depth; ++bounce) {
			const int in = (bounce - 1) & 1, out = bounce & 1;
			cudaMemsetAsync(cnt + out, 0, sizeof(unsigned), s);
			if (count_tests) k_wf_extend<true><<<grid, 256, 0, s>>>(a, q[in], cap, cnt + in, hits);
			else k_wf_extend<false><<<grid, 256, 0, s>>>(a, q[in], cap, cnt + in, hits);
			k_wf_shade<<<grid, 256, 0, s>>>(a, q[in], cap, cnt + in, hits, q[out], cnt + out);
			launches += 2;
		}
	}
	return launches;
}

}  // namespace areb
