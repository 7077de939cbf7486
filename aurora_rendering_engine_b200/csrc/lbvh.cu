// lbvh.cu — BVH2 construction on the device: Morton codes -> radix sort -> Karras' parallel
// radix tree -> bottom-up boxes, written straight into the traversal layout of dev_types.h (64-byte nodes holding both
// children's boxes as (centre, half-extent), one hot primitive per leaf, leaf reference = ~(slot | kind << 29)).
//
// The reference has no hierarchy at all (its ObjectSet is a std::vector<Triangle*> scanned linearly,
// /root/reference/include/object/object_set.h:10-12, src/object/object.cpp:7-17); the host compiler of scene.cpp
// builds a binned-SAH tree with all host threads (0.75 s for 1 M primitives).  This builder trades a few per cent of
// tree quality for a build that takes milliseconds, for scenes that change every frame.
//
//   k_lbvh_codes     one thread per hot item: 30-bit Morton code of the box centre inside the centroid bounds
//   cub radix sort   (code, item) pairs — stable, so equal codes keep item order and the tree is deterministic
//   k_lbvh_slots     leaf-ordered copy of the primitive records (a box takes two slots), leaf references, leaf boxes
//   k_lbvh_tree      T. Karras, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees" (HPG 2012):
//                    one thread per inner node finds its key range and split; equal codes are split on the index bits
//   k_lbvh_boxes     one thread per leaf walks up; the second thread to reach a node owns it (atomic counter), writes
//                    both child boxes into the node, unites them and continues; tree height on the way
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "dev_types.h"
#include "kernels.h"

namespace areb {

namespace {

__device__ __forceinline__ unsigned expand10(unsigned v) {  // 10 bits -> every third bit
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

__global__ void k_lbvh_codes(int n, const f4 *__restrict__ lo, const f4 *__restrict__ hi, float3 cmin, float3 scale,
	unsigned *__restrict__ code, int *__restrict__ item, int *__restrict__ size) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const f4 a = lo[i], b = hi[i];
	const float cx = 0.5f * (a.x + b.x), cy = 0.5f * (a.y + b.y), cz = 0.5f * (a.z + b.z);
	const unsigned qx = (unsigned)fminf(fmaxf((cx - cmin.x) * scale.x, 0.0f), 1023.0f);
	const unsigned qy = (unsigned)fminf(fmaxf((cy - cmin.y) * scale.y, 0.0f), 1023.0f);
	const unsigned qz = (unsigned)fminf(fmaxf((cz - cmin.z) * scale.z, 0.0f), 1023.0f);
	code[i] = (expand10(qx) << 2) | (expand10(qy) << 1) | expand10(qz);
	item[i] = i;
	size[i] = 0;  // filled in sorted order by k_lbvh_sizes
}

__global__ void k_lbvh_sizes(int n, const int *__restrict__ item, const f4 *__restrict__ lo, int *__restrict__ size) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	size[j] = __float_as_int(lo[item[j]].w) == 0 ? 2 : 1;  // kind 0 = box: two slots
}

// leaf j (sorted order) = item[j]: records to slot[j], leaf reference, leaf box
__global__ void k_lbvh_slots(int n, const int *__restrict__ item, const int *__restrict__ slot, const f4 *__restrict__ lo, const f4 *__restrict__ hi,
	const int *__restrict__ item_slot, const HotPrim *__restrict__ item_prims, const HotIds *__restrict__ item_ids,
	HotPrim *__restrict__ prims, HotIds *__restrict__ ids, int *__restrict__ leaf_ref, f4 *__restrict__ leaf_lo, f4 *__restrict__ leaf_hi) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const int it = item[j], s = slot[j], src = item_slot[it];
	const f4 a = lo[it], b = hi[it];
	const int kind = __float_as_int(a.w);
	prims[s] = item_prims[src];
	ids[s] = item_ids[src];
	if (kind == 0) { prims[s + 1] = item_prims[src + 1]; ids[s + 1] = item_ids[src + 1]; }
	leaf_ref[j] = ~(s | (kind << 29));
	leaf_lo[j] = a;
	leaf_hi[j] = b;
}

// length of the common prefix of keys i and j, equal codes continued on the index bits; -1 outside the array
__device__ __forceinline__ int prefix(const unsigned *__restrict__ code, int n, int i, int j) {
	if (j < 0 || j >= n) return -1;
	const unsigned x = code[i] ^ code[j];
	return x ? __clz(x) : 32 + __clz((unsigned)(i ^ j));
}

// children are recorded as: >= 0 inner node, < 0 leaf ~j (sorted leaf index)
__global__ void k_lbvh_tree(int n, const unsigned *__restrict__ code, int2 *__restrict__ child, int *__restrict__ parent_inner, int *__restrict__ parent_leaf) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	const int d = prefix(code, n, i, i + 1) - prefix(code, n, i, i - 1) >= 0 ? 1 : -1;
	const int dmin = prefix(code, n, i, i - d);
	int lmax = 2;
	while (prefix(code, n, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int t = lmax >> 1; t >= 1; t >>= 1)
		if (prefix(code, n, i, i + (l + t) * d) > dmin) l += t;
	const int j = i + l * d;
	const int dnode = prefix(code, n, i, j);
	int s = 0, t = l;
	do {
		t = (t + 1) >> 1;
		if (prefix(code, n, i, i + (s + t) * d) > dnode) s += t;
	} while (t > 1);
	const int g = i + s * d + min(d, 0);
	const int first = min(i, j), last = max(i, j);
	const int left = first == g ? ~g : g, right = last == g + 1 ? ~(g + 1) : g + 1;
	child[i] = make_int2(left, right);
	if (left < 0) parent_leaf[g] = i; else parent_inner[g] = i;
	if (right < 0) parent_leaf[g + 1] = i; else parent_inner[g + 1] = i;
	if (i == 0) parent_inner[0] = -1;
}

// (lo, hi) -> (centre, half-extent), padded outwards as the host builder pads (scene.cpp put_box): fp32 rounding of the
// bounds and of the slab arithmetic must never cut a primitive off
__device__ __forceinline__ void centre_half(float lo, float hi, float &c, float &h) {
	const float m = fmaxf(fabsf(lo), fabsf(hi));
	const float pad = __fmaf_ru(1e-5f, __fadd_ru(m, __fsub_ru(hi, lo)), 1e-7f);
	const float a = __fsub_rd(lo, pad), b = __fadd_ru(hi, pad);
	c = 0.5f * (a + b);
	h = fmaxf(__fsub_ru(b, c), __fsub_ru(c, a));
	h = __fmul_ru(h, 1.0000002f);
}

__global__ void k_lbvh_boxes(int n, const int2 *__restrict__ child, const int *__restrict__ parent_inner, const int *__restrict__ parent_leaf,
	const int *__restrict__ leaf_ref, const f4 *__restrict__ leaf_lo, const f4 *__restrict__ leaf_hi,
	f4 *node_lo, f4 *node_hi, int *node_height, int *arrivals, BvhNode *__restrict__ nodes) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	int p = parent_leaf[j];
	while (p >= 0) {
		__threadfence();  // my subtree's box / height are visible before I announce myself
		if (atomicAdd(&arrivals[p], 1) == 0) return;  // first to arrive: the sibling subtree's thread takes over
		__threadfence();
		const int2 ch = child[p];
		f4 lo[2], hi[2];
		int hgt[2], ref[2];
		const int c2[2] = { ch.x, ch.y };
#pragma unroll
		for (int k = 0; k < 2; ++k) {
			const int c = c2[k];
			if (c < 0) { lo[k] = leaf_lo[~c]; hi[k] = leaf_hi[~c]; hgt[k] = 0; ref[k] = leaf_ref[~c]; }
			else {  // written by another thread of this launch: read past L1
				const float4 a = __ldcg(reinterpret_cast<const float4 *>(node_lo) + c), b = __ldcg(reinterpret_cast<const float4 *>(node_hi) + c);
				lo[k] = { a.x, a.y, a.z, 0.f };
				hi[k] = { b.x, b.y, b.z, 0.f };
				hgt[k] = __ldcg(node_height + c);
				ref[k] = c;
			}
		}
		BvhNode nd;
		centre_half(lo[0].x, hi[0].x, nd.b0.x, nd.b0.y);
		centre_half(lo[0].y, hi[0].y, nd.b0.z, nd.b0.w);
		centre_half(lo[1].x, hi[1].x, nd.b1.x, nd.b1.y);
		centre_half(lo[1].y, hi[1].y, nd.b1.z, nd.b1.w);
		centre_half(lo[0].z, hi[0].z, nd.b2.x, nd.b2.y);
		centre_half(lo[1].z, hi[1].z, nd.b2.z, nd.b2.w);
		nd.child[0] = ref[0]; nd.child[1] = ref[1];
		nd.meta[0] = 0; nd.meta[1] = 0;
		nodes[p] = nd;
		node_lo[p] = { fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.f };
		node_hi[p] = { fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.f };
		node_height[p] = 1 + max(hgt[0], hgt[1]);
		p = parent_inner[p];
	}
}

// refit: leaf of every item (inverse of the sort permutation), then the moved items' records and boxes into their leaves
__global__ void k_lbvh_leaf_of_item(int n, const int *__restrict__ item_sorted, int *__restrict__ leaf_of_item) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n) leaf_of_item[item_sorted[j]] = j;
}
__global__ void k_lbvh_refit_leaves(int m, const int *__restrict__ item, const HotPrim *__restrict__ rec, const f4 *__restrict__ lo, const f4 *__restrict__ hi,
	const int *__restrict__ leaf_of_item, const int *__restrict__ slot, HotPrim *__restrict__ prims, f4 *__restrict__ leaf_lo, f4 *__restrict__ leaf_hi) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	const int j = leaf_of_item[item[k]];
	prims[slot[j]] = rec[k];
	leaf_lo[j] = lo[k];
	leaf_hi[j] = hi[k];
}

__global__ void k_prim_scatter(const PrimScatterArgs a) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= a.m) return;
	const int dp = a.dp[k];
	const HotPrim r = a.rec[k];
	a.prim_plane[dp] = r;
	const double *g = a.geo64 + 9 * (size_t)k;
	if (dp < a.n_tri + a.n_quad) {
		ShadeRec sr = a.shade[dp];
		sr.r0.x = r.r0.x; sr.r0.y = r.r0.y; sr.r0.z = r.r0.z;  // the geometric normal; material bits and colour stay
		a.shade[dp] = sr;
		double *dst = dp < a.n_tri ? a.tri64 + 9 * (size_t)dp : a.quad64 + 9 * (size_t)(dp - a.n_tri);
		for (int i = 0; i < 9; ++i) dst[i] = g[i];
		if (dp < a.n_tri)
			for (int i = 0; i < 3; ++i) a.rt_tris[3 * (size_t)dp + i] = a.rt[3 * (size_t)k + i];
	} else {
		double *dst = a.sph64 + 4 * (size_t)(dp - a.n_tri - a.n_quad);
		for (int i = 0; i < 4; ++i) dst[i] = g[i];
	}
}

// carves 256-byte aligned arrays out of the caller's workspace
struct Carver {
	unsigned char *base;
	size_t off = 0;
	template <typename T>
	T *get(size_t n) {
		T *p = reinterpret_cast<T *>(base + off);
		off += (n * sizeof(T) + 255) & ~(size_t)255;
		return p;
	}
};

cudaError_t cub_temp_bytes(int n, size_t &bytes) {
	size_t sort_bytes = 0, scan_bytes = 0;
	cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned *)nullptr, (unsigned *)nullptr, (const int *)nullptr, (int *)nullptr, n, 0, 30, nullptr);
	if (e != cudaSuccess) return e;
	e = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int *)nullptr, (int *)nullptr, n, nullptr);
	bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
	return e;
}

}  // namespace

size_t lbvh_workspace_bytes(int n_items) {
	const size_t n = n_items > 0 ? (size_t)n_items : 1;
	size_t cub_bytes = 0;
	if (cub_temp_bytes((int)n, cub_bytes) != cudaSuccess) return 0;
	auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
	// 2 code arrays, 9 int arrays, 1 int2 array, 4 f4 arrays, cub's temporary storage
	return 2 * al(n * 4) + 9 * al(n * 4) + al(n * 8) + 4 * al(n * 16) + al(cub_bytes) + 256;
}

#define LB(call)                                                          \
	do {                                                                  \
		cudaError_t e_ = (call);                                          \
		if (e_ != cudaSuccess) { err = cudaGetErrorString(e_); return -1; } \
	} while (0)

// The workspace layout, shared by build and refit (the refit reads what the build left behind).
struct LbvhArrays {
	unsigned *code, *code_sorted;
	int *item, *item_sorted, *size, *slot, *leaf_ref, *parent_inner, *parent_leaf, *node_height, *arrivals;
	int2 *child;
	f4 *leaf_lo, *leaf_hi, *node_lo, *node_hi;
	explicit LbvhArrays(void *workspace, int n) {
		Carver sc{ static_cast<unsigned char *>(workspace) };
		code = sc.get<unsigned>(n); code_sorted = sc.get<unsigned>(n);
		item = sc.get<int>(n); item_sorted = sc.get<int>(n); size = sc.get<int>(n); slot = sc.get<int>(n); leaf_ref = sc.get<int>(n);
		parent_inner = sc.get<int>(n); parent_leaf = sc.get<int>(n); node_height = sc.get<int>(n); arrivals = sc.get<int>(n);
		child = sc.get<int2>(n);
		leaf_lo = sc.get<f4>(n); leaf_hi = sc.get<f4>(n); node_lo = sc.get<f4>(n); node_hi = sc.get<f4>(n);
	}
};

void launch_prim_scatter(const PrimScatterArgs &a, cudaStream_t s) {
	if (a.m > 0) k_prim_scatter<<<(a.m + 255) / 256, 256, 0, s>>>(a);
}

int lbvh_refit(void *workspace, size_t workspace_bytes, int n_items, int m, const int *item, const HotPrim *rec, const f4 *lo, const f4 *hi,
	HotPrim *prims, BvhNode *nodes, cudaStream_t s, std::string &err) {
	const int n = n_items;
	if (n <= 1 || m <= 0) return 0;
	if (!workspace || workspace_bytes < lbvh_workspace_bytes(n)) { err = "device BVH refit: workspace too small"; return -1; }
	LbvhArrays w(workspace, n);
	const int B = 256;
	// `item` (unsorted ids, dead after the sort) is reused as the leaf-of-item map; `code` is free as well
	int *leaf_of_item = w.item;
	k_lbvh_leaf_of_item<<<(n + B - 1) / B, B, 0, s>>>(n, w.item_sorted, leaf_of_item);
	LB(cudaGetLastError());
	k_lbvh_refit_leaves<<<(m + B - 1) / B, B, 0, s>>>(m, item, rec, lo, hi, leaf_of_item, w.slot, prims, w.leaf_lo, w.leaf_hi);
	LB(cudaGetLastError());
	LB(cudaMemsetAsync(w.arrivals, 0, (size_t)n * sizeof(int), s));
	k_lbvh_boxes<<<(n + B - 1) / B, B, 0, s>>>(n, w.child, w.parent_inner, w.parent_leaf, w.leaf_ref, w.leaf_lo, w.leaf_hi, w.node_lo, w.node_hi, w.node_height, w.arrivals, nodes);
	LB(cudaGetLastError());
	return 3;
}

int lbvh_build(const LbvhInput &in, LbvhOutput &out, void *workspace, size_t workspace_bytes, cudaStream_t s, std::string &err) {
	const int n = in.n_items;
	out.height = 0;
	out.root_leaf_ref = 0;
	out.n_nodes = 0;
	if (n <= 0) return 0;
	if (!workspace || workspace_bytes < lbvh_workspace_bytes(n)) { err = "device BVH build: workspace too small"; return -1; }
	LbvhArrays w(workspace, n);
	unsigned *code = w.code, *code_sorted = w.code_sorted;
	int *item = w.item, *item_sorted = w.item_sorted, *size = w.size, *slot = w.slot, *leaf_ref = w.leaf_ref;
	int *parent_inner = w.parent_inner, *parent_leaf = w.parent_leaf, *node_height = w.node_height, *arrivals = w.arrivals;
	int2 *child = w.child;
	f4 *leaf_lo = w.leaf_lo, *leaf_hi = w.leaf_hi, *node_lo = w.node_lo, *node_hi = w.node_hi;
	size_t cub_bytes = 0;
	LB(cub_temp_bytes(n, cub_bytes));
	// cub's temporary storage sits behind the arrays
	unsigned char *tmp = reinterpret_cast<unsigned char *>(w.node_hi) + (((size_t)n * sizeof(f4) + 255) & ~(size_t)255);
	const int B = 256, G = (n + B - 1) / B;
	float3 cmin = make_float3(in.cmin[0], in.cmin[1], in.cmin[2]), scale;
	const float ext[3] = { in.cmax[0] - in.cmin[0], in.cmax[1] - in.cmin[1], in.cmax[2] - in.cmin[2] };
	scale.x = ext[0] > 0.f ? 1024.0f / ext[0] : 0.f; scale.y = ext[1] > 0.f ? 1024.0f / ext[1] : 0.f; scale.z = ext[2] > 0.f ? 1024.0f / ext[2] : 0.f;
	k_lbvh_codes<<<G, B, 0, s>>>(n, in.item_lo, in.item_hi, cmin, scale, code, item, size);
	LB(cudaGetLastError());
	{
		size_t tb = cub_bytes;
		LB(cub::DeviceRadixSort::SortPairs(tmp, tb, code, code_sorted, item, item_sorted, n, 0, 30, s));
		k_lbvh_sizes<<<G, B, 0, s>>>(n, item_sorted, in.item_lo, size);
		LB(cudaGetLastError());
		size_t sb = cub_bytes;
		LB(cub::DeviceScan::ExclusiveSum(tmp, sb, size, slot, n, s));
	}
	k_lbvh_slots<<<G, B, 0, s>>>(n, item_sorted, slot, in.item_lo, in.item_hi, in.item_slot, in.item_prims, in.item_ids, out.prims, out.ids, leaf_ref, leaf_lo, leaf_hi);
	LB(cudaGetLastError());
	if (n == 1) {  // the root IS the leaf
		LB(cudaMemcpyAsync(&out.root_leaf_ref, leaf_ref, sizeof(int), cudaMemcpyDeviceToHost, s));
		LB(cudaStreamSynchronize(s));
		return 0;
	}
	LB(cudaMemsetAsync(arrivals, 0, (size_t)n * sizeof(int), s));
	k_lbvh_tree<<<G, B, 0, s>>>(n, code_sorted, child, parent_inner, parent_leaf);
	LB(cudaGetLastError());
	k_lbvh_boxes<<<G, B, 0, s>>>(n, child, parent_inner, parent_leaf, leaf_ref, leaf_lo, leaf_hi, node_lo, node_hi, node_height, arrivals, out.nodes);
	LB(cudaGetLastError());
	LB(cudaMemcpyAsync(&out.height, node_height, sizeof(int), cudaMemcpyDeviceToHost, s));
	LB(cudaStreamSynchronize(s));
	out.n_nodes = n - 1;
	return 5 + 2;  // kernels launched: five of ours + cub's sort and scan passes (counted as two)
}

// ---- quantised copy of the hierarchy (dev_types.h: BvhNodeQ) --------------------------------------------------
// Grid: the root box (union of node 0's child boxes) widened by four steps on every side, 65535 steps per axis.
// Planes: lo -> floor - 1, hi -> ceil + 1 (in double: the (centre, half-extent) pair of an fp32 node denotes c -+ h
// exactly), clamped to the grid — the widening makes the clamp unreachable for boxes inside the root box.
__global__ void k_quant_grid(const BvhNode *__restrict__ nodes, QGrid *grid) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	const BvhNode n = nodes[0];
	const double c[2][3] = { { n.b0.x, n.b0.z, n.b2.x }, { n.b1.x, n.b1.z, n.b2.z } }, h[2][3] = { { n.b0.y, n.b0.w, n.b2.y }, { n.b1.y, n.b1.w, n.b2.w } };
	double lo[3], hi[3], widest = 0.0;
	for (int k = 0; k < 3; ++k) {
		lo[k] = fmin(c[0][k] - h[0][k], c[1][k] - h[1][k]);
		hi[k] = fmax(c[0][k] + h[0][k], c[1][k] + h[1][k]);
		widest = fmax(widest, hi[k] - lo[k]);
	}
	for (int k = 0; k < 3; ++k) {
		// a scene with no extent along an axis (coplanar primitives) still gets a grid: at least 1e-4 of its widest axis, and
		// never finer than the fp32 coordinates themselves resolve
		const double mag = fmax(fabs(lo[k]), fabs(hi[k]));
		const double ext = fmax(fmax(hi[k] - lo[k], 1e-4 * widest), fmax(mag * 1e-4, 1e-20));
		const float step = (float)(ext / 65520.0);   // 65520 + 2 * 4 widening steps < 65535
		grid->step[k] = step;
		grid->lo[k] = (float)(lo[k] - 4.0 * (double)step);
		grid->hi[k] = (float)((double)grid->lo[k] + 65535.0 * (double)step);
	}
	grid->area32 = grid->area_q = grid->pad_ = 0.f;
}
__device__ __forceinline__ unsigned quant_pair(double c, double h, double glo, double inv_step) {
	double a = floor((c - h - glo) * inv_step) - 1.0, b = ceil((c + h - glo) * inv_step) + 1.0;
	a = fmin(fmax(a, 0.0), 65535.0);
	b = fmin(fmax(b, 0.0), 65535.0);
	return (unsigned)a | ((unsigned)b << 16);
}
__global__ void k_quant_nodes(int n, const BvhNode *__restrict__ nodes, QGrid *__restrict__ grid, BvhNodeQ *__restrict__ out) {
	const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = i0 < n;
	const int i = live ? i0 : n - 1;  // idle threads of the last warp stay for the warp sums below (they store nothing and add no area)
	const BvhNode nd = nodes[i];
	const double gx = grid->lo[0], gy = grid->lo[1], gz = grid->lo[2];
	const double ix = 1.0 / (double)grid->step[0], iy = 1.0 / (double)grid->step[1], iz = 1.0 / (double)grid->step[2];
	BvhNodeQ q;
	q.q[0] = quant_pair(nd.b0.x, nd.b0.y, gx, ix); q.q[1] = quant_pair(nd.b0.z, nd.b0.w, gy, iy); q.q[2] = quant_pair(nd.b2.x, nd.b2.y, gz, iz);
	q.q[3] = quant_pair(nd.b1.x, nd.b1.y, gx, ix); q.q[4] = quant_pair(nd.b1.z, nd.b1.w, gy, iy); q.q[5] = quant_pair(nd.b2.z, nd.b2.w, gz, iz);
	q.child[0] = nd.child[0]; q.child[1] = nd.child[1];
	if (live) out[i] = q;
	// what the padding costs: per child box, surface area as stored here over that of the fp32 box (a ray that reaches the
	// parent visits the child in proportion to it); the commit reads the mean over all child boxes
	const float sx = grid->step[0], sy = grid->step[1], sz = grid->step[2];
	float a32 = 0.f, aq = 0.f;  // number of boxes counted, sum of their ratios
#pragma unroll
	for (int c = 0; c < 2; ++c) {
		const float wx = 2.f * (c ? nd.b1.y : nd.b0.y), wy = 2.f * (c ? nd.b1.w : nd.b0.w), wz = 2.f * (c ? nd.b2.w : nd.b2.y);
		const float qx = sx * (float)((q.q[3 * c] >> 16) - (q.q[3 * c] & 0xffffu)), qy = sy * (float)((q.q[3 * c + 1] >> 16) - (q.q[3 * c + 1] & 0xffffu)),
			qz = sz * (float)((q.q[3 * c + 2] >> 16) - (q.q[3 * c + 2] & 0xffffu));
		const float s32 = wx * wy + wy * wz + wz * wx, sq = qx * qy + qy * qz + qz * qx;
		if (live && s32 > 0.f) { a32 += 1.f; aq += fminf(sq / s32, 1e6f); }
	}
	for (int off = 16; off > 0; off >>= 1) { a32 += __shfl_down_sync(0xffffffffu, a32, off); aq += __shfl_down_sync(0xffffffffu, aq, off); }
	if ((threadIdx.x & 31) == 0) { atomicAdd(&grid->area32, a32); atomicAdd(&grid->area_q, aq); }
}
// nodes -> (grid, out); two kernels on s
int launch_quantize_nodes(const BvhNode *nodes, int n, BvhNodeQ *out, QGrid *grid, cudaStream_t s) {
	if (n <= 0) return 0;
	k_quant_grid<<<1, 32, 0, s>>>(nodes, grid);
	k_quant_nodes<<<(n + 255) / 256, 256, 0, s>>>(n, nodes, grid, out);
	return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

}  // namespace areb
