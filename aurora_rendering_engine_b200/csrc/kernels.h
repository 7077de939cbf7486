// kernels.h — host-callable launchers of the sm_100a kernels (implemented in harness64.cu / render.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "dev_types.h"
#include "philox.cuh"
#include "render_args.h"

namespace areb {


// fp64 harness (harness64.cu, -fmad=false)
void launch_hit64(const DevScene &sc, int n, const double *Q, const double *D, double tmin, int *prim, double *t, double *P, double *N, double *uv, cudaStream_t s);
void launch_scatter64(const DevScene &sc, int n, const int *mat, const int *tex, const double *wi, const double *N, const double *P, const double *uv,
	const double *rnd, double *wo, double *att, double *emit, int *alive, cudaStream_t s);
void launch_texture64(const DevScene &sc, int n, const int *tex, const double *uv, const double *P, double *rgb, cudaStream_t s);
void launch_camera64(const CamBasis &cb, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D, cudaStream_t s);
void launch_philox(int n, uint64_t seed, const uint32_t *ctr, uint32_t *out, cudaStream_t s);

// library routines off the render loop, fp64 bit-exact batches (harness64.cu)
void launch_plane64(int n, const double *plane4, const double *Q, const double *D, int *hit, double *P, cudaStream_t s);
void launch_point_in64(int n, const double *tri9, const double *pts, int *inside, cudaStream_t s);
void launch_material_reflect64(int kind, int n, const double *plane4, const double *origin, int *ok, double *out, cudaStream_t s);

// Texture::paste (harness64.cu): destination / source images as w*h*3 doubles on the device
struct PasteArgs {
	double *dst;
	const double *src;
	int dw, dh, sw, sh;
	int x0, y0, x1, y1;       // clipped bounding box, inclusive
	double qx[4], qy[4];      // destination quad in boundary order LT, RT, RB, LB
	double hinv[9];           // inverse homography (destination -> source), row-major
};
void launch_paste(const PasteArgs &a, cudaStream_t s);

// fp32 harness: the render kernels' own device routines (render.cu)
void launch_hit32(const DevScene &sc, int n, const double *Q, const double *D, double tmin, int mode, int *prim, double *t, double *P, double *N, double *uv, cudaStream_t s);
void launch_scatter32(const DevScene &sc, int n, const int *mat, const int *tex, const double *wi, const double *N, const double *P, const double *uv,
	const double *rnd, double *wo, double *att, double *emit, int *alive, cudaStream_t s);
void launch_texture32(const DevScene &sc, int n, const int *tex, const double *uv, const double *P, double *rgb, cudaStream_t s);
void launch_camera32(const CamBasis &cb, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D, cudaStream_t s);

// render kernels. Returns the number of kernels launched, <0 on a launch-configuration error.
// mode: 0 brute force (shared memory), 1 BVH2, 2 compressed 8-wide BVH
int launch_render_path(const RenderArgs &a, int mode, bool count_tests, cudaStream_t s);
bool render_path_is_lean(const RenderArgs &a);  // brute force: the lean kernel is the one launched
bool render_path_lean_dims(const RenderArgs &a, int &blocks, int &threads, size_t &smem);  // launch shape of the lean (and baked) kernel
int render_path_big_nodes();                    // BVH_BIG_NODES
bool render_path_is_big(const RenderArgs &a);   // BVH: the high-occupancy build is the one launched
int launch_render_rtao(const RenderArgs &a, cudaStream_t s);
// wavefront schedule of the same path integrator (wavefront.cu): generate / extend / shade + compact over ray queues in HBM
size_t wavefront_workspace_bytes(int W, int H, int batch_spp);
int launch_render_wavefront(const RenderArgs &a, bool count_tests, int batch_spp, void *workspace, cudaStream_t s);
// gamma_thr: 256 floats, [v-1] = smallest c with the rt.cpp gamma encode >= v (v = 1..255), [255] = +inf (encoder 0 only)
void launch_tonemap(const float *accum, int W, int H, double inv_spp, int encoder, const float *gamma_thr, uint8_t *out, cudaStream_t s);
// dst[i] += sum_s src[s][i] over float4 indices [begin4, end4) and floats [tail_begin, tail_end): the per-device slice of
// the multi-device accumulator sum.  dst / src may live on other GPUs (peer-mapped memory over NVLink).
struct PeerReduceArgs {
	float *dst;
	const float *src[15];
	int n_src;
	size_t begin4, end4, tail_begin, tail_end;
};
void launch_peer_reduce(const PeerReduceArgs &a, int sm_count, cudaStream_t s);
// register-resident FFMA loop; returns FLOPs executed per launch
double launch_fp32_peak(float *sink, int sm_count, int iters, cudaStream_t s);

// streams an L2-resident buffer `passes` times with ld.global.cg.v4; returns bytes read per launch
double launch_l2_peak(const float *buf, size_t bytes, int sm_count, int passes, float *sink, cudaStream_t s);

size_t brute_smem_limit_prims();

// BVH2 construction on the device (lbvh.cu).  All pointers are device memory; the outputs are allocated by the caller:
// nodes[n_items - 1], prims[n_slots], ids[n_slots].  Returns the number of kernels launched, < 0 on a CUDA error (err).
struct LbvhInput {
	const f4 *item_lo, *item_hi;    // per hot item: conservative fp32 bounds; lo.w = the item's test kind as int bits (0 box, 1 quad, 2 triangle, 3 sphere)
	const HotPrim *item_prims;      // records in item order (a box takes two slots)
	const HotIds *item_ids;
	const int *item_slot;           // first slot of each item in item_prims / item_ids
	int n_items, n_slots;
	float cmin[3], cmax[3];         // bounds of the item box centres
};
struct LbvhOutput {
	BvhNode *nodes;
	HotPrim *prims;
	HotIds *ids;
	int n_nodes, height, root_leaf_ref;
};
// `workspace`: at least lbvh_workspace_bytes(n_items) bytes of device memory (scratch; reusable across builds)
size_t lbvh_workspace_bytes(int n_items);
int lbvh_build(const LbvhInput &in, LbvhOutput &out, void *workspace, size_t workspace_bytes, cudaStream_t s, std::string &err);

// Refit of the hierarchy lbvh_build left in `workspace` (same n_items; nothing else may have used the workspace since):
// m primitives moved — item index item[k] (= device primitive index in a scene without fused items), new record rec[k]
// and conservative bounds lo[k] / hi[k] — their leaf records and boxes are replaced and every node box is recomputed
// bottom-up over the unchanged topology.  Returns kernels launched, < 0 on a CUDA error.
// ... and of the per-primitive device arrays the shading stage and the fp64 harness read (device primitive index dp[k]):
// plane form / sphere record, the geometric normal inside the shading record, the fp64 (Q,u,v) / (c,r) copy (geo64: 9
// doubles per update, spheres use 4), the rt.cpp-style triangle record.
struct PrimScatterArgs {
	int m;
	const int *dp;
	const HotPrim *rec;
	const double *geo64;
	const f4 *rt;  // 3 per update
	HotPrim *prim_plane;
	ShadeRec *shade;
	double *tri64, *quad64, *sph64;
	f4 *rt_tris;
	int n_tri, n_quad;
};
void launch_prim_scatter(const PrimScatterArgs &a, cudaStream_t s);
// quantised copy of a BVH2 (dev_types.h: BvhNodeQ): grid from the root box, then every node; returns kernels launched (< 0: error)
int launch_quantize_nodes(const BvhNode *nodes, int n, BvhNodeQ *out, QGrid *grid, cudaStream_t s);
int lbvh_refit(void *workspace, size_t workspace_bytes, int n_items, int m, const int *item, const HotPrim *rec, const f4 *lo, const f4 *hi,
	HotPrim *prims, BvhNode *nodes, cudaStream_t s, std::string &err);

}  // namespace areb
