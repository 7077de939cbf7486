// patch.h — the reference's patch-as-viewport renderer (/root/reference/experiments/rt10.cpp) on the GPU.
//
// The algorithm the library's Object::trace_texture was meant to run (reference include/object/object.h:37-38,
// experiments/Request.md:14): a reflective triangle is a viewport seen from the eye point mirrored across its plane;
// the triangles visible through it are painted far-to-near into its texture, recursively.  The reference interleaves
// recursion and rasterisation; here the host only PLANS (the geometry of every footprint, a few hundred flops each),
// and the texel work — every node texture of one recursion depth in one launch, deepest level first, then the
// camera image — runs on the device in fp64 with the reference's operation order (-fmad=false), so images are
// bit-identical to the reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace areb {

struct PatchSceneView {
	int n_tri = 0, n_mat = 0;
	const double *P = nullptr;         // n_tri * 9
	const double *UV = nullptr;        // n_tri * 6
	const int *material = nullptr;     // n_tri, out of range = none (white, diffuse)
	const int *mat_type = nullptr;     // n_mat, 0 diffuse / 1 reflective
	const double *mat_albedo = nullptr;     // n_mat * 3
	const double *mat_metalness = nullptr;  // n_mat
};
struct PatchCfg {
	int max_depth = 4;
	double min_area_px = 4.0;
	int max_res = 256, min_res = 16;
	double env[3] = { 0.08, 0.08, 0.10 };
	double gamma = 2.2;
};
struct PatchStats {
	uint64_t nodes = 0;         // reflective node textures rendered
	uint64_t node_texels = 0;   // their texels
	uint64_t ops = 0;           // warp triangles (fan triangles of clipped footprints)
	uint64_t levels = 0;        // recursion depths that held a node
	uint64_t launches = 0;      // kernels launched
	uint64_t h2d_bytes = 0, d2h_bytes = 0;
	double plan_ms = 0.0;       // host planning
	double kernel_ms = 0.0;     // device time of all launches (CUDA events on the launch stream)
};

// One warp triangle: destination triangle in pixel space of the texture being painted, source uv per vertex, and where
// the source texels are (arena offset in texels, or a solid colour).  160 bytes, read-only, broadcast across a warp.
struct PatchOp {
	double px[3], py[3];    // destination vertices (pixels)
	double area;            // cross(p1-p0, p2-p0)
	double inv_area;        // 1/area: only for the conservative early reject, never for a value that reaches the image
	double su[3], sv[3];    // source uv of the vertices
	double solid[3];        // colour of a solid source
	long long src_off;      // first texel of the source texture in the arena, -1 = solid
	int x0, x1, y0, y1;     // inclusive pixel bounding box, clamped as the reference clamps it
	int src_w, src_h;
};
struct PatchNode {
	long long off;          // first texel in the arena
	int w, h;
	int op_begin, op_end;   // painted in this order; the last op covering a texel wins
	double base[3];
	double metal;           // clamped to [0,1]
};
struct PatchTile { int node, tx, ty; };  // 16x16 texel tile of a node

struct PatchPlan {
	std::vector<PatchOp> ops;
	std::vector<PatchNode> nodes;                 // reflective nodes only
	std::vector<std::vector<int>> level_nodes;    // node indices per recursion depth
	long long arena_texels = 0;
	// camera pass (patch_render): ops of viewport triangle A then B
	int vp_op_begin[2] = { 0, 0 }, vp_op_end[2] = { 0, 0 };
	// trace_texture: the root node index, or -1 when the root is a solid fill (then root_solid/w/h describe it)
	int root = -1;
	int root_w = 0, root_h = 0;
	double root_solid[3] = { 0, 0, 0 };
};

// Host planners (no device work): exposed so CPU-only tests can check plans without a GPU.
void patch_plan_camera(const PatchSceneView &sc, const double origin[3], const double *vp_P, const double *vp_UV, int W, int H, const PatchCfg &cfg,
	PatchPlan &plan);
void patch_plan_texture(const PatchSceneView &sc, const double origin[3], int current, int tex_w, int tex_h, double est_area_px, const PatchCfg &cfg,
	PatchPlan &plan);

// Grow-only device / pinned workspace owned by the context.
struct PatchWorkspace {
	void *d_ops = nullptr, *d_nodes = nullptr, *d_tiles = nullptr, *d_arena = nullptr, *d_rgb = nullptr, *d_rgb8 = nullptr;
	void *d_thr = nullptr;   // 255 encode thresholds of thr_gamma (8-bit output)
	double thr_gamma = 0.0;
	size_t cap_ops = 0, cap_nodes = 0, cap_tiles = 0, cap_arena = 0, cap_rgb = 0, cap_rgb8 = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	void release();
};

// Execute a plan.  Return 0, or a cudaError_t (as int) with a message in err.
int patch_run_camera(PatchWorkspace &ws, const PatchPlan &plan, const double *vp_UV, int W, int H, const PatchCfg &cfg, double *out_rgb,
	uint8_t *out_rgb8, PatchStats &st, cudaStream_t s, std::string &err);
int patch_run_texture(PatchWorkspace &ws, const PatchPlan &plan, const PatchCfg &cfg, double *out_tex, PatchStats &st, cudaStream_t s, std::string &err);

}  // namespace areb
