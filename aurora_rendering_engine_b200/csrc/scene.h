// scene.h — host-side scene container and scene compiler (flattening to the device layout of dev_types.h).
//
// The container holds exactly what a reference-style host program hands over: the (Q,u,v) of every
// are::Triangle in its are::ObjectSet (include/object/triangle.h:19, object_set.h:10-12) plus the material and
// texture each one points at, extended with the quad / sphere / material / texture kinds the north_star adds.
// compile() turns it into structure-of-arrays blocks ready for one cudaMemcpy each:
//   * device primitive order = triangles, quads, spheres (user ids kept in PrimInfo)
//   * plane-form fp32 record per primitive (dev_types.h)
//   * "hot" primitive list = what the render loop tests: coplanar triangle pairs forming a parallelogram are
//     fused into ONE quad test (halves the work on box-like scenes; the owning triangle is recovered after the
//     loop), everything else is passed through
//   * type-sorted brute-force list (small scenes) and a binned-SAH BVH2 with leaf-ordered primitives
#pragma once
#include <string>
#include <vector>

#include "dev_types.h"

namespace areb {

struct HostTexture {
	int kind = 0;
	double p[8] = { 0 };
	int w = 0, h = 0;
	std::vector<double> rgb;
};
struct HostMaterial {
	int kind = 0;
	double p[8] = { 0 };
};
struct HostPrim {
	int type = 0, mat = 0, tex = 0;
	double Q[3], u[3], v[3];  // sphere: Q = centre, u[0] = radius
	double uv[6] = { 0, 0, 1, 0, 0, 1 };
};

struct HostScene {
	std::vector<HostTexture> textures;
	std::vector<HostMaterial> materials;
	std::vector<HostPrim> prims;  // user order; index = user primitive id
};

struct CompiledScene {
	std::vector<HotPrim> brute;
	std::vector<HotIds> brute_ids;
	HotRange brute_range = { 0, 0, 0, 0, 0 };
	std::vector<HotIds> box_faces;  // 6 per box
	int n_boxes = 0;
	// Lean form of the brute list: valid (lean_ok) when the scene has no spheres, at most LEAN_MAX boxes / quad tests /
	// triangle tests, and every surface shades from its ShadeRec alone (solid colour + simple lobe, both halves of every
	// fused pair alike).  lean_shade holds one record per brute slot, six per box (one per face, zeros for an absent face).
	bool lean_ok = false;
	int lean_n_open = 0;  // the first lean_n_open boxes of the brute list have an absent face
	std::vector<ShadeRec> lean_shade;
	std::vector<int> lean_sbase;
	std::vector<HotPrim> bvh_prims;
	std::vector<HotIds> bvh_ids;
	std::vector<BvhNode> nodes;
	std::vector<Bvh4Node> nodes4;  // 4-wide collapse of `nodes` (CompileOptions::build_bvh4; empty otherwise or when too deep for its stack)
	int bvh4_depth = 0;
	std::vector<WideNode> wnodes;  // compressed 8-wide collapse of `nodes` (empty when the scene has a single hot item)
	std::vector<HotPrim> wide_prims;
	std::vector<HotIds> wide_ids;
	std::vector<unsigned char> wide_kinds;
	int wide_depth = 0;
	int root_leaf_meta = 0;
	int n_hot = 0, n_fused_pairs = 0;
	std::vector<PrimInfo> info;
	std::vector<HotPrim> prim_plane;
	std::vector<ShadeRec> shade;
	std::vector<float> tri_uv;
	std::vector<f4> rt_tris;  // 3 per triangle: rt.cpp-style fp32 record (v0, e1, e2, n)
	std::vector<double> tri64, quad64, sph64, tri_uv64;
	int n_tri = 0, n_quad = 0, n_sph = 0;
	std::vector<MaterialRec> mats;
	std::vector<TextureRec> texs;
	std::vector<float> tex_data;
	int bvh_depth = 0;
	// Input of the device BVH builder (CompileOptions::device_bvh; lbvh.cu) instead of nodes / bvh_prims / bvh_ids:
	// per hot item conservative fp32 bounds (lo.w = test kind as int bits), records in item order, first slot per item.
	std::vector<f4> lb_lo, lb_hi;
	std::vector<HotPrim> lb_prims;
	std::vector<HotIds> lb_ids;
	std::vector<int> lb_slot;
	float lb_cmin[3] = { 0, 0, 0 }, lb_cmax[3] = { 0, 0, 0 };  // bounds of the item box centres
	double host_bvh_ms = 0.0;  // time spent in the host BVH builders (0 under device_bvh)

	// Empty the scene but keep every array's storage: a re-commit of a large scene then writes into memory that is
	// already mapped instead of page-faulting ~150 MB per million primitives in again.
	void reset() {
		brute.clear(); brute_ids.clear(); box_faces.clear(); bvh_prims.clear(); bvh_ids.clear(); nodes.clear(); nodes4.clear(); wnodes.clear();
		bvh4_depth = 0;
		wide_prims.clear(); wide_ids.clear(); wide_kinds.clear(); mats.clear(); texs.clear(); tex_data.clear();
		lean_shade.clear(); lean_sbase.clear();
		// NOT cleared: the per-primitive arrays (info, prim_plane, shade, tri_uv, rt_tris, tri64, quad64, sph64, tri_uv64) and
		// the device-builder input (lb_*).  compile_scene resizes them to the new counts and overwrites every element, so a
		// re-commit of a scene of the same size does not zero-fill ~250 bytes per primitive on one thread first.
		brute_range = { 0, 0, 0, 0, 0 };
		n_boxes = 0; lean_ok = false; lean_n_open = 0; wide_depth = 0; root_leaf_meta = 0; n_hot = 0; n_fused_pairs = 0;
		n_tri = n_quad = n_sph = 0; bvh_depth = 0; host_bvh_ms = 0.0;
		for (int k = 0; k < 3; ++k) { lb_cmin[k] = 0.f; lb_cmax[k] = 0.f; }
	}
};

struct CompileOptions {
	bool fuse_parallelograms = true;
	bool fuse_boxes = true;
	int brute_max = 1024;
	bool build_wide = false;  // also build the compressed 8-wide BVH for scenes beyond 65536 BVH2 nodes
	bool build_bvh4 = false;  // also collapse the host-built BVH2 into 128-byte 4-wide nodes (ARE_TRAVERSAL_BVH4)
	bool device_bvh = false;  // leave the hierarchy to the device builder (lbvh.cu): emit its input instead of nodes
};

// Validation mirroring are::Triangle's ctor (src/object/triangle.cpp:22-36): returns nullptr when fine, else the
// reference's exception message.
const char *validate_edges(const double u[3], const double v[3]);

// Host threads used by the scene compiler (0 = all hardware threads).  The compiled scene does not depend on the count.
void set_build_threads(int n);

// Returns false and fills err on inconsistent ids etc.
bool compile_scene(const HostScene &hs, const CompileOptions &opt, CompiledScene &out, std::string &err);

// Refit support (are_cuda_refit): the derived records of ONE primitive whose geometry changed, by the formulas
// compile_scene uses — plane form / sphere record, conservative fp32 bounds (lo.w = the test kind as int bits, the
// device builder's convention), rt.cpp-style record for triangles.  False for geometry the validation rejects.
struct PrimUpdate {
	HotPrim rec;
	f4 lo, hi;
	f4 rt[3];
};
bool prim_update_records(const HostPrim &p, PrimUpdate &out);

// Perlin tables: 256 unit gradients (xyz doubles) + 3x256 permutations from Philox(seed); spec shared with the oracle.
void make_noise_tables(uint64_t seed, double *grad768, int *perm768);

// camera basis exactly as experiments/rt.cpp:339-343 computes it, in fp64
void make_cam_basis(const double pos[3], const double target[3], const double up[3], double vfov_deg, double focus_dist,
	double defocus_angle_deg, int jitter, int W, int H, CamBasis &out);

// the same camera with rt.cpp's own fp32 operations (experiments/rt.cpp:339-343), for the RT_AO integrator
void make_rt_cam(const double pos[3], const double target[3], const double up[3], double vfov_deg, int W, int H, RtCam &out);

}  // namespace areb
