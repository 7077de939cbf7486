// dev_types.h — plain records shared by the host scene compiler (g++) and the sm_100a kernels (nvcc).
//
// HBM layout.  Everything the fp32 render kernels touch per ray is packed in 16-byte lanes so
// one LDS.128 / LDG.128 moves a quarter of a primitive:
//
//   "plane form" primitive, 3 x float4 (48 B) — used for triangles, parallelograms (quads) and fused
//   triangle pairs alike:
//       r0 = (n.x, n.y, n.z, d0)     unit plane normal, d0 = n·Q           t = (d0 - n·o) / (n·d)
//       r1 = (A.x, A.y, A.z, aQ)     alpha = A·p - aQ   with A = (v x N)/(N·N), N = u x v
//       r2 = (B.x, B.y, B.z, bQ)     beta  = B·p - bQ   with B = (N x u)/(N·N)
//   sphere, same 48 B slot:  r0 = (c.x, c.y, c.z, r), r1.x = r*r
//   parallelepiped ("box": up to six parallelograms — fused triangle pairs or quads — that are the faces of one
//   parallelepiped, e.g. the Cornell boxes, or the room itself with its front face absent), TWO slots (96 B):
//       r0..r2 = (n_i.x, n_i.y, n_i.z, c_i)   unit slab normals, c_i = n_i·centre          i = 0,1,2
//       r3     = (h_0, h_1, h_2, face mask)   slab half-widths; mask bit 2i / 2i+1 = face at c_i -/+ h_i present
//     one test = three slab pairs = at most two candidate hits (entry, exit) instead of six plane-form tests.
//
// The fp64 copies (Q,u,v as the host gave them) feed only the decision-exact harness kernels.
#pragma once
#include "rtc_compat.h"

#if defined(__CUDACC__)
#define ARE_HD __host__ __device__
#else
#define ARE_HD
#endif

namespace areb {

enum PrimType : int { PT_TRIANGLE = 0, PT_QUAD = 1, PT_SPHERE = 2 };
enum MatKind : int { MK_DIFFUSE = 0, MK_REFLECTIVE = 1, MK_LAMBERTIAN = 2, MK_METAL = 3, MK_DIELECTRIC = 4, MK_LIGHT = 5 };
enum TexKind : int { TK_SOLID = 0, TK_CHECKER_UV = 1, TK_CHECKER_3D = 2, TK_NOISE = 3, TK_IMAGE = 4 };

struct f4 { float x, y, z, w; };  // layout-compatible with CUDA's float4 (16-byte aligned by the allocator)

struct HotPrim { f4 r0, r1, r2; };  // 48 B

// ids of the user primitives behind one hot primitive: b >= 0 only for a fused triangle pair
// (a = the triangle on the alpha >= beta side of the diagonal and the lower user id).
// b <= -2 marks a pair whose two triangles shade identically (same material, same position-only texture, same
// normal): the second triangle is -2 - b and the render loop need not find out which half was hit.
// For a box slot: a = -1 - box index (its six per-face HotIds live in DevScene::box_faces[6*index + face]).
struct HotIds { int a, b; };

struct PrimInfo { int user_id, mat, tex, type; };  // per device primitive (device order: triangles, quads, spheres)

// Everything the shading stage needs for the common case, 32 B per device primitive (two LDG.128 per hit instead
// of ~10 dependent loads through info -> material -> texture tables):
//   r0 = (N.x, N.y, N.z, bits)   geometric unit normal (spheres: unused), bits = material kind | SHADE_FAST << 8
//   r1 = (c.r, c.g, c.b, p0)     albedo (emitted radiance for lights, already scaled); p0 = fuzz | ior
// SHADE_FAST is set when the effective texture is a solid colour and the material is Diffuse / Lambertian / Metal /
// Dielectric / DiffuseLight; other primitives take the general path (texture evaluation, Reflective lobe choice).
// bits: [0..7] material kind, [8] SHADE_FAST, [9..11] kind of the texture that colours the primitive, [12..27] its id
// (so the render loop knows "this hit needs noise texture 3" without walking info -> material -> texture tables)
struct ShadeRec { f4 r0, r1; };
enum { SHADE_FAST = 1, SHADE_TEXKIND_SHIFT = 9, SHADE_TEXID_SHIFT = 12, SHADE_TEXID_MASK = 0xffff };

struct MaterialRec {
	int kind, pad_;
	double p[8];
	float pf[8];
};

struct TextureRec {
	int kind, w, h, pad_;
	long long data_off;  // IMAGE: float offset into tex_data (rgb per texel). NOISE: float offset of 256 float4 gradients, followed by 768 ints and 768 doubles
	double p[8];
	float pf[8];
};

// BVH2 node, 64 B, both children's boxes inline (one node fetch = 4 x LDG.128), each axis as (centre, half-extent):
//   b0 = (c0.cx, c0.hx, c0.cy, c0.hy)   b1 = (c1.cx, c1.hx, c1.cy, c1.hy)   b2 = (c0.cz, c0.hz, c1.cz, c1.hz)
//   slab distances = centre*inv -/+ half*|inv|: FMA-pipe work only, no per-axis min/max on the (binding) ALU pipe
//   child[i] >= 0: inner node index.  child[i] < 0: leaf, first hot primitive = ~child[i],
//   meta[i] = nq | nt << 8 | ns << 16 | nb << 24  (boxes first — two slots each — then quads+fused pairs,
//   triangles, spheres)
struct BvhNode {
	f4 b0, b1, b2;
	int child[2];
	int meta[2];
};

// Quantised BVH2 node, 32 B = ONE 256-bit load (sm_100 LDG.E.256): both children's boxes as 16-bit planes on a grid over
// the root box (QGrid: plane = lo + q * step per axis), padded outwards by one step.  The traversal of a hierarchy that
// lives in L2 is bound by the L1 data pipe — load requests x distinct lines per request, 91–96 % busy on the
// 1 M-primitive scene — not by latency and not by issue slots (50 %): halving the requests per node visit is worth the few
// extra integer instructions of the decode.  Built on the device from the fp32 nodes (lbvh.cu: k_quant_nodes) for
// hierarchies beyond BVH_BIG_NODES nodes; the fp32 nodes stay (harnesses, refit, the fall-back kernel).
//   q[3 * c + axis] = lo16 | hi16 << 16 of child c;  child[] as in BvhNode
struct BvhNodeQ {
	unsigned q[6];
	int child[2];
};
struct QGrid {
	float lo[3], step[3];
	float hi[3];  // lo + 65535 * step: the far corner (the launch-time guard on the camera's distance reads it)
	float area32, area_q;  // k_quant_nodes: child boxes counted / sum of (surface area as quantised : surface area of the fp32 box)
	float pad_;
};

// Uncompressed 4-wide BVH node, 128 B = one cache line, eight 16-byte loads: the boxes of four children as fp32
// (centre, half-extent) per axis, child k in component k.  Collapsed from the BVH2 (a node absorbs its inner child of
// largest area until it has four), so one visit replaces up to two dependent BVH2 visits — for hierarchies that live in
// L2, where the traversal is bound by the latency of its dependent node fetches, not by instruction issue.
//   child[k] >= 0: inner BVH4 node; < 0: leaf reference as in the BVH2; an empty slot has half-extent -1e30 (never hit)
struct Bvh4Node {
	f4 cx, hx, cy, hy, cz, hz;
	int child[4];
	int pad_[4];
};
#define ARE_BVH4_STACK 96  // entries of the per-thread BVH4 stack: up to three postponed children per level

// Compressed 8-wide BVH node, 80 B = five 16-byte loads (layout after Ylitie, Karras, Laine, "Efficient incoherent
// ray traversal on GPUs through compressed wide BVHs", HPG 2017): child boxes are 8-bit offsets on a per-node grid
// origin + 2^e per axis, so 1 M primitives need ~20 MB of nodes instead of 64 MB of BVH2 nodes and a ray makes
// a third of the dependent memory round trips.
//   n0 = (origin.x, origin.y, origin.z, e_x | e_y << 8 | e_z << 16 | imask << 24)   e_k = biased float exponent of the grid step
//   n1 = (first inner child node, first leaf primitive slot, meta[0..3], meta[4..7])
//        meta byte: 0 = empty slot; bit 5 = present; bits 0-4 = bit position of the child in the hit mask:
//        inner child in slot s -> 24 + s (slot order = spatial octant order, so XOR with the ray octant gives a
//        front-to-back visiting order), leaf -> offset of its primitive from the node's first slot (0..23)
//   n2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])   n3 = (qlo.z.., qlo.z.., qhi.x.., qhi.x..)
//   n4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// imask bit s = slot s holds an inner node; inner children are stored contiguously in slot order.
struct u4 { unsigned x, y, z, w; };
struct WideNode {
	f4 n0;
	u4 n1, n2, n3, n4;
};

// A contiguous run of hot primitives sorted by test kind: nb boxes (two slots each) from `first`, then nq quad
// tests, nt triangle tests, ns sphere tests.  The brute-force list is one big range; every BVH leaf is a small one.
struct HotRange { int first, nq, nt, ns, nb; };

#define ARE_BVH_STACK 48  // entries of the per-thread BVH2 traversal stack = the deepest hierarchy the kernels accept

enum { LEAN_MAX = 4 };  // boxes, quad tests and triangle tests (each) that the lean render kernel unrolls

struct DevScene {
	// --- fp32 render data ---
	const HotPrim *brute;      // type-sorted hot primitives (nullptr when the scene is too big for the brute path)
	const HotIds *brute_ids;
	HotRange brute_range;
	const HotPrim *bvh_prims;  // leaf-ordered hot primitives
	const HotIds *bvh_ids;
	const BvhNode *nodes;
	int n_nodes;
	int root_leaf_meta;        // when the whole scene is one leaf (n_nodes == 0)
	const BvhNodeQ *nodes_q;   // quantised copy of `nodes` (nullptr when not built: small hierarchies, option off)
	const QGrid *qgrid;        // its grid
	const Bvh4Node *nodes4;    // 4-wide collapse of `nodes` over the same leaves (nullptr when not built)
	int n_nodes4;
	const WideNode *wnodes;    // compressed 8-wide hierarchy over the same hot items (nullptr when n_nodes == 0)
	const HotPrim *wide_prims; // its leaf-ordered primitives, ids and test kinds (0 box, 1 quad / fused pair, 2 triangle, 3 sphere)
	const HotIds *wide_ids;
	const unsigned char *wide_kinds;
	int n_wnodes;
	int n_hot;                 // slots in the hot arrays (a box takes two)
	const HotIds *box_faces;   // 6 per box: the fused pair / quad behind each face, {-1,-1} when the face is absent
	// "lean" tables of a small flat-shaded scene (scene.h: CompiledScene::lean_ok): the shading record of every brute
	// slot — six per box, one per face — so the render loop goes hit -> record without any id / owner resolution
	const ShadeRec *lean_shade;
	const int *lean_sbase;     // per brute slot: index of its first record in lean_shade
	int n_lean_shade;
	int lean_ok;
	int lean_n_open;           // brute list: the first lean_n_open boxes are open (one face absent), the others closed
	// --- per device primitive ---
	const PrimInfo *info;
	const ShadeRec *shade;     // per device primitive
	const HotPrim *prim_plane; // plane form (or sphere record) of every user primitive, device order
	const float *tri_uv;       // 6 floats per triangle
	// rt.cpp-style fp32 triangle records for the RT_AO integrator, 3 x float4 per triangle:
	//   (v0.xyz, e1.x) (e1.yz, e2.xy) (e2.z, n.xyz)   with n = normalize(e1 x e2) computed in fp32 as rt.cpp:121 does
	const f4 *rt_tris;
	int n_tri, n_quad, n_sph;
	// --- fp64 harness data (host values, untouched) ---
	const double *tri64;       // 9 doubles per triangle: Q, u, v
	const double *quad64;      // 9 doubles per quad
	const double *sph64;       // 4 doubles per sphere: c, r
	const double *tri_uv64;    // 6 doubles per triangle
	// --- materials / textures ---
	const MaterialRec *mats;
	const TextureRec *texs;
	const float *tex_data;
	int n_mat, n_tex;
	int has_noise;             // some texture is TK_NOISE: the render loop runs its cooperative turbulence stage
};

// camera basis precomputed on the host in fp64 (mirrors oracle cam_setup; experiments/rt.cpp:339-343)
struct CamBasis {
	double pos[3], fwd[3], right[3], up[3];
	double sx, sy, lens_r, focus;
	int jitter, pad_;
};

// the same basis rounded to fp32 once on the host: what the render kernels read (no F2F in the loop)
struct CamF {
	float pos[3], fwd[3], right[3], up[3];
	float sx, sy, lens_r, focus;
	int jitter, pad_;
};

// camera of the RT_AO integrator: the basis computed in fp32 with the very operations of experiments/rt.cpp:339-343
struct RtCam {
	float pos[3], fwd[3], right[3], up[3];
	float aspect, scale;
};

}  // namespace areb
