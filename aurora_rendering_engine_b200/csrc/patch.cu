// patch.cu — patch-as-viewport renderer: host planner + fp64 device kernels (see patch.h).
//
// Built with -fmad=false (device) and -ffp-contract=off (host): every fp64 expression below keeps the operation
// order of /root/reference/experiments/rt10.cpp, so plans and texels are the reference's bit for bit.  Reference
// lines are cited per routine.  Nothing here is a CPU renderer: the planner only computes footprint geometry; all
// texels are produced by the kernels at the bottom.
#include "patch.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

namespace areb {

namespace {

// ---------------------------------------------------------------------------------------------------------
// small fp64 algebra shared by host and device (identical rounding on both: + - * / sqrt only)
// ---------------------------------------------------------------------------------------------------------
struct P2 { double x, y; };
struct P3 { double x, y, z; };
__host__ __device__ inline P3 operator-(P3 a, P3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
__host__ __device__ inline P3 operator+(P3 a, P3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
__host__ __device__ inline P3 operator*(P3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
__host__ __device__ inline P3 operator/(P3 a, double s) { return { a.x / s, a.y / s, a.z / s }; }
__host__ __device__ inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline P3 cross(P3 a, P3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
__host__ __device__ inline P2 operator-(P2 a, P2 b) { return { a.x - b.x, a.y - b.y }; }
__host__ __device__ inline P2 operator+(P2 a, P2 b) { return { a.x + b.x, a.y + b.y }; }
__host__ __device__ inline P2 operator*(P2 a, double s) { return { a.x * s, a.y * s }; }
__host__ __device__ inline double cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
// std::max(lo, std::min(hi, x)) as libstdc++ evaluates it (rt10.cpp:20-25)
__host__ __device__ inline double clampd(double x, double lo, double hi) {
	const double m = (x < hi) ? x : hi;
	return (lo < m) ? m : lo;
}
__host__ __device__ inline int clampi(int x, int lo, int hi) {
	const int m = (x < hi) ? x : hi;
	return (lo < m) ? m : lo;
}

const double kEps = 1e-9;  // rt10.cpp:18

// ---------------------------------------------------------------------------------------------------------
// host planner
// ---------------------------------------------------------------------------------------------------------
struct TriRef {
	P3 p[3];
	P2 uv[3];
};
struct MatRef {
	bool present, reflective;
	double albedo[3], metal;
};
struct Foot { P2 dst, src; };  // a footprint vertex: where it lands in the viewport's uv, which source uv it carries

inline P3 normal_of(const TriRef &t) {  // rt10.cpp:158-160 with normalize :71-75
	const P3 c = cross(t.p[1] - t.p[0], t.p[2] - t.p[0]);
	const double len = std::sqrt(dot(c, c));
	if (len < kEps) return { 0, 0, 0 };
	return c / len;
}
inline P3 centroid_of(const TriRef &t) { return (t.p[0] + t.p[1] + t.p[2]) / 3.0; }  // :161-163

// barycentrics of a point in the triangle's plane, rt10.cpp:172-191
inline void plane_bary(P3 q, const TriRef &t, double bc[3]) {
	const P3 e0 = t.p[1] - t.p[0], e1 = t.p[2] - t.p[0], e2 = q - t.p[0];
	const double d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(e2, e0), d21 = dot(e2, e1);
	const double den = d00 * d11 - d01 * d01;
	if (std::fabs(den) < kEps) { bc[0] = bc[1] = bc[2] = -1; return; }
	const double v = (d11 * d20 - d01 * d21) / den;
	const double w = (d00 * d21 - d01 * d20) / den;
	bc[0] = 1.0 - v - w; bc[1] = v; bc[2] = w;
}

// rt10.cpp:279-307: central projection of `point` from `eye` onto the viewport triangle's plane -> viewport uv
inline bool project(P3 eye, P3 point, const TriRef &vp, P2 &uv, P3 *hit) {
	const P3 n = normal_of(vp);
	const P3 dir = point - eye;
	const double den = dot(n, dir);
	if (std::fabs(den) < kEps) return false;
	const double t = dot(n, vp.p[0] - eye) / den;
	if (!(t > 1e-7 && t < 1.0 - 1e-7)) return false;  // the plane must separate eye and point
	const P3 h = eye + dir * t;
	double bc[3];
	plane_bary(h, vp, bc);
	if (bc[0] < -1e6 || bc[1] < -1e6 || bc[2] < -1e6) return false;
	uv = vp.uv[0] * bc[0] + vp.uv[1] * bc[1] + vp.uv[2] * bc[2];
	if (hit) *hit = h;
	return true;
}

// Sutherland-Hodgman against one edge, "left of A->B is inside" (rt10.cpp:326-376)
inline bool keeps(P2 A, P2 B, P2 q) { return cross2(B - A, q - A) >= -1e-12; }
inline Foot split(const Foot &S, const Foot &E, P2 A, P2 B) {
	const double ds = cross2(B - A, S.dst - A);
	const double de = cross2(B - A, E.dst - A);
	double t = ds / (ds - de + 1e-30);
	t = clampd(t, 0.0, 1.0);
	return { S.dst + (E.dst - S.dst) * t, S.src + (E.src - S.src) * t };
}
void clip_edge(P2 A, P2 B, const std::vector<Foot> &in, std::vector<Foot> &out) {
	out.clear();
	if (in.empty()) return;
	Foot S = in.back();
	bool s_in = keeps(A, B, S.dst);
	for (const Foot &E : in) {
		const bool e_in = keeps(A, B, E.dst);
		if (e_in) {
			if (!s_in) out.push_back(split(S, E, A, B));
			out.push_back(E);
		} else if (s_in) out.push_back(split(S, E, A, B));
		S = E;
		s_in = e_in;
	}
}

class Planner {
public:
	Planner(const PatchSceneView &sc, const PatchCfg &cfg, PatchPlan &plan) : cfg_(cfg), plan_(plan) {
		tris_.resize(sc.n_tri);
		mats_.resize(sc.n_tri);
		for (int i = 0; i < sc.n_tri; ++i) {
			for (int k = 0; k < 3; ++k) {
				tris_[i].p[k] = { sc.P[9 * i + 3 * k], sc.P[9 * i + 3 * k + 1], sc.P[9 * i + 3 * k + 2] };
				tris_[i].uv[k] = { sc.UV[6 * i + 2 * k], sc.UV[6 * i + 2 * k + 1] };
			}
			const int m = sc.material ? sc.material[i] : -1;
			MatRef &r = mats_[i];
			r.present = m >= 0 && m < sc.n_mat;
			r.reflective = r.present && sc.mat_type[m] != 0;
			for (int c = 0; c < 3; ++c) r.albedo[c] = r.present ? sc.mat_albedo[3 * m + c] : 1.0;  // rt10.cpp:565
			r.metal = r.present ? sc.mat_metalness[m] : 0.0;
		}
	}

	// What a painted footprint samples from: a node texture in the arena or a solid fill.
	struct Source {
		long long off;
		int w, h;
		double solid[3];
	};

	// rt10.cpp:462-485 + :591-595: triangles with a vertex that projects inside the viewport triangle, far to near.
	// The order is std::sort's (unstable introsort) on the reference's comparator — ties exist in symmetric scenes.
	void candidates(P3 eye, const TriRef &vp, std::vector<int> &out) const {
		out.clear();
		for (int i = 0; i < (int)tris_.size(); ++i) {
			bool any = false;
			for (int k = 0; k < 3 && !any; ++k) {
				P2 uv;
				P3 h;
				if (!project(eye, tris_[i].p[k], vp, uv, &h)) continue;
				double bc[3];
				plane_bary(h, vp, bc);
				any = bc[0] >= -1e-8 && bc[1] >= -1e-8 && bc[2] >= -1e-8;  // :193-196
			}
			if (any) out.push_back(i);
		}
		std::vector<double> dist(tris_.size());
		for (int i : out) {
			const P3 v = centroid_of(tris_[i]) - eye;
			dist[i] = std::sqrt(dot(v, v));
		}
		std::sort(out.begin(), out.end(), [&](int a, int b) { return dist[a] > dist[b]; });
	}

	// projection + clip of triangle t into viewport vp (rt10.cpp:603-623 / :711-731); false = nothing visible
	bool footprint(P3 eye, const TriRef &t, const TriRef &vp, std::vector<Foot> &poly) const {
		std::vector<Foot> a(3), b;
		for (int k = 0; k < 3; ++k) {
			P2 uv;
			if (!project(eye, t.p[k], vp, uv, nullptr)) return false;
			a[k] = { uv, t.uv[k] };
		}
		P2 c0 = vp.uv[0], c1 = vp.uv[1], c2 = vp.uv[2];
		if (cross2(c1 - c0, c2 - c0) < 0.0) std::swap(c1, c2);  // :354-358
		clip_edge(c0, c1, a, b);
		clip_edge(c1, c2, b, a);
		clip_edge(c2, c0, a, poly);
		return poly.size() >= 3;
	}
	static double footprint_area(const std::vector<Foot> &poly) {  // shoelace, rt10.cpp:314-324
		double a = 0.0;
		for (size_t i = 0; i < poly.size(); ++i) {
			const P2 p = poly[i].dst, q = poly[(i + 1) % poly.size()].dst;
			a += p.x * q.y - p.y * q.x;
		}
		return std::fabs(0.5 * a);
	}

	// Fan-triangulate a footprint into warp triangles for a dst_w x dst_h destination (rt10.cpp:384-460: vertex
	// scaling, bounding box and its clamping, degenerate-area rejection).
	void emit_ops(const std::vector<Foot> &poly, int dst_w, int dst_h, const Source &src, std::vector<PatchOp> &ops) const {
		for (size_t i = 1; i + 1 < poly.size(); ++i) {
			const Foot *v[3] = { &poly[0], &poly[i], &poly[i + 1] };
			PatchOp op;
			for (int k = 0; k < 3; ++k) {
				op.px[k] = v[k]->dst.x * (dst_w - 1);
				op.py[k] = v[k]->dst.y * (dst_h - 1);
				op.su[k] = v[k]->src.x;
				op.sv[k] = v[k]->src.y;
			}
			const double minx = std::floor(std::min({ op.px[0], op.px[1], op.px[2] })), maxx = std::ceil(std::max({ op.px[0], op.px[1], op.px[2] }));
			const double miny = std::floor(std::min({ op.py[0], op.py[1], op.py[2] })), maxy = std::ceil(std::max({ op.py[0], op.py[1], op.py[2] }));
			op.x0 = clampi((int)minx, 0, dst_w - 1); op.x1 = clampi((int)maxx, 0, dst_w - 1);
			op.y0 = clampi((int)miny, 0, dst_h - 1); op.y1 = clampi((int)maxy, 0, dst_h - 1);
			const P2 p0 = { op.px[0], op.py[0] }, p1 = { op.px[1], op.py[1] }, p2 = { op.px[2], op.py[2] };
			op.area = cross2(p1 - p0, p2 - p0);
			if (std::fabs(op.area) < 1e-12) continue;
			op.inv_area = 1.0 / op.area;
			op.src_off = src.off;
			op.src_w = src.w;
			op.src_h = src.h;
			std::memcpy(op.solid, src.solid, sizeof op.solid);
			ops.push_back(op);
		}
	}

	// rt10.cpp:551-664.  Returns what the parent samples: a solid fill for every terminating case, else a node.
	Source node(P3 eye, int cur, int tw, int th, int depth, double est_area) {
		tw = clampi(tw, cfg_.min_res, cfg_.max_res);
		th = clampi(th, cfg_.min_res, cfg_.max_res);
		const MatRef &m = mats_[cur];
		Source solid = { -1, tw, th, { m.albedo[0], m.albedo[1], m.albedo[2] } };
		if (!m.reflective) return solid;
		if (depth >= cfg_.max_depth) return solid;
		if (est_area > 0.0 && est_area < cfg_.min_area_px) return solid;
		for (int id : stack_)
			if (id == cur) return solid;  // mirror-sees-mirror cycle guard
		stack_.push_back(cur);

		const TriRef &vp = tris_[cur];
		const P3 n = normal_of(vp);
		const P3 mirrored = eye - n * (2.0 * dot(n, eye - vp.p[0]));  // :268-273
		std::vector<int> cand;
		candidates(mirrored, vp, cand);
		std::vector<PatchOp> ops;
		std::vector<Foot> poly;
		for (int t : cand) {
			if (t == cur) continue;
			if (!footprint(mirrored, tris_[t], vp, poly)) continue;
			const double area_px = footprint_area(poly) * (double)tw * (double)th;
			if (area_px < 0.5) continue;
			const int res = clampi((int)std::lround(std::sqrt(area_px) * 1.2), cfg_.min_res, cfg_.max_res);
			const Source child = node(mirrored, t, res, res, depth + 1, area_px);
			emit_ops(poly, tw, th, child, ops);
		}
		stack_.pop_back();

		PatchNode nd;
		nd.off = plan_.arena_texels;
		nd.w = tw;
		nd.h = th;
		nd.op_begin = (int)plan_.ops.size();
		plan_.ops.insert(plan_.ops.end(), ops.begin(), ops.end());
		nd.op_end = (int)plan_.ops.size();
		std::memcpy(nd.base, m.albedo, sizeof nd.base);
		nd.metal = clampd(m.metal, 0.0, 1.0);  // :651
		plan_.arena_texels += (long long)tw * th;
		if ((int)plan_.level_nodes.size() <= depth) plan_.level_nodes.resize(depth + 1);
		plan_.level_nodes[depth].push_back((int)plan_.nodes.size());
		plan_.nodes.push_back(nd);
		return Source{ nd.off, tw, th, { 0, 0, 0 } };
	}

	// rt10.cpp:683-742: one viewport triangle of the camera
	void camera_triangle(P3 eye, const TriRef &vp, int W, int H, int which) {
		std::vector<int> cand;
		candidates(eye, vp, cand);
		std::vector<PatchOp> ops;
		std::vector<Foot> poly;
		for (int t : cand) {
			if (!footprint(eye, tris_[t], vp, poly)) continue;
			const double area_px = footprint_area(poly) * (double)W * (double)H;
			if (area_px < 0.5) continue;
			const int res = clampi((int)std::lround(std::sqrt(area_px) * 1.0), cfg_.min_res, cfg_.max_res);
			const Source child = node(eye, t, res, res, 0, area_px);
			emit_ops(poly, W, H, child, ops);
		}
		plan_.vp_op_begin[which] = (int)plan_.ops.size();
		plan_.ops.insert(plan_.ops.end(), ops.begin(), ops.end());
		plan_.vp_op_end[which] = (int)plan_.ops.size();
	}

private:
	const PatchCfg &cfg_;
	PatchPlan &plan_;
	std::vector<TriRef> tris_;
	std::vector<MatRef> mats_;
	std::vector<int> stack_;
};

TriRef viewport_from(const double *P, const double *UV) {
	TriRef t;
	for (int k = 0; k < 3; ++k) {
		t.p[k] = { P[3 * k], P[3 * k + 1], P[3 * k + 2] };
		t.uv[k] = { UV[2 * k], UV[2 * k + 1] };
	}
	return t;
}

}  // namespace

void patch_plan_camera(const PatchSceneView &sc, const double origin[3], const double *vp_P, const double *vp_UV, int W, int H, const PatchCfg &cfg,
	PatchPlan &plan) {
	plan = PatchPlan();
	Planner pl(sc, cfg, plan);
	const P3 eye = { origin[0], origin[1], origin[2] };
	pl.camera_triangle(eye, viewport_from(vp_P, vp_UV), W, H, 0);
	pl.camera_triangle(eye, viewport_from(vp_P + 9, vp_UV + 6), W, H, 1);
}

void patch_plan_texture(const PatchSceneView &sc, const double origin[3], int current, int tex_w, int tex_h, double est_area_px, const PatchCfg &cfg,
	PatchPlan &plan) {
	plan = PatchPlan();
	Planner pl(sc, cfg, plan);
	const Planner::Source s = pl.node({ origin[0], origin[1], origin[2] }, current, tex_w, tex_h, 0, est_area_px);
	plan.root_w = s.w;
	plan.root_h = s.h;
	std::memcpy(plan.root_solid, s.solid, sizeof plan.root_solid);
	plan.root = s.off < 0 ? -1 : (int)plan.nodes.size() - 1;  // the root is planned last (post-order)
}

// =========================================================================================================
// device side
// =========================================================================================================
namespace {

struct C3 { double r, g, b; };
__device__ __forceinline__ C3 lerp3(C3 a, C3 b, double t) {  // a*(1-t) + b*t, rt10.cpp:76-78
	const double s = 1.0 - t;
	return { a.r * s + b.r * t, a.g * s + b.g * t, a.b * s + b.b * t };
}
__device__ __forceinline__ C3 texel(const double *__restrict__ arena, long long off, int w, int x, int y) {
	const double *p = arena + 3 * (off + (long long)y * w + x);
	return { p[0], p[1], p[2] };
}

// Image::sampleBilinear (rt10.cpp:98-116) of the op's source at (u,v).  A solid source goes through the same
// arithmetic: c*(1-t)+c*t is not always c in floating point, and the reference does compute it.
__device__ __forceinline__ C3 sample_source(const PatchOp &op, const double *__restrict__ arena, double u, double v) {
	u = clampd(u, 0.0, 1.0);
	v = clampd(v, 0.0, 1.0);
	const double fx = u * (op.src_w - 1), fy = v * (op.src_h - 1);
	const int x0 = (int)floor(fx), y0 = (int)floor(fy);
	const double tx = fx - x0, ty = fy - y0;
	C3 c00, c10, c01, c11;
	if (op.src_off < 0) {
		c00 = c10 = c01 = c11 = { op.solid[0], op.solid[1], op.solid[2] };
	} else {
		const int x1 = min(x0 + 1, op.src_w - 1), y1 = min(y0 + 1, op.src_h - 1);
		c00 = texel(arena, op.src_off, op.src_w, x0, y0);
		c10 = texel(arena, op.src_off, op.src_w, x1, y0);
		c01 = texel(arena, op.src_off, op.src_w, x0, y1);
		c11 = texel(arena, op.src_off, op.src_w, x1, y1);
	}
	return lerp3(lerp3(c00, c10, tx), lerp3(c01, c11, tx), ty);
}

// The painter's loop turned inside out: the reference paints ops [begin,end) in order and the last writer of a texel
// wins (rt10.cpp:420-436, overwrite mode), so a texel only needs the LAST op that covers it — walk backwards, stop
// at the first hit.  Coverage test = the reference's: inside the clamped integer bounding box and all three
// pixel-centre barycentrics >= -1e-6.
// One op against one texel: the reference's coverage test and fetch.
__device__ __forceinline__ bool paint_one(const PatchOp &op, const double *__restrict__ arena, int x, int y, P2 c, C3 &out) {
	if (x < op.x0 || x > op.x1 || y < op.y0 || y > op.y1) return false;
	const P2 p0 = { op.px[0], op.py[0] }, p1 = { op.px[1], op.py[1] }, p2 = { op.px[2], op.py[2] };
	const double c0 = cross2(p1 - c, p2 - c), c1 = cross2(p2 - c, p0 - c);
	{  // Conservative early reject without the two fp64 divisions: c*inv_area is within 3 ulp of c/area, so a barycentric
		// that misses the reference's -1e-6 threshold by more than the margin below misses it exactly as well.  Every
		// pixel that might be covered still takes the exact test, and only exact quotients reach the image.
		const double q0 = c0 * op.inv_area, q1 = c1 * op.inv_area;
		const double slack = 1e-12 * (1.0 + fabs(q0) + fabs(q1));
		if (q0 < -1e-6 - slack || q1 < -1e-6 - slack || 1.0 - q0 - q1 < -1e-6 - slack) return false;
	}
	const double w0 = c0 / op.area;
	const double w1 = c1 / op.area;
	const double w2 = 1.0 - w0 - w1;
	if (w0 < -1e-6 || w1 < -1e-6 || w2 < -1e-6) return false;
	const P2 s0 = { op.su[0], op.sv[0] }, s1 = { op.su[1], op.sv[1] }, s2 = { op.su[2], op.sv[2] };
	const P2 suv = s0 * w0 + s1 * w1 + s2 * w2;
	out = sample_source(op, arena, suv.x, suv.y);
	return true;
}

// Per-tile op culling: the CTA marks, in a shared bit set, the ops of [begin, end) whose bounding box touches its
// 16x16 tile; a texel then visits only those (typically 5-10 of ~60), still in the reference's order.
#define PATCH_CULL_WORDS 64  // up to 2048 ops per list through the bit set; longer lists fall back to the plain walk
__device__ __forceinline__ void cull_ops(const PatchOp *__restrict__ ops, int begin, int end, int tx0, int ty0, unsigned *bits) {
	const int n = end - begin;
	for (int w = threadIdx.x; w < PATCH_CULL_WORDS; w += blockDim.x) bits[w] = 0u;
	__syncthreads();
	if (n <= 32 * PATCH_CULL_WORDS)
		for (int i = threadIdx.x; i < n; i += blockDim.x) {
			const PatchOp &op = ops[begin + i];
			if (op.x1 >= tx0 && op.x0 <= tx0 + 15 && op.y1 >= ty0 && op.y0 <= ty0 + 15) atomicOr(&bits[i >> 5], 1u << (i & 31));
		}
	__syncthreads();
}
// The painter's loop turned inside out: the reference paints ops [begin,end) in order and the last writer of a texel
// wins (rt10.cpp:420-436, overwrite mode), so a texel only needs the LAST op that covers it — walk backwards, stop
// at the first hit.  Coverage test = the reference's: inside the clamped integer bounding box and all three
// pixel-centre barycentrics >= -1e-6.
__device__ __forceinline__ bool painted(const PatchOp *__restrict__ ops, int begin, int end, const unsigned *bits, const double *__restrict__ arena, int x,
	int y, C3 &out) {
	const P2 c = { (double)x + 0.5, (double)y + 0.5 };
	const int n = end - begin;
	if (n > 32 * PATCH_CULL_WORDS) {
		for (int i = end - 1; i >= begin; --i)
			if (paint_one(ops[i], arena, x, y, c, out)) return true;
		return false;
	}
	for (int w = (n - 1) >> 5; w >= 0; --w) {
		unsigned m = bits[w];
		while (m) {
			const int b = 31 - __clz(m);
			m &= ~(1u << b);
			if (paint_one(ops[begin + (w << 5) + b], arena, x, y, c, out)) return true;
		}
	}
	return false;
}

// All node textures of one recursion depth: one CTA per 16x16 texel tile.  reflection -> tint -> metalness blend
// (rt10.cpp:650-661).
__global__ void __launch_bounds__(256) k_patch_nodes(const PatchOp *__restrict__ ops, const PatchNode *__restrict__ nodes, const PatchTile *__restrict__ tiles,
	double *__restrict__ arena, double env_r, double env_g, double env_b) {
	__shared__ unsigned s_bits[PATCH_CULL_WORDS];
	const PatchTile t = tiles[blockIdx.x];
	const PatchNode &nd = nodes[t.node];
	const int x = t.tx * 16 + (threadIdx.x & 15), y = t.ty * 16 + (threadIdx.x >> 4);
	cull_ops(ops, nd.op_begin, nd.op_end, t.tx * 16, t.ty * 16, s_bits);
	if (x >= nd.w || y >= nd.h) return;
	C3 r = { env_r, env_g, env_b };
	painted(ops, nd.op_begin, nd.op_end, s_bits, arena, x, y, r);
	const double m = nd.metal, k = 1.0 - m;
	const C3 tint = { r.r * nd.base[0], r.g * nd.base[1], r.b * nd.base[2] };
	double *o = arena + 3 * (nd.off + (long long)y * nd.w + x);
	o[0] = nd.base[0] * k + tint.r * m;
	o[1] = nd.base[1] * k + tint.g * m;
	o[2] = nd.base[2] * k + tint.b * m;
}

struct CameraArgs {
	int W, H;
	int op_begin[2], op_end[2];
	double uv[2][6];    // the two viewport triangles' uv
	double env[3];
	const double *thresholds;  // 255 doubles: thresholds[v-1] = smallest c with encode(c) >= v (see encode_thresholds)
};

// pointInTriangle2D, rt10.cpp:202-218
__device__ __forceinline__ bool in_uv_triangle(P2 q, const double *uv, double eps) {
	const P2 A = { uv[0], uv[1] }, B = { uv[2], uv[3] }, C = { uv[4], uv[5] };
	const double c0 = cross2(B - A, q - A), c1 = cross2(C - B, q - B), c2 = cross2(A - C, q - C);
	const bool neg = (c0 < -eps) || (c1 < -eps) || (c2 < -eps);
	const bool pos = (c0 > eps) || (c1 > eps) || (c2 > eps);
	return !(neg && pos);
}
// Image::writePPM's 8-bit encode, rt10.cpp:124-141: lround(255 * pow(clamp(c,0,1), 1/gamma)).  The byte only depends on
// which of 255 thresholds c has passed, and those thresholds are found on the HOST with the very libm pow the reference
// calls (encode_thresholds below), so the device neither needs a pow (three per pixel were ~40 % of the kernel) nor has
// to agree with glibc's pow to the last ulp: encode(c) = number of thresholds <= c, an 8-step binary search.
__device__ __forceinline__ uint8_t encode8(double c, const double *__restrict__ thr) {
	c = clampd(c, 0.0, 1.0);
	int lo = 0, hi = 255;  // invariant: thr[0..lo) <= c, thr[hi..255) > c
#pragma unroll
	for (int step = 0; step < 8; ++step) {
		const int mid = (lo + hi) >> 1;
		if (lo < hi && __ldg(thr + mid) <= c) lo = mid + 1;
		else hi = lo < hi ? mid : hi;
	}
	return (uint8_t)lo;
}

// Camera image: both viewport triangles per pixel (each black outside its own triangle), summed and clamped
// (rt10.cpp:744-772), plus the gamma-encoded P6 payload.
__global__ void __launch_bounds__(256) k_patch_camera(const PatchOp *__restrict__ ops, const double *__restrict__ arena, CameraArgs a, double *__restrict__ rgb,
	uint8_t *__restrict__ rgb8) {
	__shared__ unsigned s_bits[2][PATCH_CULL_WORDS];
	const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
	cull_ops(ops, a.op_begin[0], a.op_end[0], blockIdx.x * 16, blockIdx.y * 16, s_bits[0]);
	cull_ops(ops, a.op_begin[1], a.op_end[1], blockIdx.x * 16, blockIdx.y * 16, s_bits[1]);
	if (x >= a.W || y >= a.H) return;
	const P2 uv = { (double)x / (a.W - 1), (double)y / (a.H - 1) };
	C3 part[2];
#pragma unroll
	for (int v = 0; v < 2; ++v) {
		C3 c = { a.env[0], a.env[1], a.env[2] };
		painted(ops, a.op_begin[v], a.op_end[v], s_bits[v], arena, x, y, c);
		if (!in_uv_triangle(uv, a.uv[v], 1e-10)) c = { 0.0, 0.0, 0.0 };
		part[v] = c;
	}
	const double r = clampd(part[0].r + part[1].r, 0.0, 1.0), g = clampd(part[0].g + part[1].g, 0.0, 1.0), b = clampd(part[0].b + part[1].b, 0.0, 1.0);
	const size_t i = ((size_t)y * a.W + x) * 3;
	if (rgb) { rgb[i] = r; rgb[i + 1] = g; rgb[i + 2] = b; }
	if (rgb8) { rgb8[i] = encode8(r, a.thresholds); rgb8[i + 1] = encode8(g, a.thresholds); rgb8[i + 2] = encode8(b, a.thresholds); }
}

__global__ void k_patch_fill(double *__restrict__ out, long long texels, double r, double g, double b) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= texels) return;
	out[3 * i] = r; out[3 * i + 1] = g; out[3 * i + 2] = b;
}

// thresholds[v-1] = the smallest double c in [0,1] whose reference encode is >= v, v = 1..255, by bisection over the bit
// patterns of the non-negative doubles (monotone in the value) with the host's pow — the function the reference calls.
void encode_thresholds(double gamma, double *thr255) {
	const double inv_gamma = 1.0 / gamma;  // rt10.cpp:131
	auto enc = [&](double c) { return clampi((int)std::lround(std::pow(clampd(c, 0.0, 1.0), inv_gamma) * 255.0), 0, 255); };
	auto as_bits = [](double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; };
	auto as_double = [](uint64_t u) { double d; std::memcpy(&d, &u, 8); return d; };
	const uint64_t one = as_bits(1.0);
	for (int v = 1; v <= 255; ++v) {
		if (enc(1.0) < v) { thr255[v - 1] = std::numeric_limits<double>::infinity(); continue; }  // never reached
		uint64_t lo = 0, hi = one;  // enc(as_double(hi)) >= v; enc(0) = 0 < v
		while (hi - lo > 1) {
			const uint64_t mid = lo + (hi - lo) / 2;
			if (enc(as_double(mid)) >= v) hi = mid;
			else lo = mid;
		}
		thr255[v - 1] = as_double(hi);
	}
}

int grow(void **p, size_t *cap, size_t bytes, std::string &err) {
	if (bytes <= *cap) return 0;
	if (*p) cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	const size_t want = bytes + bytes / 4 + 256;
	cudaError_t e = cudaMalloc(p, want);
	if (e != cudaSuccess) { err = std::string("cudaMalloc (patch workspace): ") + cudaGetErrorString(e); return (int)e; }
	*cap = want;
	return 0;
}
#define PCK(call)                                                                       \
	do {                                                                                \
		cudaError_t e_ = (call);                                                        \
		if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return (int)e_; } \
	} while (0)

// upload ops / nodes / tiles and launch every recursion level, deepest first; returns the number of launches
int run_levels(PatchWorkspace &ws, const PatchPlan &plan, const PatchCfg &cfg, PatchStats &st, cudaStream_t s, std::string &err) {
	std::vector<PatchTile> tiles;
	std::vector<std::pair<int, int>> ranges;  // per level: [first tile, count)
	for (int lv = (int)plan.level_nodes.size() - 1; lv >= 0; --lv) {
		const int first = (int)tiles.size();
		for (int ni : plan.level_nodes[lv]) {
			const PatchNode &nd = plan.nodes[ni];
			for (int ty = 0; ty < (nd.h + 15) / 16; ++ty)
				for (int tx = 0; tx < (nd.w + 15) / 16; ++tx) tiles.push_back({ ni, tx, ty });
		}
		if ((int)tiles.size() > first) ranges.push_back({ first, (int)tiles.size() - first });
	}
	int rc;
	if ((rc = grow(&ws.d_ops, &ws.cap_ops, plan.ops.size() * sizeof(PatchOp) + 1, err))) return rc;
	if ((rc = grow(&ws.d_nodes, &ws.cap_nodes, plan.nodes.size() * sizeof(PatchNode) + 1, err))) return rc;
	if ((rc = grow(&ws.d_tiles, &ws.cap_tiles, tiles.size() * sizeof(PatchTile) + 1, err))) return rc;
	if ((rc = grow(&ws.d_arena, &ws.cap_arena, (size_t)plan.arena_texels * 3 * sizeof(double) + 8, err))) return rc;
	if (!plan.ops.empty()) PCK(cudaMemcpyAsync(ws.d_ops, plan.ops.data(), plan.ops.size() * sizeof(PatchOp), cudaMemcpyHostToDevice, s));
	if (!plan.nodes.empty()) PCK(cudaMemcpyAsync(ws.d_nodes, plan.nodes.data(), plan.nodes.size() * sizeof(PatchNode), cudaMemcpyHostToDevice, s));
	if (!tiles.empty()) PCK(cudaMemcpyAsync(ws.d_tiles, tiles.data(), tiles.size() * sizeof(PatchTile), cudaMemcpyHostToDevice, s));
	st.h2d_bytes += plan.ops.size() * sizeof(PatchOp) + plan.nodes.size() * sizeof(PatchNode) + tiles.size() * sizeof(PatchTile);
	if (!ws.ev0) { PCK(cudaEventCreate(&ws.ev0)); PCK(cudaEventCreate(&ws.ev1)); }
	PCK(cudaEventRecord(ws.ev0, s));
	for (const auto &r : ranges) {
		k_patch_nodes<<<r.second, 256, 0, s>>>(static_cast<const PatchOp *>(ws.d_ops), static_cast<const PatchNode *>(ws.d_nodes),
			static_cast<const PatchTile *>(ws.d_tiles) + r.first, static_cast<double *>(ws.d_arena), cfg.env[0], cfg.env[1], cfg.env[2]);
		PCK(cudaGetLastError());
		++st.launches;
	}
	st.nodes = plan.nodes.size();
	st.node_texels = (uint64_t)plan.arena_texels;
	st.ops = plan.ops.size();
	st.levels = ranges.size();
	// (cudaMemcpyAsync from pageable memory returns once the data is staged, so `tiles` may go out of scope here)
	return 0;
}

}  // namespace

void PatchWorkspace::release() {
	void **bufs[] = { &d_ops, &d_nodes, &d_tiles, &d_arena, &d_rgb, &d_rgb8, &d_thr };
	thr_gamma = 0.0;
	for (void **b : bufs) {
		if (*b) cudaFree(*b);
		*b = nullptr;
	}
	cap_ops = cap_nodes = cap_tiles = cap_arena = cap_rgb = cap_rgb8 = 0;
	if (ev0) cudaEventDestroy(ev0);
	if (ev1) cudaEventDestroy(ev1);
	ev0 = ev1 = nullptr;
}

int patch_run_camera(PatchWorkspace &ws, const PatchPlan &plan, const double *vp_UV, int W, int H, const PatchCfg &cfg, double *out_rgb,
	uint8_t *out_rgb8, PatchStats &st, cudaStream_t s, std::string &err) {
	int rc = run_levels(ws, plan, cfg, st, s, err);
	if (rc) return rc;
	const size_t n = (size_t)W * H * 3;
	if (out_rgb && (rc = grow(&ws.d_rgb, &ws.cap_rgb, n * sizeof(double), err))) return rc;
	if (out_rgb8 && (rc = grow(&ws.d_rgb8, &ws.cap_rgb8, n, err))) return rc;
	CameraArgs a;
	a.W = W;
	a.H = H;
	for (int v = 0; v < 2; ++v) {
		a.op_begin[v] = plan.vp_op_begin[v];
		a.op_end[v] = plan.vp_op_end[v];
		std::memcpy(a.uv[v], vp_UV + 6 * v, 6 * sizeof(double));
	}
	std::memcpy(a.env, cfg.env, sizeof a.env);
	if (out_rgb8) {
		if (!ws.d_thr || ws.thr_gamma != cfg.gamma) {  // 255 x ~62 host pow calls: once per gamma value, kept in the workspace
			double thr[256];
			encode_thresholds(cfg.gamma, thr);
			thr[255] = std::numeric_limits<double>::infinity();  // sentinel: the search may look one past the last threshold
			if (!ws.d_thr) PCK(cudaMalloc(&ws.d_thr, sizeof thr));
			PCK(cudaMemcpyAsync(ws.d_thr, thr, sizeof thr, cudaMemcpyHostToDevice, s));
			PCK(cudaStreamSynchronize(s));  // thr lives on this stack frame
			ws.thr_gamma = cfg.gamma;
		}
	}
	a.thresholds = static_cast<const double *>(ws.d_thr);
	const dim3 grid((W + 15) / 16, (H + 15) / 16);
	k_patch_camera<<<grid, 256, 0, s>>>(static_cast<const PatchOp *>(ws.d_ops), static_cast<const double *>(ws.d_arena), a,
		out_rgb ? static_cast<double *>(ws.d_rgb) : nullptr, out_rgb8 ? static_cast<uint8_t *>(ws.d_rgb8) : nullptr);
	PCK(cudaGetLastError());
	++st.launches;
	PCK(cudaEventRecord(ws.ev1, s));
	if (out_rgb) { PCK(cudaMemcpyAsync(out_rgb, ws.d_rgb, n * sizeof(double), cudaMemcpyDeviceToHost, s)); st.d2h_bytes += n * sizeof(double); }
	if (out_rgb8) { PCK(cudaMemcpyAsync(out_rgb8, ws.d_rgb8, n, cudaMemcpyDeviceToHost, s)); st.d2h_bytes += n; }
	PCK(cudaStreamSynchronize(s));
	float ms = 0.f;
	PCK(cudaEventElapsedTime(&ms, ws.ev0, ws.ev1));
	st.kernel_ms = ms;
	return 0;
}

int patch_run_texture(PatchWorkspace &ws, const PatchPlan &plan, const PatchCfg &cfg, double *out_tex, PatchStats &st, cudaStream_t s, std::string &err) {
	int rc = run_levels(ws, plan, cfg, st, s, err);
	if (rc) return rc;
	const long long texels = (long long)plan.root_w * plan.root_h;
	const double *src;
	if (plan.root < 0) {  // every terminating case of rt10.cpp:567-581: the texture is the base colour
		if ((rc = grow(&ws.d_rgb, &ws.cap_rgb, (size_t)texels * 3 * sizeof(double), err))) return rc;
		k_patch_fill<<<(unsigned)((texels + 255) / 256), 256, 0, s>>>(static_cast<double *>(ws.d_rgb), texels, plan.root_solid[0], plan.root_solid[1],
			plan.root_solid[2]);
		PCK(cudaGetLastError());
		++st.launches;
		src = static_cast<const double *>(ws.d_rgb);
	} else src = static_cast<const double *>(ws.d_arena) + 3 * plan.nodes[plan.root].off;
	PCK(cudaEventRecord(ws.ev1, s));
	PCK(cudaMemcpyAsync(out_tex, src, (size_t)texels * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
	st.d2h_bytes += (size_t)texels * 3 * sizeof(double);
	PCK(cudaStreamSynchronize(s));
	float ms = 0.f;
	PCK(cudaEventElapsedTime(&ms, ws.ev0, ws.ev1));
	st.kernel_ms = ms;
	return 0;
}

}  // namespace areb
