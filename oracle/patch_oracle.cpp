// oracle/patch_oracle.cpp — TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's patch-as-viewport renderer
// (/root/reference/experiments/rt10.cpp), the algorithm the library's Object::trace_texture was meant to run
// (include/object/object.h:37-38, experiments/Request.md:14).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may load it; the product (csrc/patch.cu) never does.
//
// Pinned (tests/test_patch.py): (1) the rt10 scene rendered through this file is byte-identical to the reference's
// shipped experiments/output_rt10.ppm (sha256 committed in tests/golden/patch_vectors.npz, full compare when
// /root/reference is present); (2) random scenes / configurations vs the REAL rt10.cpp compiled into
// oracle/_ref/librt10_ref.so (ref_patch_harness.cpp), recorded in tests/golden/patch_vectors.npz.
//
// C++ rather than C for one reason: the far-to-near order of the candidates is whatever std::sort (libstdc++
// introsort, unstable) makes of the comparator "distance(a) > distance(b)" (rt10.cpp:591-595, :698-702); ties do
// occur (mirror-symmetric scenes), so the restatement has to run the same algorithm on the same sequence.
// Arithmetic is fp64, written operation for operation in the reference's order; built with -ffp-contract=off.
//
// Section map (reference lines in experiments/rt10.cpp):
//   bilinear fetch                      :98-116      fetch_bilinear
//   projection to viewport uv           :279-307     project_uv (barycentric3D :172-191)
//   Sutherland-Hodgman in uv space      :330-382     clip_to_triangle
//   warp rasteriser                     :384-460     raster_fan
//   candidate collection                :462-485     collect
//   recursive node render               :551-664     node_texture
//   camera                              :677-772     orc_patch_render
//   8-bit encode                        :124-141     encode8
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

const double EPS9 = 1e-9;

struct D2 { double x, y; };
struct D3 { double x, y, z; };
inline D3 sub(D3 a, D3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline D3 add(D3 a, D3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline D3 scl(D3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
inline D3 dvd(D3 a, double s) { return { a.x / s, a.y / s, a.z / s }; }
inline double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 crs(D3 a, D3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double len3(D3 a) { return std::sqrt(dot3(a, a)); }
inline D3 unit(D3 a) {  // :71-75
	double l = len3(a);
	if (l < EPS9) return { 0, 0, 0 };
	return dvd(a, l);
}
inline D3 mix(D3 a, D3 b, double t) { return add(scl(a, 1.0 - t), scl(b, t)); }  // :76-78
inline double cr2(D2 a, D2 b) { return a.x * b.y - a.y * b.x; }
inline D2 sub2(D2 a, D2 b) { return { a.x - b.x, a.y - b.y }; }
inline D2 add2(D2 a, D2 b) { return { a.x + b.x, a.y + b.y }; }
inline D2 scl2(D2 a, double s) { return { a.x * s, a.y * s }; }
// std::max(lo, std::min(hi, x)) with the comparison directions of libstdc++ (NaN behaviour included)  :20-25
inline double clampd(double x, double lo, double hi) {
	double m = (x < hi) ? x : hi;
	return (lo < m) ? m : lo;
}
inline int clampi(int x, int lo, int hi) {
	int m = (x < hi) ? x : hi;
	return (lo < m) ? m : lo;
}

struct Img {
	int w = 0, h = 0;
	std::vector<D3> px;
	Img() {}
	Img(int w_, int h_, D3 fill) : w(w_), h(h_), px((size_t)w_ * (size_t)h_, fill) {}
	D3 at(int x, int y) const { return px[(size_t)clampi(y, 0, h - 1) * (size_t)w + (size_t)clampi(x, 0, w - 1)]; }
	void put(int x, int y, D3 c) {
		if (x < 0 || x >= w || y < 0 || y >= h) return;
		px[(size_t)y * (size_t)w + (size_t)x] = c;
	}
};

D3 fetch_bilinear(const Img &im, double u, double v) {
	u = clampd(u, 0.0, 1.0);
	v = clampd(v, 0.0, 1.0);
	double fx = u * (im.w - 1), fy = v * (im.h - 1);
	int x0 = (int)std::floor(fx), y0 = (int)std::floor(fy);
	int x1 = std::min(x0 + 1, im.w - 1), y1 = std::min(y0 + 1, im.h - 1);
	double tx = fx - x0, ty = fy - y0;
	D3 top = mix(im.at(x0, y0), im.at(x1, y0), tx);
	D3 bot = mix(im.at(x0, y1), im.at(x1, y1), tx);
	return mix(top, bot, ty);
}

struct Mat { int reflective; D3 albedo; double metalness; };
struct Tri { int id; D3 p[3]; D2 uv[3]; const Mat *mat; };
struct Cfg { int max_depth; double min_area; int max_res, min_res; D3 env; };

inline D3 tri_normal(const Tri &t) { return unit(crs(sub(t.p[1], t.p[0]), sub(t.p[2], t.p[0]))); }
inline D3 tri_centroid(const Tri &t) { return dvd(add(add(t.p[0], t.p[1]), t.p[2]), 3.0); }

bool bary3(D3 P, const Tri &t, double bc[3]) {  // :172-191; false = degenerate (-1,-1,-1)
	D3 a = sub(t.p[1], t.p[0]), b = sub(t.p[2], t.p[0]), c = sub(P, t.p[0]);
	double aa = dot3(a, a), ab = dot3(a, b), bb = dot3(b, b), ca = dot3(c, a), cb = dot3(c, b);
	double den = aa * bb - ab * ab;
	if (std::fabs(den) < EPS9) { bc[0] = bc[1] = bc[2] = -1; return false; }
	double v = (bb * ca - ab * cb) / den;
	double w = (aa * cb - ab * ca) / den;
	bc[0] = 1.0 - v - w; bc[1] = v; bc[2] = w;
	return true;
}
bool inside3(D3 P, const Tri &t, double eps) {  // :193-196
	double bc[3];
	bary3(P, t, bc);
	return bc[0] >= -eps && bc[1] >= -eps && bc[2] >= -eps;
}

// :279-307 — the viewport plane must lie strictly between origin and point
bool project_uv(D3 origin, D3 point, const Tri &vp, D2 &uv, D3 *hit_out) {
	D3 n = tri_normal(vp);
	D3 dir = sub(point, origin);
	double den = dot3(n, dir);
	if (std::fabs(den) < EPS9) return false;
	double t = dot3(n, sub(vp.p[0], origin)) / den;
	if (!(t > 1e-7 && t < 1.0 - 1e-7)) return false;
	D3 hit = add(origin, scl(dir, t));
	double bc[3];
	bary3(hit, vp, bc);
	if (bc[0] < -1e6 || bc[1] < -1e6 || bc[2] < -1e6) return false;
	uv = add2(add2(scl2(vp.uv[0], bc[0]), scl2(vp.uv[1], bc[1])), scl2(vp.uv[2], bc[2]));
	if (hit_out) *hit_out = hit;
	return true;
}

struct PV { D2 dst, src; };

inline bool left_of(D2 A, D2 B, D2 P) { return cr2(sub2(B, A), sub2(P, A)) >= -1e-12; }  // :326-329
PV cut(const PV &S, const PV &E, D2 A, D2 B) {  // :331-346
	double dS = cr2(sub2(B, A), sub2(S.dst, A));
	double dE = cr2(sub2(B, A), sub2(E.dst, A));
	double t = dS / (dS - dE + 1e-30);
	t = clampd(t, 0.0, 1.0);
	PV I;
	I.dst = add2(S.dst, scl2(sub2(E.dst, S.dst), t));
	I.src = add2(S.src, scl2(sub2(E.src, S.src), t));
	return I;
}
std::vector<PV> clip_edge(D2 A, D2 B, const std::vector<PV> &in) {  // :360-376
	std::vector<PV> out;
	if (in.empty()) return out;
	PV S = in.back();
	bool s_in = left_of(A, B, S.dst);
	for (const PV &E : in) {
		bool e_in = left_of(A, B, E.dst);
		if (e_in) {
			if (!s_in) out.push_back(cut(S, E, A, B));
			out.push_back(E);
		} else if (s_in) out.push_back(cut(S, E, A, B));
		S = E;
		s_in = e_in;
	}
	return out;
}
std::vector<PV> clip_to_triangle(std::vector<PV> poly, D2 c0, D2 c1, D2 c2) {  // :348-382
	if (cr2(sub2(c1, c0), sub2(c2, c0)) < 0.0) std::swap(c1, c2);
	poly = clip_edge(c0, c1, poly);
	poly = clip_edge(c1, c2, poly);
	poly = clip_edge(c2, c0, poly);
	return poly;
}
double area_uv(const std::vector<PV> &poly) {  // :314-324
	double a = 0.0;
	for (size_t i = 0; i < poly.size(); ++i) {
		D2 p = poly[i].dst, q = poly[(i + 1) % poly.size()].dst;
		a += p.x * q.y - p.y * q.x;
	}
	return 0.5 * a;
}

void raster_triangle(Img &dst, D2 d0, D2 d1, D2 d2, const Img &src, D2 s0, D2 s1, D2 s2) {  // :384-438 (overwrite = true)
	D2 p0 = { d0.x * (dst.w - 1), d0.y * (dst.h - 1) }, p1 = { d1.x * (dst.w - 1), d1.y * (dst.h - 1) }, p2 = { d2.x * (dst.w - 1), d2.y * (dst.h - 1) };
	double minx = std::floor(std::min({ p0.x, p1.x, p2.x })), maxx = std::ceil(std::max({ p0.x, p1.x, p2.x }));
	double miny = std::floor(std::min({ p0.y, p1.y, p2.y })), maxy = std::ceil(std::max({ p0.y, p1.y, p2.y }));
	int x0 = clampi((int)minx, 0, dst.w - 1), x1 = clampi((int)maxx, 0, dst.w - 1);
	int y0 = clampi((int)miny, 0, dst.h - 1), y1 = clampi((int)maxy, 0, dst.h - 1);
	double area = cr2(sub2(p1, p0), sub2(p2, p0));
	if (std::fabs(area) < 1e-12) return;
	for (int y = y0; y <= y1; ++y)
		for (int x = x0; x <= x1; ++x) {
			D2 P = { (double)x + 0.5, (double)y + 0.5 };
			double w0 = cr2(sub2(p1, P), sub2(p2, P)) / area;
			double w1 = cr2(sub2(p2, P), sub2(p0, P)) / area;
			double w2 = 1.0 - w0 - w1;
			if (w0 < -1e-6 || w1 < -1e-6 || w2 < -1e-6) continue;
			D2 suv = add2(add2(scl2(s0, w0), scl2(s1, w1)), scl2(s2, w2));
			dst.put(x, y, fetch_bilinear(src, suv.x, suv.y));
		}
}
void raster_fan(Img &dst, const std::vector<PV> &poly, const Img &src) {  // :440-460
	if (poly.size() < 3) return;
	for (size_t i = 1; i + 1 < poly.size(); ++i)
		raster_triangle(dst, poly[0].dst, poly[i].dst, poly[i + 1].dst, src, poly[0].src, poly[i].src, poly[i + 1].src);
}

std::vector<const Tri *> collect(const std::vector<Tri> &scene, D3 origin, const Tri &vp) {  // :462-485
	std::vector<const Tri *> out;
	for (const Tri &t : scene) {
		bool any = false;
		for (int i = 0; i < 3 && !any; ++i) {
			D2 uv;
			D3 hit;
			if (!project_uv(origin, t.p[i], vp, uv, &hit)) continue;
			if (inside3(hit, vp, 1e-8)) any = true;
		}
		if (any) out.push_back(&t);
	}
	return out;
}
void sort_far_to_near(std::vector<const Tri *> &c, D3 origin) {  // :591-595, :698-702
	std::sort(c.begin(), c.end(), [&](const Tri *a, const Tri *b) {
		double da = len3(sub(tri_centroid(*a), origin));
		double db = len3(sub(tri_centroid(*b), origin));
		return da > db;
	});
}
// projection + clip of `t` into the viewport triangle `vp` seen from `origin`; false = nothing to draw
bool footprint(D3 origin, const Tri &t, const Tri &vp, std::vector<PV> &clipped) {
	std::vector<PV> poly(3);
	for (int i = 0; i < 3; ++i) {
		D2 uv;
		if (!project_uv(origin, t.p[i], vp, uv, nullptr)) return false;
		poly[i].dst = uv;
		poly[i].src = t.uv[i];
	}
	clipped = clip_to_triangle(poly, vp.uv[0], vp.uv[1], vp.uv[2]);
	return clipped.size() >= 3;
}

struct Counters { uint64_t nodes = 0, texels = 0, raster_px = 0; };

Img node_texture(const std::vector<Tri> &scene, D3 origin, const Tri &cur, int tw, int th, int depth, double est_area, std::vector<int> &stack,
	const Cfg &cfg, Counters &cnt) {  // :551-664
	tw = clampi(tw, cfg.min_res, cfg.max_res);
	th = clampi(th, cfg.min_res, cfg.max_res);
	D3 base = cur.mat ? cur.mat->albedo : D3{ 1, 1, 1 };
	Img solid(tw, th, base);
	if (!cur.mat || !cur.mat->reflective) return solid;
	if (depth >= cfg.max_depth) return solid;
	if (est_area > 0.0 && est_area < cfg.min_area) return solid;
	for (int id : stack)
		if (id == cur.id) return solid;
	stack.push_back(cur.id);
	++cnt.nodes;
	cnt.texels += (uint64_t)tw * th;

	D3 n = tri_normal(cur);
	D3 mirrored = sub(origin, scl(n, 2.0 * dot3(n, sub(origin, cur.p[0]))));  // :268-273
	Img refl(tw, th, cfg.env);
	std::vector<const Tri *> cand = collect(scene, mirrored, cur);
	sort_far_to_near(cand, mirrored);
	for (const Tri *t : cand) {
		if (t->id == cur.id) continue;
		std::vector<PV> clipped;
		if (!footprint(mirrored, *t, cur, clipped)) continue;
		double area_px = std::fabs(area_uv(clipped)) * (double)tw * (double)th;
		if (area_px < 0.5) continue;
		int res = clampi((int)std::lround(std::sqrt(area_px) * 1.2), cfg.min_res, cfg.max_res);
		Img child = node_texture(scene, mirrored, *t, res, res, depth + 1, area_px, stack, cfg, cnt);
		raster_fan(refl, clipped, child);
	}
	stack.pop_back();

	double m = clampd(cur.mat->metalness, 0.0, 1.0);
	Img out(tw, th, D3{ 0, 0, 0 });
	for (int y = 0; y < th; ++y)
		for (int x = 0; x < tw; ++x) {
			D3 r = refl.at(x, y);
			D3 tinted = { r.x * base.x, r.y * base.y, r.z * base.z };
			out.put(x, y, add(scl(base, 1.0 - m), scl(tinted, m)));
		}
	return out;
}

bool inside2(D2 P, D2 A, D2 B, D2 C, double eps) {  // :202-218
	double c0 = cr2(sub2(B, A), sub2(P, A)), c1 = cr2(sub2(C, B), sub2(P, B)), c2 = cr2(sub2(A, C), sub2(P, C));
	bool neg = (c0 < -eps) || (c1 < -eps) || (c2 < -eps);
	bool pos = (c0 > eps) || (c1 > eps) || (c2 > eps);
	return !(neg && pos);
}

Img viewport_image(const std::vector<Tri> &scene, D3 origin, const Tri &vp, int W, int H, const Cfg &cfg, Counters &cnt) {  // :683-753
	Img img(W, H, cfg.env);
	std::vector<const Tri *> cand = collect(scene, origin, vp);
	sort_far_to_near(cand, origin);
	std::vector<int> stack;
	for (const Tri *t : cand) {
		std::vector<PV> clipped;
		if (!footprint(origin, *t, vp, clipped)) continue;
		double area_px = std::fabs(area_uv(clipped)) * (double)W * (double)H;
		if (area_px < 0.5) continue;
		int res = clampi((int)std::lround(std::sqrt(area_px) * 1.0), cfg.min_res, cfg.max_res);
		Img tex = node_texture(scene, origin, *t, res, res, 0, area_px, stack, cfg, cnt);
		raster_fan(img, clipped, tex);
	}
	for (int y = 0; y < H; ++y)
		for (int x = 0; x < W; ++x) {
			D2 uv = { (double)x / (W - 1), (double)y / (H - 1) };
			if (!inside2(uv, vp.uv[0], vp.uv[1], vp.uv[2], 1e-10)) img.put(x, y, D3{ 0, 0, 0 });
		}
	return img;
}

struct Scene {
	std::vector<Mat> mats;
	std::vector<Tri> tris;
};
void load_scene(Scene &s, int n_tri, const double *P, const double *UV, const int *material, int n_mat, const int *mat_type, const double *mat_albedo,
	const double *mat_metalness) {
	s.mats.resize(n_mat);
	for (int i = 0; i < n_mat; ++i) s.mats[i] = Mat{ mat_type[i] ? 1 : 0, D3{ mat_albedo[3 * i], mat_albedo[3 * i + 1], mat_albedo[3 * i + 2] }, mat_metalness[i] };
	s.tris.resize(n_tri);
	for (int i = 0; i < n_tri; ++i) {
		Tri &t = s.tris[i];
		t.id = i + 1;
		for (int k = 0; k < 3; ++k) {
			t.p[k] = D3{ P[9 * i + 3 * k], P[9 * i + 3 * k + 1], P[9 * i + 3 * k + 2] };
			t.uv[k] = D2{ UV[6 * i + 2 * k], UV[6 * i + 2 * k + 1] };
		}
		t.mat = (material[i] >= 0 && material[i] < n_mat) ? &s.mats[material[i]] : nullptr;
	}
}
Tri load_viewport(int id, const double *P, const double *UV) {
	Tri t;
	t.id = id;
	for (int k = 0; k < 3; ++k) {
		t.p[k] = D3{ P[3 * k], P[3 * k + 1], P[3 * k + 2] };
		t.uv[k] = D2{ UV[2 * k], UV[2 * k + 1] };
	}
	t.mat = nullptr;
	return t;
}
Cfg load_cfg(const double *c) { return Cfg{ (int)c[0], c[1], (int)c[2], (int)c[3], D3{ c[4], c[5], c[6] } }; }

inline uint8_t encode8(double c, double gamma) {  // :124-141
	c = clampd(c, 0.0, 1.0);
	c = std::pow(c, 1.0 / gamma);
	return (uint8_t)clampi((int)std::lround(c * 255.0), 0, 255);
}

}  // namespace

extern "C" {

// cfg = { maxDepth, minAreaPxToRecurse, maxTriTexRes, minTriTexRes, env.r, env.g, env.b, gamma }
// counters (may be NULL) = { reflective nodes rendered, node texels, 0 }
int orc_patch_render(int n_tri, const double *P, const double *UV, const int *material, int n_mat, const int *mat_type, const double *mat_albedo,
	const double *mat_metalness, const double *origin, const double *vp_P, const double *vp_UV, int width, int height, const double *cfg,
	double *out_rgb, uint8_t *out_rgb8, uint64_t *counters) {
	Scene s;
	load_scene(s, n_tri, P, UV, material, n_mat, mat_type, mat_albedo, mat_metalness);
	const Cfg c = load_cfg(cfg);
	const D3 o = { origin[0], origin[1], origin[2] };
	Counters cnt;
	Img a = viewport_image(s.tris, o, load_viewport(-100, vp_P, vp_UV), width, height, c, cnt);
	Img b = viewport_image(s.tris, o, load_viewport(-101, vp_P + 9, vp_UV + 6), width, height, c, cnt);
	for (int y = 0; y < height; ++y)  // :755-772
		for (int x = 0; x < width; ++x) {
			D3 v = add(a.at(x, y), b.at(x, y));
			v = D3{ clampd(v.x, 0.0, 1.0), clampd(v.y, 0.0, 1.0), clampd(v.z, 0.0, 1.0) };
			size_t i = (size_t)y * width + x;
			if (out_rgb) { out_rgb[3 * i] = v.x; out_rgb[3 * i + 1] = v.y; out_rgb[3 * i + 2] = v.z; }
			if (out_rgb8) { out_rgb8[3 * i] = encode8(v.x, cfg[7]); out_rgb8[3 * i + 1] = encode8(v.y, cfg[7]); out_rgb8[3 * i + 2] = encode8(v.z, cfg[7]); }
		}
	if (counters) { counters[0] = cnt.nodes; counters[1] = cnt.texels; counters[2] = 0; }
	return 0;
}

int orc_patch_trace_texture(int n_tri, const double *P, const double *UV, const int *material, int n_mat, const int *mat_type, const double *mat_albedo,
	const double *mat_metalness, const double *origin, int current, int tex_w, int tex_h, double est_area_px, const double *cfg, double *out_tex,
	int *out_wh) {
	Scene s;
	load_scene(s, n_tri, P, UV, material, n_mat, mat_type, mat_albedo, mat_metalness);
	if (current < 0 || current >= n_tri) return -1;
	std::vector<int> stack;
	Counters cnt;
	Img t = node_texture(s.tris, D3{ origin[0], origin[1], origin[2] }, s.tris[current], tex_w, tex_h, 0, est_area_px, stack, load_cfg(cfg), cnt);
	out_wh[0] = t.w;
	out_wh[1] = t.h;
	for (size_t i = 0; i < t.px.size(); ++i) { out_tex[3 * i] = t.px[i].x; out_tex[3 * i + 1] = t.px[i].y; out_tex[3 * i + 2] = t.px[i].z; }
	return 0;
}

}  // extern "C"
