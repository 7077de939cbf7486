/* oracle/are_oracle.c — CPU restatement of the reference arithmetic on the path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under aurora_rendering_engine_b200/ or include/ may include, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * do, and only as the checker.  The product has no CPU path.
 *
 * Two layers:
 *   (1) "lib_*"  — the reference LIBRARY's arithmetic, restated in plain C fp64, each function citing the
 *       /root/reference file:line it follows.  PINNED: tests/test_oracle_vs_reference.py checks every one of
 *       these against the real reference compiled into oracle/_ref/libare_ref.so (oracle/ref_harness.cpp) and
 *       against the golden vectors committed in tests/golden/ (generated from that library by
 *       tests/golden/make_golden.py), including the known answers recorded in SURVEY.md §8c.
 *   (2) "orc_*"  — the fp64 CPU twin of the NEW surface the north_star asks for and the reference does not
 *       contain (sphere, quad, lambertian / metal / dielectric / light scatter, procedural textures, camera
 *       jitter + lens, Philox sampling, spp/bounce loop).  PARITY UNPINNED by any reference test for those
 *       parts (no reference implementation exists); every piece that CAN be tied to reference
 *       arithmetic is: triangle hit = Triangle::intersect_ray, quad plane step = Plane::intersect_ray,
 *       metal = are::reflect, dielectric = are::refract, lambertian sampler = rt.cpp cosine gather,
 *       camera = rt.cpp ray set-up, uv checker = rt.cpp CheckerTexture, Philox = Random123 known answers.
 *   (3) "rtao_*" — restatement of experiments/rt.cpp's own shading loop (fp32 like the original) driven by
 *       Philox instead of mt19937(random_device); pinned statistically against oracle/_ref/rt_ref images.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define LIB_EPS 1e-12 /* are::GEOMETRY_EPSILON, include/basic/math.h:7 */

typedef struct { double x, y, z; } v3;

/* ===================================================================================================== */
/* (1) library arithmetic                                                                                  */
/* ===================================================================================================== */

static inline v3 V(double x, double y, double z) { v3 r = { x, y, z }; return r; }
static inline v3 vnan(void) { return V(NAN, NAN, NAN); }
static inline v3 vld(const double *p) { return V(p[0], p[1], p[2]); }
static inline void vst(double *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
/* src/basic/vec3.cpp:149-171 */
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(double t, v3 a) { return V(t * a.x, t * a.y, t * a.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* src/basic/vec3.cpp:174-179: division multiplies by the reciprocal; /0 -> NaN vector */
static inline v3 vdiv(v3 a, double t) { if (t == 0.0) return vnan(); return vscale(1 / t, a); }
/* src/basic/vec3.cpp:132-134 */
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* src/basic/vec3.cpp:137-140 */
static inline v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* src/basic/vec3.cpp:102-109 */
static inline double vlen2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline double vlen(v3 a) { return sqrt(vlen2(a)); }
/* src/basic/vec3.cpp:123-129: normalized() = *this / len, NaN vector when len == 0 */
static inline v3 vnormalized(v3 a) { double l = vlen(a); if (l == 0.0) return vnan(); return vdiv(a, l); }
/* src/basic/vec3.cpp:143-146 */
static inline int vnear_zero(v3 a) { const double s = 1e-8; return fabs(a.x) < s && fabs(a.y) < s && fabs(a.z) < s; }
/* src/basic/vec3.cpp:182-184:  v - 2*v.dot(n)*n  (evaluated as (2*dot)*n) */
static inline v3 lib_reflect1(v3 v, v3 n) { return vsub(v, vscale(2 * vdot(v, n), n)); }
/* src/basic/vec3.cpp:188-199: fabs() under the sqrt means total internal reflection is NOT detected */
static inline v3 lib_refract1(v3 uv, v3 n, double eta) {
	double cos_theta = fmin(vdot(vneg(uv), n), 1.0);
	v3 perp = vscale(eta, vadd(uv, vscale(cos_theta, n)));
	v3 par = vscale(-sqrt(fabs(1.0 - vlen2(perp))), n);
	double l = vlen(par);
	if (isnan(l) || isinf(l)) return vnan();
	return vadd(perp, par);
}

/* op codes mirror oracle/ref_harness.cpp ref_vec3_binary / ref_vec3_scalar */
void lib_vec3_binary(int op, int n, const double *a, const double *b, const double *s, double *out) {
	for (int i = 0; i < n; ++i) {
		v3 A = vld(a + 3 * i), B = b ? vld(b + 3 * i) : V(0, 0, 0), r;
		double k = s ? s[i] : 0.0;
		switch (op) {
		case 0: r = vadd(A, B); break;
		case 1: r = vsub(A, B); break;
		case 2: r = vmul(A, B); break;
		case 3: r = vcross(A, B); break;
		case 4: r = vscale(k, A); break;
		case 5: r = vdiv(A, k); break;
		case 6: r = vnormalized(A); break;
		default: r = vneg(A); break;
		}
		vst(out + 3 * i, r);
	}
}
void lib_vec3_scalar(int op, int n, const double *a, const double *b, double *out) {
	for (int i = 0; i < n; ++i) {
		v3 A = vld(a + 3 * i), B = b ? vld(b + 3 * i) : V(0, 0, 0);
		switch (op) {
		case 0: out[i] = vdot(A, B); break;
		case 1: out[i] = vlen(A); break;
		case 2: out[i] = vlen2(A); break;
		default: out[i] = vnear_zero(A) ? 1.0 : 0.0; break;
		}
	}
}
void lib_reflect(int n, const double *v, const double *nrm, double *out) {
	for (int i = 0; i < n; ++i) vst(out + 3 * i, lib_reflect1(vld(v + 3 * i), vld(nrm + 3 * i)));
}
void lib_refract(int n, const double *uv, const double *nrm, const double *eta, double *out) {
	for (int i = 0; i < n; ++i) vst(out + 3 * i, lib_refract1(vld(uv + 3 * i), vld(nrm + 3 * i), eta[i]));
}
/* src/basic/ray.cpp:5-10 */
void lib_ray(int n, const double *Q, const double *D, const double *t, double *outD, double *outAt) {
	for (int i = 0; i < n; ++i) {
		v3 d = vnormalized(vld(D + 3 * i));
		vst(outD + 3 * i, d);
		if (outAt) vst(outAt + 3 * i, vadd(vld(Q + 3 * i), vscale(t ? t[i] : 0.0, vnormalized(d))));
	}
}
/* src/basic/plane.cpp:10-11 */
void lib_plane_from_point_normal(int n, const double *p, const double *nrm, double *out4) {
	for (int i = 0; i < n; ++i) {
		v3 nn = vnormalized(vld(nrm + 3 * i));
		vst(out4 + 4 * i, nn);
		out4[4 * i + 3] = -vdot(nn, vld(p + 3 * i));
	}
}
/* src/basic/plane.cpp:13-27 */
static inline int lib_plane_hit1(v3 pn, double pd, v3 Q, v3 D, v3 *X) {
	double denom = vdot(pn, D);
	if (fabs(denom) < LIB_EPS) return 0;
	double t = -(vdot(pn, Q) + pd) / denom;
	if (t < LIB_EPS) return 0;
	*X = vadd(Q, vscale(t, D));
	return 1;
}
void lib_plane_intersect(int n, const double *plane4, const double *Q, const double *D, int *hit, double *P) {
	for (int i = 0; i < n; ++i) {
		v3 x = vnan();
		hit[i] = lib_plane_hit1(vld(plane4 + 4 * i), plane4[4 * i + 3], vld(Q + 3 * i), vnormalized(vld(D + 3 * i)), &x);
		vst(P + 3 * i, x);
	}
}
/* src/material/diffuse.cpp:5-7 (kind 0) and src/material/reflective.cpp:9-26 (kind 1) */
void lib_material_reflect(int kind, double reflectivity, int n, const double *plane4, const double *origin, int *ok, double *out) {
	(void)reflectivity;
	for (int i = 0; i < n; ++i) {
		v3 o = vnan();
		ok[i] = 0;
		if (kind == 1) {
			v3 pn = vld(plane4 + 4 * i);
			double denom = vlen2(pn);
			if (!(denom < LIB_EPS)) {
				double numer = vdot(pn, vld(origin + 3 * i)) + plane4[4 * i + 3];
				double k = 2.0 * numer / denom;
				o = vsub(vld(origin + 3 * i), vscale(k, pn));
				ok[i] = 1;
			}
		}
		vst(out + 3 * i, o);
	}
}
/* src/object/triangle.cpp:9-46. 0 ok, 1 invalid_argument. flags bit0/bit1: null material / texture */
int lib_triangle_ctor(const double *Q, const double *u, const double *v, int flags, double *verts9) {
	if (flags & 3) return 1;
	v3 q = vld(Q), uu = vld(u), vv = vld(v);
	if (vnear_zero(uu) || vnear_zero(vv) || vnear_zero(vcross(uu, vv))) return 1;
	if (verts9) { vst(verts9, q); vst(verts9 + 3, vadd(q, uu)); vst(verts9 + 6, vadd(q, vv)); }
	return 0;
}
/* src/object/triangle.cpp:82-121 — Möller–Trumbore, two-sided, ±eps slack, returns the hit POINT */
static inline int lib_tri_hit1(v3 Q, v3 u, v3 v, v3 rQ, v3 rD, v3 *X, double *tt, double *al, double *be) {
	v3 h = vcross(rD, v);
	double a = vdot(u, h);
	if (fabs(a) < LIB_EPS) return 0;
	double f = 1.0 / a;
	v3 s = vsub(rQ, Q);
	double alpha = f * vdot(s, h);
	if (alpha < -LIB_EPS || alpha > 1.0 + LIB_EPS) return 0;
	v3 q = vcross(s, u);
	double beta = f * vdot(rD, q);
	if (beta < -LIB_EPS || alpha + beta > 1.0 + LIB_EPS) return 0;
	double t = f * vdot(v, q);
	if (t > LIB_EPS) {
		*X = vadd(rQ, vscale(t, rD));
		if (tt) *tt = t;
		if (al) *al = alpha;
		if (be) *be = beta;
		return 1;
	}
	return 0;
}
/* src/object/triangle.cpp:49-79 */
static inline int lib_tri_point_in1(v3 Q, v3 u, v3 v, v3 p) {
	v3 v0 = v, v1 = u, v2 = vsub(p, Q);
	double d00 = vdot(v0, v0), d01 = vdot(v0, v1), d02 = vdot(v0, v2), d11 = vdot(v1, v1), d12 = vdot(v1, v2);
	double inv = d00 * d11 - d01 * d01;
	if (fabs(inv) < LIB_EPS) return 0;
	inv = 1.0 / inv;
	double alpha = (d11 * d02 - d01 * d12) * inv, beta = (d00 * d12 - d01 * d02) * inv;
	return alpha >= -LIB_EPS && beta >= -LIB_EPS && alpha + beta <= 1.0 + LIB_EPS;
}
void lib_triset_hit_matrix(int ntri, const double *TQ, const double *Tu, const double *Tv, int nrays,
	const double *Q, const double *D, int *hit, double *P) {
	for (int r = 0; r < nrays; ++r) {
		v3 rq = vld(Q + 3 * r), rd = vnormalized(vld(D + 3 * r));
		for (int k = 0; k < ntri; ++k) {
			v3 x = vnan();
			hit[(size_t)r * ntri + k] = lib_tri_hit1(vld(TQ + 3 * k), vld(Tu + 3 * k), vld(Tv + 3 * k), rq, rd, &x, 0, 0, 0);
			vst(P + ((size_t)r * ntri + k) * 3, x);
		}
	}
}
/* the renderer's drive of the list: t recovered as (P-Q)·D, min t, ties keep the lower index */
long lib_triset_closest_hit(int ntri, const double *TQ, const double *Tu, const double *Tv, int nrays,
	const double *Q, const double *D, int *prim, double *t, double *P) {
	long nhit = 0;
	for (int r = 0; r < nrays; ++r) {
		v3 rq = vld(Q + 3 * r), rd = vnormalized(vld(D + 3 * r)), bp = vnan();
		int best = -1;
		double bt = INFINITY;
		for (int k = 0; k < ntri; ++k) {
			v3 x;
			if (lib_tri_hit1(vld(TQ + 3 * k), vld(Tu + 3 * k), vld(Tv + 3 * k), rq, rd, &x, 0, 0, 0)) {
				double tk = vdot(vsub(x, rq), rd);
				if (tk < bt) { bt = tk; best = k; bp = x; }
			}
		}
		prim[r] = best;
		t[r] = best >= 0 ? bt : NAN;
		vst(P + 3 * r, bp);
		nhit += best >= 0;
	}
	return nhit;
}
void lib_triset_point_in(const double *TQ, const double *Tu, const double *Tv, int tri, int n, const double *pts, int *inside) {
	for (int i = 0; i < n; ++i)
		inside[i] = lib_tri_point_in1(vld(TQ + 3 * tri), vld(Tu + 3 * tri), vld(Tv + 3 * tri), vld(pts + 3 * i));
}

/* src/texture.cpp:9-50 — binary P6, maxval 255 only, channel = byte / 255.0.  0 ok, 1 runtime_error, 3 cap */
int lib_texture_load(const char *path, int *w, int *h, double *rgb, long cap) {
	FILE *fp = fopen(path, "rb");
	if (!fp) return 1;
	char fmt[3] = { 0 };
	int W, H, mx;
	if (fscanf(fp, "%2s", fmt) != 1 || strcmp(fmt, "P6") != 0) { fclose(fp); return 1; }
	if (fscanf(fp, "%d %d %d", &W, &H, &mx) != 3 || mx != 255) { fclose(fp); return 1; }
	fgetc(fp);
	*w = W; *h = H;
	if (rgb && (long)W * H * 3 > cap) { fclose(fp); return 3; }
	for (long i = 0; i < (long)W * H; ++i) {
		unsigned char c[3];
		if (fread(c, 1, 3, fp) != 3) { fclose(fp); return 1; }
		if (rgb) { rgb[3 * i] = c[0] / 255.0; rgb[3 * i + 1] = c[1] / 255.0; rgb[3 * i + 2] = c[2] / 255.0; }
	}
	fclose(fp);
	return 0;
}
/* src/texture.cpp:383-386: static_cast<unsigned char>(clamp(c*255, 0, 255)) — truncation, no gamma */
static inline unsigned char lib_quant_linear(double c) {
	double x = c * 255.0;
	x = x < 0.0 ? 0.0 : (x > 255.0 ? 255.0 : x); /* std::clamp: NaN compares false twice and passes through */
	return (unsigned char)x;
}
void lib_encode_linear(long n, const double *c, unsigned char *out) {
	for (long i = 0; i < n; ++i) out[i] = lib_quant_linear(c[i]);
}
/* src/texture.cpp:362-395. 1 saved, 0 refused (.ppm suffix / open failure) */
int lib_texture_save(const char *path, int w, int h, const double *rgb) {
	size_t L = strlen(path);
	if (L < 4 || strcmp(path + L - 4, ".ppm") != 0) return 0;
	FILE *fp = fopen(path, "wb");
	if (!fp) return 0;
	fprintf(fp, "P6\n%d %d\n255\n", w, h);
	for (long i = 0; i < (long)w * h * 3; ++i) { unsigned char b = lib_quant_linear(rgb[i]); fwrite(&b, 1, 1, fp); }
	fclose(fp);
	return 1;
}
/* src/texture.cpp:85-360 — Texture::paste: warp src into the quad (lt, rt, lb, rb) of dst through the 4-corner
 * homography (8x8 Gauss-Jordan with partial pivoting, 3x3 inverse by cofactors), even-odd point-in-quad at pixel
 * centres, bilinear fetch.  dst/src = w*h*3 doubles, corners = {lt.x, lt.y, rt.x, rt.y, lb.x, lb.y, rb.x, rb.y}. */
void lib_texture_paste(double *dst, int dw, int dh, const double *src, int sw_i, int sh_i, const int *corners) {
	if (!dst || sw_i <= 0 || sh_i <= 0) return;
	const double lt[2] = { corners[0], corners[1] }, rt[2] = { corners[2], corners[3] }, lb[2] = { corners[4], corners[5] }, rb[2] = { corners[6], corners[7] };
	const double qx[4] = { lt[0], rt[0], rb[0], lb[0] }, qy[4] = { lt[1], rt[1], rb[1], lb[1] };
	double min_x = fmin(fmin(qx[0], qx[1]), fmin(qx[2], qx[3])), max_x = fmax(fmax(qx[0], qx[1]), fmax(qx[2], qx[3]));
	double min_y = fmin(fmin(qy[0], qy[1]), fmin(qy[2], qy[3])), max_y = fmax(fmax(qy[0], qy[1]), fmax(qy[2], qy[3]));
	if (max_x < 0.0 || max_y < 0.0 || min_x > (double)(dw - 1) || min_y > (double)(dh - 1)) return;
	int x0 = (int)floor(min_x), x1 = (int)ceil(max_x), y0 = (int)floor(min_y), y1 = (int)ceil(max_y);
	if (x0 < 0) x0 = 0;
	if (x1 > dw - 1) x1 = dw - 1;
	if (y0 < 0) y0 = 0;
	if (y1 > dh - 1) y1 = dh - 1;
	const double sw = sw_i, sh = sh_i;
	const double sxy[4][2] = { { 0.0, 0.0 }, { sw - 1.0, 0.0 }, { 0.0, sh - 1.0 }, { sw - 1.0, sh - 1.0 } };
	const double dxy[4][2] = { { lt[0], lt[1] }, { rt[0], rt[1] }, { lb[0], lb[1] }, { rb[0], rb[1] } };
	double M[8][9];
	for (int k = 0; k < 4; ++k) {
		double x = sxy[k][0], y = sxy[k][1], u = dxy[k][0], v = dxy[k][1];
		double r0[9] = { x, y, 1.0, 0.0, 0.0, 0.0, -u * x, -u * y, u }, r1[9] = { 0.0, 0.0, 0.0, x, y, 1.0, -v * x, -v * y, v };
		memcpy(M[2 * k], r0, sizeof r0);
		memcpy(M[2 * k + 1], r1, sizeof r1);
	}
	for (int col = 0; col < 8; ++col) {
		int pivot = col;
		double best = fabs(M[col][col]);
		for (int r = col + 1; r < 8; ++r)
			if (fabs(M[r][col]) > best) { best = fabs(M[r][col]); pivot = r; }
		if (best < LIB_EPS) return;
		if (pivot != col)
			for (int c = col; c < 9; ++c) { double t = M[col][c]; M[col][c] = M[pivot][c]; M[pivot][c] = t; }
		double div = M[col][col];
		for (int c = col; c < 9; ++c) M[col][c] /= div;
		for (int r = 0; r < 8; ++r) {
			if (r == col) continue;
			double factor = M[r][col];
			if (fabs(factor) < LIB_EPS) continue;
			for (int c = col; c < 9; ++c) M[r][c] -= factor * M[col][c];
		}
	}
	const double a = M[0][8], b = M[1][8], c = M[2][8], d = M[3][8], e = M[4][8], f = M[5][8], g = M[6][8], h = M[7][8], i = 1.0;
	const double A = (e * i - f * h), B = -(d * i - f * g), C = (d * h - e * g), D = -(b * i - c * h), E = (a * i - c * g), F = -(a * h - b * g);
	const double G = (b * f - c * e), Hc = -(a * f - c * d), I = (a * e - b * d);
	const double det = a * A + b * B + c * C;
	if (fabs(det) < LIB_EPS) return;
	const double id = 1.0 / det;
	const double hi[9] = { A * id, D * id, G * id, B * id, E * id, Hc * id, C * id, F * id, I * id };
	for (int y = y0; y <= y1; ++y)
		for (int x = x0; x <= x1; ++x) {
			const double X = (double)x + 0.5, Y = (double)y + 0.5;
			int inside = 0;
			for (int p = 0, q = 3; p < 4; q = p++) {
				const double xi = qx[p], yi = qy[p], xj = qx[q], yj = qy[q];
				const int cross = ((yi > Y) != (yj > Y)) && (X < (xj - xi) * (Y - yi) / ((yj - yi) == 0.0 ? 1e-30 : (yj - yi)) + xi);
				inside ^= cross;
			}
			if (!inside) continue;
			const double denom = hi[6] * X + hi[7] * Y + hi[8];
			if (fabs(denom) < LIB_EPS) continue;
			double sx = (hi[0] * X + hi[1] * Y + hi[2]) / denom, sy = (hi[3] * X + hi[4] * Y + hi[5]) / denom;
			if (sx < 0.0 || sy < 0.0 || sx > sw - 1.0 || sy > sh - 1.0) continue;
			sx = sx < 0.0 ? 0.0 : (sw - 1.0 < sx ? sw - 1.0 : sx);
			sy = sy < 0.0 ? 0.0 : (sh - 1.0 < sy ? sh - 1.0 : sy);
			int ix = (int)floor(sx), iy = (int)floor(sy);
			int ix1 = ix + 1 < sw_i - 1 ? ix + 1 : sw_i - 1, iy1 = iy + 1 < sh_i - 1 ? iy + 1 : sh_i - 1;
			double tx = sx - (double)ix, ty = sy - (double)iy;
			double w00 = (1.0 - tx) * (1.0 - ty), w10 = tx * (1.0 - ty), w01 = (1.0 - tx) * ty, w11 = tx * ty;
			const double *c00 = src + ((size_t)iy * sw_i + ix) * 3, *c10 = src + ((size_t)iy * sw_i + ix1) * 3;
			const double *c01 = src + ((size_t)iy1 * sw_i + ix) * 3, *c11 = src + ((size_t)iy1 * sw_i + ix1) * 3;
			double *out = dst + ((size_t)y * dw + x) * 3;
			for (int k = 0; k < 3; ++k) out[k] = w00 * c00[k] + w10 * c10[k] + w01 * c01[k] + w11 * c11[k];
		}
}
/* experiments/rt.cpp:72-76,383-386: clamp to [0,1] (fmax/fmin), (unsigned char)(powf(c, 1/2.2f) * 255) in fp32 */
void lib_encode_gamma22(long n, const float *c, unsigned char *out) {
	for (long i = 0; i < n; ++i) {
		float x = fmaxf(0.0f, fminf(1.0f, c[i]));
		out[i] = (unsigned char)(powf(x, 1 / 2.2f) * 255);
	}
}
/* gamma-2 encoder (NEW, RTIOW convention): int(256 * clamp(sqrt(c), 0, 0.999)) */
void lib_encode_sqrt(long n, const float *c, unsigned char *out) {
	for (long i = 0; i < n; ++i) {
		float x = c[i] > 0.0f ? sqrtf(c[i]) : 0.0f;
		x = x < 0.0f ? 0.0f : (x > 0.999f ? 0.999f : x);
		out[i] = (unsigned char)(256.0f * x);
	}
}

/* ===================================================================================================== */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 constants).  Pinned by the Random123 known-answer  */
/* vectors in tests/golden/philox_kat.json.                                                                */
/* ===================================================================================================== */
static inline void philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t o[4]) {
	uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
	for (int r = 0; r < 10; ++r) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
void orc_philox(int n, uint64_t seed, const uint32_t *ctr, uint32_t *out) {
	for (int i = 0; i < n; ++i) philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), ctr + 4 * i, out + 4 * i);
}
/* 24-bit uniforms in [0,1): (x >> 8) * 2^-24 — exact in fp32 and fp64 alike */
static inline void rnd4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t stream, double r[4]) {
	uint32_t c[4] = { pixel, sample, bounce, stream }, o[4];
	philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c, o);
	for (int i = 0; i < 4; ++i) r[i] = (double)(o[i] >> 8) * (1.0 / 16777216.0);
}

/* ===================================================================================================== */
/* (2) fp64 twin of the path-tracing surface                                                               */
/* ===================================================================================================== */
enum { MAT_DIFFUSE = 0, MAT_REFLECTIVE = 1, MAT_LAMBERTIAN = 2, MAT_METAL = 3, MAT_DIELECTRIC = 4, MAT_LIGHT = 5 };
enum { TEX_SOLID = 0, TEX_CHECKER_UV = 1, TEX_CHECKER_3D = 2, TEX_NOISE = 3, TEX_IMAGE = 4 };
enum { PRIM_TRIANGLE = 0, PRIM_QUAD = 1, PRIM_SPHERE = 2 };

typedef struct {
	int kind;
	double p[8];
	int w, h;
	double *rgb; /* IMAGE */
	v3 *ranvec; /* NOISE: 256 unit vectors */
	int *perm; /* NOISE: 3 x 256 */
} orc_texture;
typedef struct { int kind; double p[8]; } orc_material;
typedef struct {
	int type, mat, tex;
	v3 Q, u, v; /* sphere: Q = centre, u.x = radius */
	double uv[6];
} orc_prim;
/* CPU-side acceleration (NEW, test infrastructure): median-split AABB tree over the primitives, built on first use for
 * scenes of more than ORC_BVH_MIN primitives.  It only prunes: every surviving primitive goes through the same
 * tri_test / quad_test / sphere_test as the linear scan, and ties in t go to the lower primitive index exactly as the
 * scan resolves them, so closest_hit returns the same (primitive, t, a, b) with or without it. */
typedef struct { double lo[3], hi[3]; int left, right, first, count; } orc_node; /* leaf: count > 0 */
typedef struct {
	orc_texture *tex; int ntex, ctex;
	orc_material *mat; int nmat, cmat;
	orc_prim *prim; int nprim, cprim;
	orc_node *node; int nnode, cnode; int *order; double *pb; int bvh_nprim; /* tree valid for the first bvh_nprim primitives; pb = 6 bounds per primitive */
} orc_scene;

/* identical field order to are_camera / are_render_params in include/are_cuda.h */
typedef struct {
	double pos[3], target[3], up[3];
	double vfov_deg, focus_dist, defocus_angle_deg;
	int32_t jitter, pad_;
} orc_camera;
typedef struct {
	int32_t width, height, sample_begin, sample_count, max_depth, integrator, traversal, ao_samples;
	uint64_t seed;
	double t_min;
	double background_bottom[3], background_top[3];
} orc_params;
typedef struct {
	uint64_t samples, rays, tri_tests, quad_tests, sphere_tests, node_visits, box_tests;
	double kernel_ms;
	uint64_t launches;
} orc_stats;

orc_scene *orc_scene_create(void) { return (orc_scene *)calloc(1, sizeof(orc_scene)); }
void orc_scene_destroy(orc_scene *s) {
	if (!s) return;
	for (int i = 0; i < s->ntex; ++i) { free(s->tex[i].rgb); free(s->tex[i].ranvec); free(s->tex[i].perm); }
	free(s->tex); free(s->mat); free(s->prim); free(s->node); free(s->order); free(s->pb); free(s);
}
#define GROW(arr, n, c, T) do { if ((n) == (c)) { (c) = (c) ? 2 * (c) : 16; (arr) = (T *)realloc((arr), (size_t)(c) * sizeof(T)); } } while (0)

/* Perlin tables (NEW): 256 unit gradient vectors + three Fisher–Yates permutations, all drawn from
 * Philox(key = table seed) with counter (i, 0, 0, 0x5045524C).  Spec shared with the CUDA library. */
static void noise_tables(uint64_t seed, v3 *ranvec, int *perm) {
	for (int i = 0; i < 256; ++i) {
		double r[4];
		rnd4(seed, (uint32_t)i, 0, 0, 0x5045524Cu, r);
		double z = 1.0 - 2.0 * r[0], rxy = sqrt(fmax(0.0, 1.0 - z * z)), phi = 2.0 * M_PI * r[1];
		ranvec[i] = V(rxy * cos(phi), rxy * sin(phi), z);
	}
	for (int a = 0; a < 3; ++a) {
		int *p = perm + 256 * a;
		for (int i = 0; i < 256; ++i) p[i] = i;
		for (int i = 255; i > 0; --i) {
			uint32_t c[4] = { (uint32_t)i, (uint32_t)(a + 1), 0, 0x5045524Cu }, o[4];
			philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c, o);
			int target = (int)(o[0] % (uint32_t)(i + 1));
			int tmp = p[i]; p[i] = p[target]; p[target] = tmp;
		}
	}
}
int orc_add_texture(orc_scene *s, int kind, const double params[8], const double *rgb, int w, int h) {
	GROW(s->tex, s->ntex, s->ctex, orc_texture);
	orc_texture *t = &s->tex[s->ntex];
	memset(t, 0, sizeof *t);
	t->kind = kind;
	if (params) memcpy(t->p, params, 8 * sizeof(double));
	if (kind == TEX_IMAGE) {
		t->w = w; t->h = h;
		t->rgb = (double *)malloc((size_t)w * h * 3 * sizeof(double));
		memcpy(t->rgb, rgb, (size_t)w * h * 3 * sizeof(double));
	} else if (kind == TEX_NOISE) {
		t->ranvec = (v3 *)malloc(256 * sizeof(v3));
		t->perm = (int *)malloc(3 * 256 * sizeof(int));
		noise_tables((uint64_t)params[1], t->ranvec, t->perm);
	}
	return s->ntex++;
}
int orc_add_material(orc_scene *s, int kind, const double params[8]) {
	GROW(s->mat, s->nmat, s->cmat, orc_material);
	s->mat[s->nmat].kind = kind;
	memset(s->mat[s->nmat].p, 0, sizeof s->mat[s->nmat].p);
	if (params) memcpy(s->mat[s->nmat].p, params, 8 * sizeof(double));
	return s->nmat++;
}
static int add_prim(orc_scene *s, int type, v3 Q, v3 u, v3 v, int mat, int tex) {
	GROW(s->prim, s->nprim, s->cprim, orc_prim);
	orc_prim *p = &s->prim[s->nprim];
	p->type = type; p->mat = mat; p->tex = tex; p->Q = Q; p->u = u; p->v = v;
	double duv[6] = { 0, 0, 1, 0, 0, 1 };
	memcpy(p->uv, duv, sizeof duv);
	return s->nprim++;
}
int orc_add_triangle(orc_scene *s, const double Q[3], const double u[3], const double v[3], int mat, int tex) {
	if (lib_triangle_ctor(Q, u, v, 0, 0)) return -1;
	return add_prim(s, PRIM_TRIANGLE, vld(Q), vld(u), vld(v), mat, tex);
}
int orc_set_triangle_uv(orc_scene *s, int prim, const double uv[6]) {
	if (prim < 0 || prim >= s->nprim || s->prim[prim].type != PRIM_TRIANGLE) return -1;
	memcpy(s->prim[prim].uv, uv, 6 * sizeof(double));
	return 0;
}
int orc_add_quad(orc_scene *s, const double Q[3], const double u[3], const double v[3], int mat, int tex) {
	if (lib_triangle_ctor(Q, u, v, 0, 0)) return -1;
	return add_prim(s, PRIM_QUAD, vld(Q), vld(u), vld(v), mat, tex);
}
int orc_add_sphere(orc_scene *s, const double c[3], double radius, int mat, int tex) {
	if (!(radius > 0.0)) return -1;
	return add_prim(s, PRIM_SPHERE, vld(c), V(radius, 0, 0), V(0, 0, 0), mat, tex);
}
int orc_num_primitives(orc_scene *s) { return s->nprim; }

/* ---- primitive tests -------------------------------------------------------------------------------- */
typedef struct { int prim; double t, a, b; } orc_hit; /* a,b: barycentric / quad coords; unused for spheres */

/* triangle: reference arithmetic (lib_tri_hit1) + the integrator's (t_min, t_best) window */
static inline int tri_test(const orc_prim *p, v3 o, v3 d, double tmin, double tmax, double *t, double *a, double *b) {
	v3 x; double tt, al, be;
	if (!lib_tri_hit1(p->Q, p->u, p->v, o, d, &x, &tt, &al, &be)) return 0;
	if (!(tt > tmin && tt < tmax)) return 0;
	*t = tt; *a = al; *b = be;
	return 1;
}
/* quad (NEW): plane step = Plane(Q, u x v).intersect_ray (src/basic/plane.cpp:13-27), then planar coordinates
 * alpha = w·(p x v), beta = w·(u x p) with w = n/(n·n), accepted on [0,1]^2 */
static inline int quad_test(const orc_prim *p, v3 o, v3 d, double tmin, double tmax, double *t, double *a, double *b) {
	v3 n = vcross(p->u, p->v);
	v3 nn = vnormalized(n);
	double pd = -vdot(nn, p->Q);
	double denom = vdot(nn, d);
	if (fabs(denom) < LIB_EPS) return 0;
	double tt = -(vdot(nn, o) + pd) / denom;
	if (tt < LIB_EPS) return 0;
	if (!(tt > tmin && tt < tmax)) return 0;
	v3 x = vadd(o, vscale(tt, d));
	v3 w = vdiv(n, vdot(n, n));
	v3 ph = vsub(x, p->Q);
	double al = vdot(w, vcross(ph, p->v)), be = vdot(w, vcross(p->u, ph));
	if (al < 0.0 || al > 1.0 || be < 0.0 || be > 1.0) return 0;
	*t = tt; *a = al; *b = be;
	return 1;
}
/* sphere (NEW): oc = C - o; h = d·oc; c = |oc|^2 - r^2; disc = h^2 - |d|^2 c; nearest root in (tmin,tmax) */
static inline int sphere_test(const orc_prim *p, v3 o, v3 d, double tmin, double tmax, double *t) {
	v3 oc = vsub(p->Q, o);
	double r = p->u.x;
	double a = vlen2(d), h = vdot(d, oc), c = vlen2(oc) - r * r;
	double disc = h * h - a * c;
	if (disc < 0.0) return 0;
	double sq = sqrt(disc);
	double root = (h - sq) / a;
	if (!(root > tmin && root < tmax)) {
		root = (h + sq) / a;
		if (!(root > tmin && root < tmax)) return 0;
	}
	*t = root;
	return 1;
}
#define ORC_BVH_MIN 256
static inline int prim_test(const orc_prim *p, v3 o, v3 d, double tmin, double tmax, double *t, double *a, double *b, orc_stats *st) {
	*a = 0; *b = 0;
	if (p->type == PRIM_TRIANGLE) { if (st) st->tri_tests++; return tri_test(p, o, d, tmin, tmax, t, a, b); }
	if (p->type == PRIM_QUAD) { if (st) st->quad_tests++; return quad_test(p, o, d, tmin, tmax, t, a, b); }
	if (st) st->sphere_tests++;
	return sphere_test(p, o, d, tmin, tmax, t);
}
static void prim_bounds(const orc_prim *p, double lo[3], double hi[3]) {
	v3 c[4];
	int n = 0;
	if (p->type == PRIM_SPHERE) {
		double r = fabs(p->u.x);
		c[0] = V(p->Q.x - r, p->Q.y - r, p->Q.z - r); c[1] = V(p->Q.x + r, p->Q.y + r, p->Q.z + r);
		n = 2;
	} else {
		c[0] = p->Q; c[1] = vadd(p->Q, p->u); c[2] = vadd(p->Q, p->v);
		n = 3;
		if (p->type == PRIM_QUAD) c[n++] = vadd(vadd(p->Q, p->u), p->v);
	}
	lo[0] = hi[0] = c[0].x; lo[1] = hi[1] = c[0].y; lo[2] = hi[2] = c[0].z;
	for (int i = 1; i < n; ++i) {
		lo[0] = fmin(lo[0], c[i].x); hi[0] = fmax(hi[0], c[i].x);
		lo[1] = fmin(lo[1], c[i].y); hi[1] = fmax(hi[1], c[i].y);
		lo[2] = fmin(lo[2], c[i].z); hi[2] = fmax(hi[2], c[i].z);
	}
	for (int k = 0; k < 3; ++k) { /* outward pad: the tree must never cut off a hit the scan would find (the tests carry eps slack) */
		double pad = 1e-9 * (fabs(lo[k]) + fabs(hi[k]) + (hi[k] - lo[k])) + 1e-9;
		lo[k] -= pad; hi[k] += pad;
	}
}
static const orc_scene *g_sort_scene;
static int g_sort_axis;
static int cmp_centroid(const void *a, const void *b) {
	const double *pa = g_sort_scene->pb + 6 * (size_t)*(const int *)a, *pc = g_sort_scene->pb + 6 * (size_t)*(const int *)b;
	const double ca = pa[g_sort_axis] + pa[3 + g_sort_axis], cb = pc[g_sort_axis] + pc[3 + g_sort_axis];
	return ca < cb ? -1 : (ca > cb ? 1 : (*(const int *)a - *(const int *)b));
}
static int bvh_build(orc_scene *s, int first, int count) {
	GROW(s->node, s->nnode, s->cnode, orc_node);
	int me = s->nnode++;
	orc_node nd;
	for (int k = 0; k < 3; ++k) { nd.lo[k] = INFINITY; nd.hi[k] = -INFINITY; }
	for (int i = first; i < first + count; ++i) {
		const double *b = s->pb + 6 * (size_t)s->order[i];
		for (int k = 0; k < 3; ++k) { nd.lo[k] = fmin(nd.lo[k], b[k]); nd.hi[k] = fmax(nd.hi[k], b[3 + k]); }
	}
	nd.first = first; nd.count = count; nd.left = nd.right = -1;
	if (count > 4) {
		int ax = 0;
		for (int k = 1; k < 3; ++k) if (nd.hi[k] - nd.lo[k] > nd.hi[ax] - nd.lo[ax]) ax = k;
		g_sort_scene = s; g_sort_axis = ax;
		qsort(s->order + first, (size_t)count, sizeof(int), cmp_centroid);
		nd.count = 0;
		s->node[me] = nd;
		int l = bvh_build(s, first, count / 2), r = bvh_build(s, first + count / 2, count - count / 2);
		nd.left = l; nd.right = r;
	}
	s->node[me] = nd;
	return me;
}
/* not thread-safe: called from the single-threaded entry points before any worker starts */
static void ensure_bvh(const orc_scene *cs) {
	orc_scene *s = (orc_scene *)cs;
	if (s->nprim <= ORC_BVH_MIN || s->bvh_nprim == s->nprim) return;
	free(s->order);
	s->order = (int *)malloc((size_t)s->nprim * sizeof(int));
	free(s->pb);
	s->pb = (double *)malloc((size_t)s->nprim * 6 * sizeof(double));
	for (int i = 0; i < s->nprim; ++i) { s->order[i] = i; prim_bounds(&s->prim[i], s->pb + 6 * (size_t)i, s->pb + 6 * (size_t)i + 3); }
	s->nnode = 0;
	bvh_build(s, 0, s->nprim);
	s->bvh_nprim = s->nprim;
}
static inline int box_overlaps(const orc_node *n, v3 o, v3 inv, double tmin, double tmax) {
	double t0 = tmin, t1 = tmax;
	const double oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
	for (int k = 0; k < 3; ++k) {
		double a = (n->lo[k] - oo[k]) * ii[k], b = (n->hi[k] - oo[k]) * ii[k];
		if (a != a || b != b) continue; /* 0 * inf: the origin lies on a slab plane of an axis-parallel ray — do not prune */
		if (a > b) { double c = a; a = b; b = c; }
		if (a > t0) t0 = a;
		if (b < t1) t1 = b;
	}
	return t0 <= t1 * (1.0 + 1e-12) + 1e-12;
}
static orc_hit closest_hit(const orc_scene *s, v3 o, v3 d, double tmin, orc_stats *st) {
	orc_hit best = { -1, INFINITY, 0, 0 };
	if (st) st->rays++;
	if (s->nprim <= ORC_BVH_MIN || s->bvh_nprim != s->nprim) { /* linear scan: the hittable list (include/object/object_set.h:10-12) */
		for (int i = 0; i < s->nprim; ++i) {
			double t, a, b;
			if (prim_test(&s->prim[i], o, d, tmin, best.t, &t, &a, &b, st)) { best.prim = i; best.t = t; best.a = a; best.b = b; }
		}
		return best;
	}
	const v3 inv = V(1.0 / d.x, 1.0 / d.y, 1.0 / d.z);
	int stack[128], sp = 0;
	stack[sp++] = 0;
	while (sp > 0) {
		const orc_node *n = &s->node[stack[--sp]];
		if (st) st->node_visits++;
		if (!box_overlaps(n, o, inv, tmin, best.t)) continue;
		if (n->count > 0) {
			for (int k = n->first; k < n->first + n->count; ++k) {
				const int i = s->order[k];
				double t, a, b;
				/* window closed at best.t so that a tie reaches the index comparison: the scan keeps the lower index */
				if (!prim_test(&s->prim[i], o, d, tmin, nextafter(best.t, INFINITY), &t, &a, &b, st)) continue;
				if (t < best.t || (t == best.t && i < best.prim)) { best.prim = i; best.t = t; best.a = a; best.b = b; }
			}
		} else if (sp + 2 <= 128) { stack[sp++] = n->left; stack[sp++] = n->right; }
	}
	return best;
}
/* any hit in (tmin, tmax) — occlusion query */
__attribute__((unused)) static int any_hit(const orc_scene *s, v3 o, v3 d, double tmin, double tmax, orc_stats *st) {
	if (st) st->rays++;
	for (int i = 0; i < s->nprim; ++i) {
		const orc_prim *p = &s->prim[i];
		double t, a, b;
		int ok;
		if (p->type == PRIM_TRIANGLE) { ok = tri_test(p, o, d, tmin, tmax, &t, &a, &b); if (st) st->tri_tests++; }
		else if (p->type == PRIM_QUAD) { ok = quad_test(p, o, d, tmin, tmax, &t, &a, &b); if (st) st->quad_tests++; }
		else { ok = sphere_test(p, o, d, tmin, tmax, &t); if (st) st->sphere_tests++; }
		if (ok) return 1;
	}
	return 0;
}
/* geometric unit normal (not flipped) and surface coordinates at a hit */
static void surface_at(const orc_scene *s, const orc_hit *h, v3 P, v3 *N, double uv[2]) {
	const orc_prim *p = &s->prim[h->prim];
	if (p->type == PRIM_SPHERE) {
		v3 n = vdiv(vsub(P, p->Q), p->u.x);
		*N = n;
		double theta = acos(fmax(-1.0, fmin(1.0, -n.y))), phi = atan2(-n.z, n.x) + M_PI;
		uv[0] = phi / (2 * M_PI);
		uv[1] = theta / M_PI;
	} else {
		*N = vnormalized(vcross(p->u, p->v));
		if (p->type == PRIM_TRIANGLE) { /* experiments/rt.cpp:139-142 bary2uv */
			double b0 = 1.0 - h->a - h->b;
			uv[0] = p->uv[0] * b0 + p->uv[2] * h->a + p->uv[4] * h->b;
			uv[1] = p->uv[1] * b0 + p->uv[3] * h->a + p->uv[5] * h->b;
		} else { uv[0] = h->a; uv[1] = h->b; }
	}
}

void orc_hit_batch(const orc_scene *s, int n, const double *Q, const double *D, double tmin,
	int *prim, double *t, double *P, double *N, double *uv) {
	ensure_bvh(s);
	for (int i = 0; i < n; ++i) {
		v3 o = vld(Q + 3 * i), d = vnormalized(vld(D + 3 * i));
		orc_hit h = closest_hit(s, o, d, tmin, 0);
		v3 x = vnan(), nn = vnan();
		double c[2] = { NAN, NAN };
		if (h.prim >= 0) { x = vadd(o, vscale(h.t, d)); surface_at(s, &h, x, &nn, c); }
		if (prim) prim[i] = h.prim;
		if (t) t[i] = h.prim >= 0 ? h.t : NAN;
		if (P) vst(P + 3 * i, x);
		if (N) vst(N + 3 * i, nn);
		if (uv) { uv[2 * i] = c[0]; uv[2 * i + 1] = c[1]; }
	}
}

/* ---- textures --------------------------------------------------------------------------------------- */
static double perlin_noise(const orc_texture *t, v3 p) {
	double fx = floor(p.x), fy = floor(p.y), fz = floor(p.z);
	double u = p.x - fx, v = p.y - fy, w = p.z - fz;
	int i = (int)fx, j = (int)fy, k = (int)fz;
	double uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w), acc = 0.0;
	for (int di = 0; di < 2; ++di)
		for (int dj = 0; dj < 2; ++dj)
			for (int dk = 0; dk < 2; ++dk) {
				v3 g = t->ranvec[t->perm[(i + di) & 255] ^ t->perm[256 + ((j + dj) & 255)] ^ t->perm[512 + ((k + dk) & 255)]];
				v3 wv = V(u - di, v - dj, w - dk);
				acc += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) * vdot(g, wv);
			}
	return acc;
}
static double perlin_turb(const orc_texture *t, v3 p, int depth) {
	double acc = 0.0, weight = 1.0;
	for (int i = 0; i < depth; ++i) { acc += weight * perlin_noise(t, p); weight *= 0.5; p = vscale(2.0, p); }
	return fabs(acc);
}
static v3 tex_eval(const orc_scene *s, int id, const double uv[2], v3 P) {
	const orc_texture *t = &s->tex[id];
	switch (t->kind) {
	case TEX_SOLID: return V(t->p[0], t->p[1], t->p[2]);
	case TEX_CHECKER_UV: { /* experiments/rt.cpp:98-101 */
		int xx = (int)floor(uv[0] * t->p[0]), yy = (int)floor(uv[1] * t->p[0]);
		return ((xx + yy) % 2 == 0) ? V(t->p[1], t->p[2], t->p[3]) : V(t->p[4], t->p[5], t->p[6]);
	}
	case TEX_CHECKER_3D: {
		double inv = 1.0 / t->p[0];
		int xi = (int)floor(inv * P.x), yi = (int)floor(inv * P.y), zi = (int)floor(inv * P.z);
		return ((xi + yi + zi) % 2 == 0) ? V(t->p[1], t->p[2], t->p[3]) : V(t->p[4], t->p[5], t->p[6]);
	}
	case TEX_NOISE: {
		double g = 0.5 * (1.0 + sin(t->p[0] * P.z + 10.0 * perlin_turb(t, P, 7)));
		return V(g, g, g);
	}
	default: { /* IMAGE: nearest texel, v flipped, rows as are::Texture::image_[y][x] */
		double u = fmin(1.0, fmax(0.0, uv[0])), v = 1.0 - fmin(1.0, fmax(0.0, uv[1]));
		int i = (int)(u * t->w), j = (int)(v * t->h);
		if (i > t->w - 1) i = t->w - 1;
		if (j > t->h - 1) j = t->h - 1;
		const double *c = t->rgb + ((size_t)j * t->w + i) * 3;
		return V(c[0], c[1], c[2]);
	}
	}
}
void orc_texture_batch(const orc_scene *s, int n, const int *tex, const double *uv, const double *P, double *rgb) {
	for (int i = 0; i < n; ++i) vst(rgb + 3 * i, tex_eval(s, tex[i], uv + 2 * i, vld(P + 3 * i)));
}

/* ---- sampling + scatter ----------------------------------------------------------------------------- */
/* experiments/rt.cpp:285-289 + rotateToHemisphere :50-55 — cosine-weighted direction about n */
static inline v3 cosine_dir(v3 n, double r1, double r2) {
	double phi = 2 * M_PI * r1, r2s = sqrt(r2);
	double lx = r2s * cos(phi), ly = r2s * sin(phi);
	v3 up = fabs(n.z) < 0.999 ? V(0, 0, 1) : V(1, 0, 0);
	v3 tangent = vnormalized(vcross(n, up));
	v3 bitangent = vcross(n, tangent);
	double lz = sqrt(fmax(0.0, 1 - lx * lx - ly * ly));
	return vadd(vadd(vscale(lx, tangent), vscale(ly, bitangent)), vscale(lz, n));
}
/* uniform direction on the unit sphere; same parametrisation as experiments/rt.cpp:226-231 (z = 1-2v) */
static inline v3 sphere_dir(double r0, double r1) {
	double z = 1.0 - 2.0 * r0, rxy = sqrt(fmax(0.0, 1.0 - z * z)), phi = 2 * M_PI * r1;
	return V(rxy * cos(phi), rxy * sin(phi), z);
}
static inline int mat_texture(const orc_material *m, int prim_tex) {
	int o = -1;
	if (m->kind == MAT_LAMBERTIAN || m->kind == MAT_LIGHT) o = (int)m->p[0];
	else if (m->kind == MAT_METAL) o = (int)m->p[1];
	return o >= 0 ? o : prim_tex;
}
/* returns alive; wi unit incoming, Ng geometric unit normal (either side) */
static int scatter1(const orc_scene *s, int mat, int tex, v3 wi, v3 Ng, v3 P, const double uv[2], const double r[4],
	v3 *wo, v3 *att, v3 *emit) {
	const orc_material *m = &s->mat[mat];
	int front = vdot(wi, Ng) < 0.0;
	v3 nf = front ? Ng : vneg(Ng);
	*emit = V(0, 0, 0); *att = V(0, 0, 0); *wo = vnan();
	switch (m->kind) {
	case MAT_LIGHT: {
		v3 c = tex_eval(s, mat_texture(m, tex), uv, P);
		*emit = vscale(m->p[1], c);
		return 0;
	}
	case MAT_DIELECTRIC: {
		double ior = m->p[0], ri = front ? 1.0 / ior : ior;
		double cos_t = fmin(vdot(vneg(wi), nf), 1.0), sin_t = sqrt(fmax(0.0, 1.0 - cos_t * cos_t));
		double r0 = (1 - ri) / (1 + ri); r0 = r0 * r0;
		double schlick = r0 + (1 - r0) * pow(1 - cos_t, 5);
		v3 d = (ri * sin_t > 1.0 || schlick > r[0]) ? lib_reflect1(wi, nf) : lib_refract1(wi, nf, ri);
		*wo = vnormalized(d);
		*att = V(1, 1, 1);
		return 1;
	}
	case MAT_METAL: {
		v3 c = tex_eval(s, mat_texture(m, tex), uv, P);
		v3 d = vadd(vnormalized(lib_reflect1(wi, nf)), vscale(m->p[0], sphere_dir(r[0], r[1])));
		*att = c;
		if (!(vdot(d, nf) > 0.0)) return 0;
		*wo = vnormalized(d);
		return 1;
	}
	case MAT_REFLECTIVE:
		if (r[2] < m->p[0]) {
			*wo = vnormalized(lib_reflect1(wi, nf));
			*att = V(m->p[1], m->p[2], m->p[3]);
			return 1;
		}
		__attribute__((fallthrough)); /* diffuse lobe */
	default: {
		*att = tex_eval(s, mat_texture(m, tex), uv, P);
		*wo = vnormalized(cosine_dir(nf, r[0], r[1]));
		return 1;
	}
	}
}
void orc_scatter_batch(const orc_scene *s, int n, const int *mat, const int *tex, const double *wi, const double *N,
	const double *P, const double *uv, const double *rnd, double *wo, double *att, double *emit, int *alive) {
	for (int i = 0; i < n; ++i) {
		v3 o, a, e;
		alive[i] = scatter1(s, mat[i], tex[i], vld(wi + 3 * i), vld(N + 3 * i), vld(P + 3 * i), uv + 2 * i, rnd + 4 * i, &o, &a, &e);
		vst(wo + 3 * i, o); vst(att + 3 * i, a); vst(emit + 3 * i, e);
	}
}

/* ---- camera (experiments/rt.cpp:339-343,364-366 + jitter + thin lens) -------------------------------- */
typedef struct {
	v3 pos, fwd, right, up; double sx, sy, lens_r, focus; int jitter;
	/* the same camera with rt.cpp's own fp32 operations (rt.cpp:339-343): pos, forward, right, up, aspect, scale */
	float rt_pos[3], rt_fwd[3], rt_right[3], rt_up[3], rt_aspect, rt_scale;
} cam_basis;
static cam_basis cam_setup(const orc_camera *c, int W, int H) {
	cam_basis b;
	b.pos = vld(c->pos);
	b.fwd = vnormalized(vsub(vld(c->target), b.pos));
	b.right = vnormalized(vcross(b.fwd, vld(c->up)));
	b.up = vcross(b.right, b.fwd);
	double scale = tan(c->vfov_deg * 0.5 * M_PI / 180.0);
	b.sx = ((double)W / (double)H) * scale;
	b.sy = scale;
	b.focus = c->focus_dist;
	b.jitter = c->jitter;
	{ /* rt.cpp:339-343 in fp32: forward = (target - pos).normalized(); right = forward.cross(up).normalized(); up' = right.cross(forward) */
		float px = (float)c->pos[0], py = (float)c->pos[1], pz = (float)c->pos[2];
		float fx = (float)c->target[0] - px, fy = (float)c->target[1] - py, fz = (float)c->target[2] - pz;
		float il = 1.0f / sqrtf(fx * fx + fy * fy + fz * fz);
		fx = fx * il; fy = fy * il; fz = fz * il;
		float ux = (float)c->up[0], uy = (float)c->up[1], uz = (float)c->up[2];
		float rx = fy * uz - fz * uy, ry = fz * ux - fx * uz, rz = fx * uy - fy * ux;
		il = 1.0f / sqrtf(rx * rx + ry * ry + rz * rz);
		rx = rx * il; ry = ry * il; rz = rz * il;
		b.rt_pos[0] = px; b.rt_pos[1] = py; b.rt_pos[2] = pz;
		b.rt_fwd[0] = fx; b.rt_fwd[1] = fy; b.rt_fwd[2] = fz;
		b.rt_right[0] = rx; b.rt_right[1] = ry; b.rt_right[2] = rz;
		b.rt_up[0] = ry * fz - rz * fy; b.rt_up[1] = rz * fx - rx * fz; b.rt_up[2] = rx * fy - ry * fx;
		b.rt_aspect = (float)W / H;                                         /* rt.cpp:342 */
		b.rt_scale = (float)tan((float)c->vfov_deg * 0.5f * M_PI / 180.f); /* rt.cpp:343 */
	}
	b.lens_r = c->defocus_angle_deg > 0.0 ? c->focus_dist * tan(c->defocus_angle_deg * 0.5 * M_PI / 180.0) : 0.0;
	return b;
}
static inline void cam_ray(const cam_basis *b, int W, int H, int x, int y, const double r[4], v3 *o, v3 *d) {
	double fx = (2 * (x + r[0]) / W - 1) * b->sx, fy = (1 - 2 * (y + r[1]) / H) * b->sy;
	v3 dir = vadd(b->fwd, vadd(vscale(fx, b->right), vscale(fy, b->up)));
	if (b->lens_r > 0.0) {
		double rr = b->lens_r * sqrt(r[2]), phi = 2 * M_PI * r[3];
		v3 off = vadd(vscale(rr * cos(phi), b->right), vscale(rr * sin(phi), b->up));
		*o = vadd(b->pos, off);
		*d = vnormalized(vsub(vscale(b->focus, dir), off));
	} else {
		*o = b->pos;
		*d = vnormalized(dir);
	}
}
void orc_camera_rays(const orc_camera *c, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D) {
	cam_basis b = cam_setup(c, W, H);
	for (int i = 0; i < n; ++i) {
		v3 o, d;
		cam_ray(&b, W, H, px[i], py[i], rnd + 4 * i, &o, &d);
		vst(Q + 3 * i, o); vst(D + 3 * i, d);
	}
}

/* ---- path integrator (NEW; estimator = RTIOW ray_color unrolled into a loop) ---------------------------- */
static inline v3 background(const orc_params *p, v3 d) {
	double a = 0.5 * (d.y + 1.0);
	return vadd(vscale(1.0 - a, vld(p->background_bottom)), vscale(a, vld(p->background_top)));
}
static v3 path_sample(const orc_scene *s, const cam_basis *cb, const orc_params *p, int x, int y, uint32_t sample, orc_stats *st) {
	const uint32_t pixel = (uint32_t)(y * p->width + x);
	double r[4];
	rnd4(p->seed, pixel, sample, 0, 0, r);
	if (!cb->jitter) { r[0] = 0.5; r[1] = 0.5; }
	v3 o, d, thr = V(1, 1, 1), L = V(0, 0, 0);
	cam_ray(cb, p->width, p->height, x, y, r, &o, &d);
	for (int b = 1; b <= p->max_depth; ++b) {
		orc_hit h = closest_hit(s, o, d, p->t_min, st);
		if (h.prim < 0) { L = vadd(L, vmul(thr, background(p, d))); break; }
		v3 P = vadd(o, vscale(h.t, d)), N, wo, att, emit;
		double uv[2];
		surface_at(s, &h, P, &N, uv);
		rnd4(p->seed, pixel, sample, (uint32_t)b, 0, r);
		const orc_prim *pr = &s->prim[h.prim];
		int alive = scatter1(s, pr->mat, pr->tex, d, N, P, uv, r, &wo, &att, &emit);
		L = vadd(L, vmul(thr, emit));
		if (!alive) break;
		thr = vmul(thr, att);
		o = P; d = wo;
	}
	if (!(isfinite(L.x) && isfinite(L.y) && isfinite(L.z))) L = V(0, 0, 0);
	return L;
}

/* ===================================================================================================== */
/* (3) experiments/rt.cpp's own shading loop, fp32 like the original, Philox instead of mt19937            */
/*     stream layout per (pixel, sample): counter.z = ray slot, counter.w = 1 (AO) / 2 (gather)            */
/* ===================================================================================================== */
typedef struct { float x, y, z; } f3;
static inline f3 F(float x, float y, float z) { f3 r = { x, y, z }; return r; }
static inline f3 f_add(f3 a, f3 b) { return F(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 f_sub(f3 a, f3 b) { return F(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 fsc(f3 a, float s) { return F(a.x * s, a.y * s, a.z * s); }
static inline f3 fmul3(f3 a, f3 b) { return F(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float fdot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 fcross(f3 a, f3 b) { return F(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float flen(f3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline f3 fnorm(f3 a) { return fsc(a, 1.0f / flen(a)); } /* rt.cpp:46-48 */
static inline f3 fclamp01(f3 c) { return F(fmaxf(0.f, fminf(1.f, c.x)), fmaxf(0.f, fminf(1.f, c.y)), fmaxf(0.f, fminf(1.f, c.z))); }
static inline f3 tof(v3 a) { return F((float)a.x, (float)a.y, (float)a.z); }
#define RT_EPS 1e-5f /* rt.cpp:15 */
typedef struct { int tri; float t, u, v; } rt_hit;
/* rt.cpp:123-138 — fp32 Möller–Trumbore, EPS = 1e-5, strict [0,1] barycentrics */
static inline int rt_tri(const orc_prim *p, f3 o, f3 d, float *t, float *u, float *v) {
	f3 v0 = tof(p->Q), e1 = tof(p->u), e2 = tof(p->v);
	/* the original stores vertices and forms e1 = v1 - v0 in fp32; Q+u-Q == u for the scenes used here */
	f3 h = fcross(d, e2), s = f_sub(o, v0);
	float a = fdot(e1, h);
	if (fabsf(a) < RT_EPS) return 0;
	float f = 1.f / a;
	*u = f * fdot(s, h);
	if (*u < 0 || *u > 1) return 0;
	f3 q = fcross(s, e1);
	*v = f * fdot(d, q);
	if (*v < 0 || *u + *v > 1) return 0;
	*t = f * fdot(e2, q);
	return *t > RT_EPS;
}
/* rt.cpp:209-218 */
static rt_hit rt_intersect(const orc_scene *s, f3 o, f3 d, orc_stats *st) {
	rt_hit r = { -1, 1e30f, 0, 0 };
	for (int i = 0; i < s->nprim; ++i) {
		float t, u, v;
		if (rt_tri(&s->prim[i], o, d, &t, &u, &v) && t < r.t) { r.tri = i; r.t = t; r.u = u; r.v = v; }
	}
	if (st) { st->rays++; st->tri_tests += (uint64_t)s->nprim; }
	return r;
}
static inline f3 rt_normal(const orc_prim *p) { return fnorm(fcross(tof(p->u), tof(p->v))); } /* rt.cpp:121 */
static inline f3 rt_point(const orc_prim *p, float u, float v) { /* rt.cpp:258 */
	f3 v0 = tof(p->Q), v1 = f_add(v0, tof(p->u)), v2 = f_add(v0, tof(p->v));
	return f_add(f_add(fsc(v0, 1 - u - v), fsc(v1, u)), fsc(v2, v));
}
static inline f3 rt_albedo(const orc_scene *s, const orc_prim *p, float u, float v) { /* rt.cpp:139-142,85-102 */
	float b0 = 1 - u - v;
	float uu = (float)p->uv[0] * b0 + (float)p->uv[2] * u + (float)p->uv[4] * v;
	float vv = (float)p->uv[1] * b0 + (float)p->uv[3] * u + (float)p->uv[5] * v;
	const orc_texture *t = &s->tex[p->tex];
	if (t->kind == TEX_CHECKER_UV) {
		int xx = (int)floorf(uu * (float)t->p[0]), yy = (int)floorf(vv * (float)t->p[0]);
		return ((xx + yy) % 2 == 0) ? F((float)t->p[1], (float)t->p[2], (float)t->p[3]) : F((float)t->p[4], (float)t->p[5], (float)t->p[6]);
	}
	return F((float)t->p[0], (float)t->p[1], (float)t->p[2]);
}
static inline void rnd4f(uint64_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t d, float r[4]) {
	double x[4];
	rnd4(seed, a, b, c, d, x);
	for (int i = 0; i < 4; ++i) r[i] = (float)x[i];
}
/* rt.cpp:221-248 */
static float rt_ao(const orc_scene *s, f3 p, f3 n, const orc_params *pp, uint32_t pixel, uint32_t sample, uint32_t slot_base, orc_stats *st) {
	const int N = pp->ao_samples;
	int unocc = 0;
	for (int i = 0; i < N; ++i) {
		float r[4];
		rnd4f(pp->seed, pixel, sample, slot_base + (uint32_t)i, 1, r);
		float theta = (float)(2 * M_PI * r[0]); /* rt.cpp:227: the product is formed in double, then narrowed */
		float phi = acosf(1 - 2 * r[1]);
		float x = sinf(phi) * cosf(theta), y = sinf(phi) * sinf(theta), z = cosf(phi);
		if (z < 0) z = -z;
		f3 hemi = F(x, y, z), axis = fcross(F(0, 0, 1), n);
		float sa = flen(axis), ca = fdot(F(0, 0, 1), n);
		f3 d = hemi;
		if (sa > RT_EPS) {
			axis = fnorm(axis);
			float ang = acosf(ca);
			d = f_add(f_add(fsc(d, cosf(ang)), fsc(fcross(axis, d), sinf(ang))), fsc(axis, fdot(axis, d) * (1 - cosf(ang))));
		}
		d = fnorm(d);
		if (rt_intersect(s, f_add(p, fsc(n, RT_EPS)), d, st).tri == -1) unocc++;
	}
	return 0.25f + 0.75f * (unocc / (float)N);
}
/* rt.cpp:50-55 */
static inline f3 rt_rotate(f3 normal, float u, float v) {
	f3 up = fabsf(normal.z) < 0.999f ? F(0, 0, 1) : F(1, 0, 0);
	f3 tangent = fnorm(fcross(normal, up));
	f3 bitangent = fcross(normal, tangent);
	return f_add(f_add(fsc(tangent, u), fsc(bitangent, v)), fsc(normal, sqrtf(fmaxf(0.f, 1 - u * u - v * v))));
}
/* rt.cpp:251-334. Materials: MAT_REFLECTIVE(p0 = reflect_ratio, p1..3 = metal_tint) == MAT_METAL there; anything
 * else == MAT_DIFFUSE there. */
static f3 rt_trace(const orc_scene *s, f3 o, f3 d, const orc_params *pp, uint32_t pixel, uint32_t sample, int depth, orc_stats *st) {
	rt_hit rec = rt_intersect(s, o, d, st);
	if (rec.tri == -1) return F(0.06f, 0.09f, 0.14f);
	const orc_prim *tri = &s->prim[rec.tri];
	const orc_material *m = &s->mat[tri->mat];
	f3 n = rt_normal(tri), p = rt_point(tri, rec.u, rec.v), albedo = rt_albedo(s, tri, rec.u, rec.v);
	const uint32_t N = (uint32_t)pp->ao_samples;
	float ao = rt_ao(s, p, n, pp, pixel, sample, (uint32_t)depth * N, st);
	if (m->kind == MAT_REFLECTIVE && depth == 0) {
		f3 view = fnorm(f_sub(o, p));
		f3 refl = f_sub(view, fsc(n, 2 * fdot(view, n))); /* rt.cpp:269 — note: this is -reflect(d,n) */
		f3 reflected = rt_trace(s, f_add(p, fsc(n, RT_EPS)), refl, pp, pixel, sample, depth + 1, st);
		float rr = (float)m->p[0];
		f3 tint = F((float)m->p[1], (float)m->p[2], (float)m->p[3]);
		f3 ret = f_add(fsc(albedo, 1 - rr), fmul3(fsc(reflected, rr), tint));
		return fclamp01(fsc(ret, ao));
	}
	if (m->kind != MAT_REFLECTIVE) {
		const int max_bounce = 3;
		f3 accum = F(0, 0, 0);
		for (uint32_t k = 0; k < N; ++k) {
			float r[4];
			uint32_t slot = ((uint32_t)depth * N + k) * 4u;
			rnd4f(pp->seed, pixel, sample, slot, 2, r);
			float phi = (float)(2 * M_PI * r[0]), r2s = sqrtf(r[1]); /* rt.cpp:286: double product, narrowed */
			f3 dir = rt_rotate(n, r2s * cosf(phi), r2s * sinf(phi));
			f3 org = f_add(p, fsc(n, RT_EPS)), thr = albedo;
			int b = 0;
			while (b < max_bounce) {
				rt_hit br = rt_intersect(s, org, dir, st);
				if (br.tri == -1) break;
				const orc_prim *bt = &s->prim[br.tri];
				f3 bn = rt_normal(bt), bp = rt_point(bt, br.u, br.v);
				thr = fmul3(thr, rt_albedo(s, bt, br.u, br.v));
				rnd4f(pp->seed, pixel, sample, slot + 1u + (uint32_t)b, 2, r);
				float pr = fmaxf(thr.x, fmaxf(thr.y, thr.z));
				if (r[0] > pr) break;
				thr = fsc(thr, 1 / pr);
				f3 nd;
				if (s->mat[bt->mat].kind == MAT_REFLECTIVE) {
					f3 view = fnorm(f_sub(F(0, 0, 0), dir));
					nd = f_sub(view, fsc(bn, 2 * fdot(view, bn)));
				} else {
					float nphi = (float)(2 * M_PI * r[1]), nr2s = sqrtf(r[2]);
					nd = rt_rotate(bn, nr2s * cosf(nphi), nr2s * sinf(nphi));
				}
				org = f_add(bp, fsc(bn, RT_EPS));
				dir = nd;
				b++;
			}
			accum = f_add(accum, thr);
		}
		albedo = f_add(albedo, fsc(accum, 1.0f / (float)N));
	}
	return fclamp01(fsc(albedo, ao));
}
static v3 rtao_sample(const orc_scene *s, const cam_basis *cb, const orc_params *p, int x, int y, uint32_t sample, orc_stats *st) {
	/* rt.cpp:364-366, pixel centres, fp32 */
	f3 fwd = F(cb->rt_fwd[0], cb->rt_fwd[1], cb->rt_fwd[2]), right = F(cb->rt_right[0], cb->rt_right[1], cb->rt_right[2]);
	f3 up = F(cb->rt_up[0], cb->rt_up[1], cb->rt_up[2]), pos = F(cb->rt_pos[0], cb->rt_pos[1], cb->rt_pos[2]);
	float fx = (2 * (x + 0.5f) / p->width - 1) * cb->rt_aspect * cb->rt_scale;
	float fy = (1 - 2 * (y + 0.5f) / p->height) * cb->rt_scale;
	f3 dir = fnorm(f_add(f_add(fwd, fsc(right, fx)), fsc(up, fy)));
	f3 c = fclamp01(rt_trace(s, pos, dir, p, (uint32_t)(y * p->width + x), sample, 0, st));
	return V(c.x, c.y, c.z);
}

/* ===================================================================================================== */
/* render driver: rows handed out dynamically to pthreads; accum += per-pixel SUM over the sample range     */
/* ===================================================================================================== */
typedef struct {
	const orc_scene *s; const orc_camera *c; const orc_params *p;
	double *accum; int y0, y1, x0, x1; volatile int next_row; pthread_mutex_t mu; orc_stats st;
} job_t;
static void *worker(void *arg) {
	job_t *j = (job_t *)arg;
	cam_basis cb = cam_setup(j->c, j->p->width, j->p->height);
	orc_stats st;
	memset(&st, 0, sizeof st);
	for (;;) {
		pthread_mutex_lock(&j->mu);
		int y = j->next_row++;
		pthread_mutex_unlock(&j->mu);
		if (y >= j->y1) break;
		for (int x = j->x0; x < j->x1; ++x) {
			v3 sum = V(0, 0, 0);
			for (int k = 0; k < j->p->sample_count; ++k) {
				uint32_t sample = (uint32_t)(j->p->sample_begin + k);
				v3 L = j->p->integrator == 1 ? rtao_sample(j->s, &cb, j->p, x, y, sample, &st) : path_sample(j->s, &cb, j->p, x, y, sample, &st);
				sum = vadd(sum, L);
				st.samples++;
			}
			double *a = j->accum + ((size_t)y * j->p->width + x) * 3;
			a[0] += sum.x; a[1] += sum.y; a[2] += sum.z;
		}
	}
	pthread_mutex_lock(&j->mu);
	j->st.samples += st.samples; j->st.rays += st.rays; j->st.tri_tests += st.tri_tests;
	j->st.quad_tests += st.quad_tests; j->st.sphere_tests += st.sphere_tests;
	pthread_mutex_unlock(&j->mu);
	return 0;
}
/* Renders the window [x0,x1) x [y0,y1) of the full W x H image (a window lets the CPU baseline time a bounded
 * sample of a large frame); accum is the full W*H*3 double buffer and is ADDED to. */
int orc_render_window(const orc_scene *s, const orc_camera *c, const orc_params *p, int x0, int y0, int x1, int y1,
	double *accum, int nthreads, orc_stats *stats) {
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 256) nthreads = 256;
	ensure_bvh(s);
	job_t j;
	memset(&j, 0, sizeof j);
	j.s = s; j.c = c; j.p = p; j.accum = accum; j.x0 = x0; j.x1 = x1; j.y0 = y0; j.y1 = y1; j.next_row = y0;
	pthread_mutex_init(&j.mu, 0);
	pthread_t th[256];
	for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], 0, worker, &j);
	for (int i = 0; i < nthreads; ++i) pthread_join(th[i], 0);
	pthread_mutex_destroy(&j.mu);
	if (stats) *stats = j.st;
	return 0;
}
int orc_render(const orc_scene *s, const orc_camera *c, const orc_params *p, double *accum, int nthreads, orc_stats *stats) {
	return orc_render_window(s, c, p, 0, 0, p->width, p->height, accum, nthreads, stats);
}
