// harness64.cu — the fp64, decision-exact per-ray kernels behind are_cuda_hit_batch / scatter_batch /
// texture_batch / camera_rays with precision = 64, plus the Philox batch.
//
// This translation unit is compiled with -fmad=false so that +,-,*,/ and sqrt round exactly as the reference's
// x86-64 build does (no FMA contraction): hit/miss decisions, the winning primitive, t and the hit point of
// are::Triangle::intersect_ray (/root/reference/src/object/triangle.cpp:82-121) are reproduced BIT-FOR-BIT;
// only libm transcendentals (sin, cos, acos, atan2) may differ in the last ulp.
#include "dev_types.h"
#include "kernels.h"
#include "philox.cuh"
#include "shade.cuh"
#include "vec.cuh"

namespace areb {

#define LIB_EPS 1e-12  // are::GEOMETRY_EPSILON, include/basic/math.h:7

typedef V3<double> D3;

// src/object/triangle.cpp:82-121, literally, + the integrator's (tmin, tbest) window
__device__ __forceinline__ bool tri64(const double *q9, D3 o, D3 d, double tmin, double tmax, double &t, double &al, double &be) {
	D3 Q = mk<double>(q9[0], q9[1], q9[2]), u = mk<double>(q9[3], q9[4], q9[5]), v = mk<double>(q9[6], q9[7], q9[8]);
	D3 h = cross(d, v);
	double a = dot(u, h);
	if (fabs(a) < LIB_EPS) return false;
	double f = 1.0 / a;
	D3 s = o - Q;
	double alpha = f * dot(s, h);
	if (alpha < -LIB_EPS || alpha > 1.0 + LIB_EPS) return false;
	D3 q = cross(s, u);
	double beta = f * dot(d, q);
	if (beta < -LIB_EPS || alpha + beta > 1.0 + LIB_EPS) return false;
	double tt = f * dot(v, q);
	if (!(tt > LIB_EPS)) return false;
	if (!(tt > tmin && tt < tmax)) return false;
	t = tt; al = alpha; be = beta;
	return true;
}
// plane step = are::Plane(Q, u x v).intersect_ray (src/basic/plane.cpp:10-27), then planar coordinates
__device__ __forceinline__ bool quad64(const double *q9, D3 o, D3 d, double tmin, double tmax, double &t, double &al, double &be) {
	D3 Q = mk<double>(q9[0], q9[1], q9[2]), u = mk<double>(q9[3], q9[4], q9[5]), v = mk<double>(q9[6], q9[7], q9[8]);
	D3 n = cross(u, v);
	D3 nn = normalized(n);
	double pd = -dot(nn, Q);
	double denom = dot(nn, d);
	if (fabs(denom) < LIB_EPS) return false;
	double tt = -(dot(nn, o) + pd) / denom;
	if (tt < LIB_EPS) return false;
	if (!(tt > tmin && tt < tmax)) return false;
	D3 x = o + tt * d;
	D3 w = vdiv(n, dot(n, n));
	D3 ph = x - Q;
	double a = dot(w, cross(ph, v)), b = dot(w, cross(u, ph));
	if (a < 0.0 || a > 1.0 || b < 0.0 || b > 1.0) return false;
	t = tt; al = a; be = b;
	return true;
}
__device__ __forceinline__ bool sphere64(const double *c4, D3 o, D3 d, double tmin, double tmax, double &t) {
	D3 oc = mk<double>(c4[0], c4[1], c4[2]) - o;
	double r = c4[3];
	double a = len2(d), h = dot(d, oc), c = len2(oc) - r * r;
	double disc = h * h - a * c;
	if (disc < 0.0) return false;
	double sq = sqrt(disc);
	double root = (h - sq) / a;
	if (!(root > tmin && root < tmax)) {
		root = (h + sq) / a;
		if (!(root > tmin && root < tmax)) return false;
	}
	t = root;
	return true;
}

__global__ void k_hit64(DevScene sc, int n, const double *__restrict__ Q, const double *__restrict__ D, double tmin,
	int *__restrict__ prim, double *__restrict__ tout, double *__restrict__ P, double *__restrict__ N, double *__restrict__ uv) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	D3 o = ld3<double>(Q + 3 * i), d = normalized(ld3<double>(D + 3 * i));  // are::Ray ctor, src/basic/ray.cpp:5
	double best = INFINITY, ba = 0, bb = 0;
	int bdev = -1, buid = 0x7fffffff;
	const int nt = sc.n_tri, nq = sc.n_quad, ns = sc.n_sph;
	for (int k = 0; k < nt + nq + ns; ++k) {
		double t, a = 0, b = 0;
		bool ok;
		// ties at equal t go to the lower USER id, as a linear scan of the ObjectSet in insertion order would decide
		if (k < nt) ok = tri64(sc.tri64 + 9 * k, o, d, tmin, INFINITY, t, a, b);
		else if (k < nt + nq) ok = quad64(sc.quad64 + 9 * (k - nt), o, d, tmin, INFINITY, t, a, b);
		else ok = sphere64(sc.sph64 + 4 * (k - nt - nq), o, d, tmin, INFINITY, t);
		if (ok) {
			int uid = sc.info[k].user_id;
			if (t < best || (t == best && uid < buid)) { best = t; bdev = k; buid = uid; ba = a; bb = b; }
		}
	}
	const double nan = nan_t<double>();
	D3 x = mk<double>(nan, nan, nan), nn = x;
	double cu = nan, cv = nan;
	if (bdev >= 0) {
		x = o + best * d;
		if (bdev >= nt + nq) {
			const double *c4 = sc.sph64 + 4 * (bdev - nt - nq);
			nn = vdiv(x - mk<double>(c4[0], c4[1], c4[2]), c4[3]);
			double theta = acos(fmax(-1.0, fmin(1.0, -nn.y))), phi = atan2(-nn.z, nn.x) + 3.14159265358979323846;
			cu = phi / (2 * 3.14159265358979323846);
			cv = theta / 3.14159265358979323846;
		} else {
			const double *q9 = bdev < nt ? sc.tri64 + 9 * bdev : sc.quad64 + 9 * (bdev - nt);
			nn = normalized(cross(mk<double>(q9[3], q9[4], q9[5]), mk<double>(q9[6], q9[7], q9[8])));
			if (bdev < nt) {
				const double *t6 = sc.tri_uv64 + 6 * bdev;
				double b0 = 1.0 - ba - bb;
				cu = t6[0] * b0 + t6[2] * ba + t6[4] * bb;
				cv = t6[1] * b0 + t6[3] * ba + t6[5] * bb;
			} else { cu = ba; cv = bb; }
		}
	}
	if (prim) prim[i] = bdev >= 0 ? buid : -1;
	if (tout) tout[i] = bdev >= 0 ? best : nan;
	if (P) st3(P + 3 * i, x);
	if (N) st3(N + 3 * i, nn);
	if (uv) { uv[2 * i] = cu; uv[2 * i + 1] = cv; }
}

__global__ void k_scatter64(DevScene sc, int n, const int *mat, const int *tex, const double *wi, const double *N, const double *P,
	const double *uv, const double *rnd, double *wo, double *att, double *emit, int *alive) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Rnd4<double> r;
	r.x = rnd[4 * i]; r.y = rnd[4 * i + 1]; r.z = rnd[4 * i + 2]; r.w = rnd[4 * i + 3];
	D3 o, a, e;
	bool ok = scatter<double>(sc, mat[i], tex[i], ld3<double>(wi + 3 * i), ld3<double>(N + 3 * i), ld3<double>(P + 3 * i), uv[2 * i], uv[2 * i + 1], r, o, a, e);
	st3(wo + 3 * i, o); st3(att + 3 * i, a); st3(emit + 3 * i, e);
	alive[i] = ok ? 1 : 0;
}

__global__ void k_texture64(DevScene sc, int n, const int *tex, const double *uv, const double *P, double *rgb) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	st3(rgb + 3 * i, tex_eval<double>(sc, tex[i], uv[2 * i], uv[2 * i + 1], ld3<double>(P + 3 * i)));
}

__global__ void k_camera64(CamBasis cb, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	CamT<double> c = cam_from_basis<double>(cb);
	Rnd4<double> r;
	r.x = rnd[4 * i]; r.y = rnd[4 * i + 1]; r.z = rnd[4 * i + 2]; r.w = rnd[4 * i + 3];
	D3 o, d;
	cam_ray<double>(c, 1.0 / W, 1.0 / H, px[i], py[i], r, o, d);
	st3(Q + 3 * i, o); st3(D + 3 * i, d);
}

__global__ void k_philox(int n, uint64_t seed, const uint32_t *ctr, uint32_t *out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	PhiloxKey k = philox_key(seed);
	U4 o = philox4x32_10(k, ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]);
	out[4 * i] = o.x; out[4 * i + 1] = o.y; out[4 * i + 2] = o.z; out[4 * i + 3] = o.w;
}

static inline int blocks(int n) { return (n + 127) / 128; }

void launch_hit64(const DevScene &sc, int n, const double *Q, const double *D, double tmin, int *prim, double *t, double *P, double *N, double *uv, cudaStream_t s) {
	if (n > 0) k_hit64<<<blocks(n), 128, 0, s>>>(sc, n, Q, D, tmin, prim, t, P, N, uv);
}
void launch_scatter64(const DevScene &sc, int n, const int *mat, const int *tex, const double *wi, const double *N, const double *P, const double *uv,
	const double *rnd, double *wo, double *att, double *emit, int *alive, cudaStream_t s) {
	if (n > 0) k_scatter64<<<blocks(n), 128, 0, s>>>(sc, n, mat, tex, wi, N, P, uv, rnd, wo, att, emit, alive);
}
void launch_texture64(const DevScene &sc, int n, const int *tex, const double *uv, const double *P, double *rgb, cudaStream_t s) {
	if (n > 0) k_texture64<<<blocks(n), 128, 0, s>>>(sc, n, tex, uv, P, rgb);
}
void launch_camera64(const CamBasis &cb, int W, int H, int n, const int *px, const int *py, const double *rnd, double *Q, double *D, cudaStream_t s) {
	if (n > 0) k_camera64<<<blocks(n), 128, 0, s>>>(cb, W, H, n, px, py, rnd, Q, D);
}
void launch_philox(int n, uint64_t seed, const uint32_t *ctr, uint32_t *out, cudaStream_t s) {
	if (n > 0) k_philox<<<blocks(n), 128, 0, s>>>(n, seed, ctr, out);
}

}  // namespace areb

// =========================================================================================================
// Texture::paste on the device (reference src/texture.cpp:85-360): one thread per destination pixel of the clipped
// bounding box — even-odd point-in-quad at the pixel centre, inverse homography, bilinear fetch, all in fp64 with the
// reference's operation order (this TU is built with -fmad=false), so the pasted texels are bit-identical.
// The 8x8 solve and 3x3 inverse are done once on the host (api.cu).
// =========================================================================================================
namespace areb {

__global__ void k_paste(PasteArgs a) {
	const int x = a.x0 + blockIdx.x * blockDim.x + threadIdx.x;
	const int y = a.y0 + blockIdx.y * blockDim.y + threadIdx.y;
	if (x > a.x1 || y > a.y1) return;
	const double X = (double)x + 0.5, Y = (double)y + 0.5;
	bool inside = false;
	for (int i = 0, j = 3; i < 4; j = i++) {  // boundary order LT, RT, RB, LB
		const double xi = a.qx[i], yi = a.qy[i], xj = a.qx[j], yj = a.qy[j];
		const bool cross = ((yi > Y) != (yj > Y)) && (X < (xj - xi) * (Y - yi) / ((yj - yi) == 0.0 ? 1e-30 : (yj - yi)) + xi);
		inside ^= cross;
	}
	if (!inside) return;
	const double denom = a.hinv[6] * X + a.hinv[7] * Y + a.hinv[8];
	if (fabs(denom) < LIB_EPS) return;
	double sx = (a.hinv[0] * X + a.hinv[1] * Y + a.hinv[2]) / denom;
	double sy = (a.hinv[3] * X + a.hinv[4] * Y + a.hinv[5]) / denom;
	const double sw = (double)a.sw, sh = (double)a.sh;
	if (sx < 0.0 || sy < 0.0 || sx > sw - 1.0 || sy > sh - 1.0) return;
	sx = fmin(fmax(sx, 0.0), sw - 1.0);
	sy = fmin(fmax(sy, 0.0), sh - 1.0);
	const int ix = (int)floor(sx), iy = (int)floor(sy);
	const int ix1 = min(ix + 1, a.sw - 1), iy1 = min(iy + 1, a.sh - 1);
	const double tx = sx - (double)ix, ty = sy - (double)iy;
	const double w00 = (1.0 - tx) * (1.0 - ty), w10 = tx * (1.0 - ty), w01 = (1.0 - tx) * ty, w11 = tx * ty;
	const double *c00 = a.src + ((size_t)iy * a.sw + ix) * 3, *c10 = a.src + ((size_t)iy * a.sw + ix1) * 3;
	const double *c01 = a.src + ((size_t)iy1 * a.sw + ix) * 3, *c11 = a.src + ((size_t)iy1 * a.sw + ix1) * 3;
	double *out = a.dst + ((size_t)y * a.dw + x) * 3;
	for (int k = 0; k < 3; ++k) out[k] = w00 * c00[k] + w10 * c10[k] + w01 * c01[k] + w11 * c11[k];
}

// ---- library routines of SURVEY §8a that are not on the render loop, as fp64 batches (bit-exact, -fmad=false) ----
// are::Plane::intersect_ray, src/basic/plane.cpp:13-27 (the ray direction is normalised first, as are::Ray's ctor does)
__global__ void k_plane64(int n, const double *__restrict__ plane4, const double *__restrict__ Q, const double *__restrict__ D, int *__restrict__ hit,
	double *__restrict__ P) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const D3 pn = ld3<double>(plane4 + 4 * i), q = ld3<double>(Q + 3 * i), d = normalized(ld3<double>(D + 3 * i));
	const double pd = plane4[4 * i + 3];
	const double nan = nan_t<double>();
	D3 x = mk<double>(nan, nan, nan);
	int ok = 0;
	const double denom = dot(pn, d);
	if (!(fabs(denom) < LIB_EPS)) {
		const double t = -(dot(pn, q) + pd) / denom;
		if (!(t < LIB_EPS)) { x = q + t * d; ok = 1; }
	}
	hit[i] = ok;
	st3(P + 3 * i, x);
}
// are::Triangle::point_in, src/object/triangle.cpp:49-79
__global__ void k_point_in64(int n, const double *__restrict__ tri9, const double *__restrict__ pts, int *__restrict__ inside) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const D3 Q = ld3<double>(tri9), u = ld3<double>(tri9 + 3), v = ld3<double>(tri9 + 6), p = ld3<double>(pts + 3 * i);
	const D3 v0 = v, v1 = u, v2 = p - Q;
	const double d00 = dot(v0, v0), d01 = dot(v0, v1), d02 = dot(v0, v2), d11 = dot(v1, v1), d12 = dot(v1, v2);
	double inv = d00 * d11 - d01 * d01;
	int ok = 0;
	if (!(fabs(inv) < LIB_EPS)) {
		inv = 1.0 / inv;
		const double alpha = (d11 * d02 - d01 * d12) * inv, beta = (d00 * d12 - d01 * d02) * inv;
		ok = alpha >= -LIB_EPS && beta >= -LIB_EPS && alpha + beta <= 1.0 + LIB_EPS;
	}
	inside[i] = ok;
}
// are::Material::reflect: Diffuse declines (src/material/diffuse.cpp:5-7), Reflective mirrors the viewport origin across
// the plane, o' = o - 2 (n.o + d)/|n|^2 n (src/material/reflective.cpp:9-26)
__global__ void k_material_reflect64(int kind, int n, const double *__restrict__ plane4, const double *__restrict__ origin, int *__restrict__ ok,
	double *__restrict__ out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double nan = nan_t<double>();
	D3 o = mk<double>(nan, nan, nan);
	int good = 0;
	if (kind == MK_REFLECTIVE) {
		const D3 pn = ld3<double>(plane4 + 4 * i), org = ld3<double>(origin + 3 * i);
		const double denom = len2(pn);
		if (!(denom < LIB_EPS)) {
			const double numer = dot(pn, org) + plane4[4 * i + 3];
			const double k = 2.0 * numer / denom;
			o = org - k * pn;
			good = 1;
		}
	}
	ok[i] = good;
	st3(out + 3 * i, o);
}
void launch_plane64(int n, const double *plane4, const double *Q, const double *D, int *hit, double *P, cudaStream_t s) {
	if (n > 0) k_plane64<<<(n + 127) / 128, 128, 0, s>>>(n, plane4, Q, D, hit, P);
}
void launch_point_in64(int n, const double *tri9, const double *pts, int *inside, cudaStream_t s) {
	if (n > 0) k_point_in64<<<(n + 127) / 128, 128, 0, s>>>(n, tri9, pts, inside);
}
void launch_material_reflect64(int kind, int n, const double *plane4, const double *origin, int *ok, double *out, cudaStream_t s) {
	if (n > 0) k_material_reflect64<<<(n + 127) / 128, 128, 0, s>>>(kind, n, plane4, origin, ok, out);
}

void launch_paste(const PasteArgs &a, cudaStream_t s) {
	if (a.x1 < a.x0 || a.y1 < a.y0) return;
	dim3 block(32, 8), grid((a.x1 - a.x0 + 32) / 32, (a.y1 - a.y0 + 8) / 8);
	k_paste<<<grid, block, 0, s>>>(a);
}

}  // namespace areb
