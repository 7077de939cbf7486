// vec.cuh — 3-vector algebra for the kernels, templated on the scalar so that the SAME routines run in fp64
// (decision-exact harness against the reference's are::Vec3, /root/reference/src/basic/vec3.cpp) and in fp32
// (the render kernels).  Semantics follow the reference: x/0 and normalize(0) give a NaN vector
// (vec3.cpp:90-99,112-129,174-179), division multiplies by the reciprocal (vec3.cpp:178).
#pragma once
#include "rtc_compat.h"
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#include <math.h>
#endif

namespace areb {

template <typename T>
struct V3 {
	T x, y, z;
};

template <typename T> __host__ __device__ __forceinline__ V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator-(V3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(V3<T> a, V3<T> b) { return mk<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(T t, V3<T> a) { return mk<T>(t * a.x, t * a.y, t * a.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(V3<T> a, T t) { return mk<T>(t * a.x, t * a.y, t * a.z); }
template <typename T> __host__ __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __host__ __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
	return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> __host__ __device__ __forceinline__ T len2(V3<T> a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
#if defined(__CUDA_ARCH__)
// fp32 on the device: sums of products are written out with explicit fused multiply-adds.  Left to the compiler, WHICH
// product of a*b + c*d gets fused depends on the code around the expression, and the render kernels (generic, lean,
// baked: three compilations of these routines) must round every path identically — their images are compared bit for
// bit.  (The fp64 harness is built with -fmad=false and keeps the plain expressions: it mirrors the reference's x86 code.)
__device__ __forceinline__ float dot(V3<float> a, V3<float> b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float len2(V3<float> a) { return fmaf(a.z, a.z, fmaf(a.y, a.y, a.x * a.x)); }
__device__ __forceinline__ V3<float> cross(V3<float> a, V3<float> b) {
	return mk<float>(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
#endif
// s*a + b and s*a + t*b with the roundings pinned the same way (fp64: the plain expression)
template <typename T> __host__ __device__ __forceinline__ V3<T> mad(T s, V3<T> a, V3<T> b) { return s * a + b; }
template <typename T> __host__ __device__ __forceinline__ V3<T> mad2(T s, V3<T> a, T t, V3<T> b) { return s * a + t * b; }
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ V3<float> mad(float s, V3<float> a, V3<float> b) { return mk<float>(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ V3<float> mad2(float s, V3<float> a, float t, V3<float> b) {
	return mk<float>(fmaf(s, a.x, t * b.x), fmaf(s, a.y, t * b.y), fmaf(s, a.z, t * b.z));
}
#endif

__host__ __device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__host__ __device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__host__ __device__ __forceinline__ double abs_t(double x) { return fabs(x); }
__host__ __device__ __forceinline__ float abs_t(float x) { return fabsf(x); }
__host__ __device__ __forceinline__ double min_t(double a, double b) { return fmin(a, b); }
__host__ __device__ __forceinline__ float min_t(float a, float b) { return fminf(a, b); }
__host__ __device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__host__ __device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__host__ __device__ __forceinline__ double floor_t(double x) { return floor(x); }
__host__ __device__ __forceinline__ float floor_t(float x) { return floorf(x); }
__device__ __forceinline__ void sincos_t(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __forceinline__ void sincos_t(float a, float *s, float *c) { sincosf(a, s, c); }
// sin/cos of 2*pi*r for r in [0,1): the fp32 path uses sincospif (exact range reduction, no slow path)
__device__ __forceinline__ void sincos2pi_t(double r, double *s, double *c) { sincos(2.0 * 3.14159265358979323846 * r, s, c); }
// fp32: r in [0,1) -> angle 2*pi*(r - 0.5) in [-pi, pi), where the MUFU.SIN/COS pair (__sincosf) is accurate to
// ~4e-7 absolute; sin/cos(2*pi*r) are the negatives.  Two MUFU + a handful of FMUL instead of ~50 instructions.
__device__ __forceinline__ void sincos2pi_t(float r, float *s, float *c) {
	float sn, cs;
	__sincosf(6.283185307179586f * (r - 0.5f), &sn, &cs);
	*s = -sn;
	*c = -cs;
}
__host__ __device__ __forceinline__ double sin_t(double x) { return sin(x); }
// fp32 sine of the noise texture's phase (|x| of a few tens): reduced in REVOLUTIONS — t = x / 2pi, t - rint(t) is exact —
// and handed to MUFU.SIN inside [-pi, pi], where it is good to 4e-7 absolute; with the rounding of t the result is within
// ~3e-6 of sin(x).  Six instructions at ~6 active lanes where sinf's Cody-Waite reduction + polynomial cost ~25.
__device__ __forceinline__ float sin_t(float x) {
	const float t = x * 0.15915494309189535f;
	return __sinf(6.283185307179586f * (t - rintf(t)));
}
__host__ __device__ __forceinline__ double acos_t(double x) { return acos(x); }
__host__ __device__ __forceinline__ float acos_t(float x) { return acosf(x); }
__host__ __device__ __forceinline__ double atan2_t(double y, double x) { return atan2(y, x); }
__host__ __device__ __forceinline__ float atan2_t(float y, float x) { return atan2f(y, x); }

template <typename T> __host__ __device__ __forceinline__ T nan_t() { return T(NAN); }

template <typename T> __host__ __device__ __forceinline__ T length(V3<T> a) { return sqrt_t(len2(a)); }
// reference operator/ : NaN vector on t == 0, else (1/t) * v  (vec3.cpp:174-179)
template <typename T> __host__ __device__ __forceinline__ V3<T> vdiv(V3<T> a, T t) {
	if (t == T(0)) return mk<T>(nan_t<T>(), nan_t<T>(), nan_t<T>());
	return (T(1) / t) * a;
}
// reference normalized() (vec3.cpp:123-129)
template <typename T> __host__ __device__ __forceinline__ V3<T> normalized(V3<T> a) {
	T l = length(a);
	if (l == T(0)) return mk<T>(nan_t<T>(), nan_t<T>(), nan_t<T>());
	return vdiv(a, l);
}
// fast fp32 normalisation for the render loop: one MUFU.RSQ, no NaN convention needed (callers guarantee |a|>0)
__device__ __forceinline__ V3<float> normalize_fast(V3<float> a) { return rsqrtf(len2(a)) * a; }

// are::reflect (vec3.cpp:182-184): v - 2*dot(v,n)*n
template <typename T> __host__ __device__ __forceinline__ V3<T> reflect(V3<T> v, V3<T> n) { return v - (T(2) * dot(v, n)) * n; }
// are::refract (vec3.cpp:188-199): the fabs under the sqrt is the reference's, so TIR is NOT detected here
template <typename T> __host__ __device__ __forceinline__ V3<T> refract(V3<T> uv, V3<T> n, T eta) {
	T cos_theta = min_t(dot(-uv, n), T(1));
	V3<T> perp = eta * (uv + cos_theta * n);
	V3<T> par = (-sqrt_t(abs_t(T(1) - len2(perp)))) * n;
	T l = length(par);
	if (l != l || (l - l) != T(0)) return mk<T>(nan_t<T>(), nan_t<T>(), nan_t<T>());  // isnan || isinf
	return perp + par;
}

template <typename T> __host__ __device__ __forceinline__ V3<T> ld3(const double *p) { return mk<T>(T(p[0]), T(p[1]), T(p[2])); }
template <typename T> __host__ __device__ __forceinline__ void st3(double *p, V3<T> a) { p[0] = double(a.x); p[1] = double(a.y); p[2] = double(a.z); }

}  // namespace areb
