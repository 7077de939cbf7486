// philox.cuh — Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw, SC'11; Random123 constants).
// key = render seed, counter = (pixel index, GLOBAL sample index, bounce, stream): every random number of the
// job is addressable, so any GPU can render any sample range and the 1-GPU and N-GPU estimators draw the very
// same samples.  The ten round keys depend only on the seed; they are computed once per thread.
#pragma once
#include "rtc_compat.h"

namespace areb {

struct PhiloxKey {
	uint32_t k0[10], k1[10];
};

__host__ __device__ __forceinline__ PhiloxKey philox_key(uint64_t seed) {
	PhiloxKey k;
	uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		k.k0[r] = a;
		k.k1[r] = b;
		a += 0x9E3779B9u;
		b += 0xBB67AE85u;
	}
	return k;
}

struct U4 {
	uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ U4 philox4x32_10(const PhiloxKey &k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
	for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
		uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
		uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
		uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
		uint32_t n0 = h1 ^ c1 ^ k.k0[r], n2 = h0 ^ c3 ^ k.k1[r];
		c0 = n0;
		c1 = l1;
		c2 = n2;
		c3 = l0;
	}
	U4 o;
	o.x = c0; o.y = c1; o.z = c2; o.w = c3;
	return o;
}

// 24-bit uniform in [0,1): (x >> 8) * 2^-24 — exactly representable in fp32, identical on host and device
template <typename T>
__host__ __device__ __forceinline__ T u01(uint32_t x) { return T(x >> 8) * T(1.0 / 16777216.0); }

template <typename T>
struct Rnd4 {
	T x, y, z, w;
};

template <typename T>
__host__ __device__ __forceinline__ Rnd4<T> rnd4(const PhiloxKey &k, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t stream) {
	U4 o = philox4x32_10(k, pixel, sample, bounce, stream);
	Rnd4<T> r;
	r.x = u01<T>(o.x); r.y = u01<T>(o.y); r.z = u01<T>(o.z); r.w = u01<T>(o.w);
	return r;
}

}  // namespace areb
