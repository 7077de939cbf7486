// <basic/vec3.h> — are::Vec3 / Point3 / Color3, reflect, refract.
//
// Same public surface and conventions as the reference class (reference include/basic/vec3.h:9-86,
// src/basic/vec3.cpp): three doubles, standard layout (sizeof == 24), division by zero and normalising a zero
// vector yield an all-NaN vector instead of throwing, division is a multiplication by the reciprocal, near_zero uses
// a 1e-8 per-component threshold, refract() takes |1 - |r_perp|^2| under the root (so total internal reflection is
// not reported).  Everything is defined inline here; there is no vec3.cpp in this tree.
#pragma once

#include <cmath>
#include <limits>

namespace are {

class Vec3 {
public:
	Vec3() : c_{ 0.0, 0.0, 0.0 } {}
	Vec3(double x, double y, double z) : c_{ x, y, z } {}
	Vec3(const Vec3 &o) : c_{ o.c_[0], o.c_[1], o.c_[2] } {}
	Vec3 &operator=(const Vec3 &o) {
		c_[0] = o.c_[0];
		c_[1] = o.c_[1];
		c_[2] = o.c_[2];
		return *this;
	}
	~Vec3() = default;

	double *e() { return c_; }
	const double *e() const { return c_; }  // additive: read-only access for the CUDA shim
	double x() const { return c_[0]; }
	double y() const { return c_[1]; }
	double z() const { return c_[2]; }
	double operator[](int i) const { return c_[i]; }
	double &operator[](int i) { return c_[i]; }

	Vec3 operator-() const { return Vec3(-c_[0], -c_[1], -c_[2]); }
	Vec3 &operator+=(const Vec3 &v) {
		for (int i = 0; i < 3; ++i) c_[i] += v.c_[i];
		return *this;
	}
	Vec3 &operator-=(const Vec3 &v) {
		for (int i = 0; i < 3; ++i) c_[i] -= v.c_[i];
		return *this;
	}
	Vec3 &operator*=(double t) {
		for (int i = 0; i < 3; ++i) c_[i] *= t;
		return *this;
	}
	/// In-place division really divides (unlike the free operator/, which multiplies by 1/t); t == 0 -> NaN vector.
	Vec3 &operator/=(double t) {
		if (t == 0.0) return poison();
		for (int i = 0; i < 3; ++i) c_[i] /= t;
		return *this;
	}

	double length_squared() const { return c_[0] * c_[0] + c_[1] * c_[1] + c_[2] * c_[2]; }
	double length() const { return std::sqrt(length_squared()); }
	Vec3 &normalize() {
		const double len = length();
		if (len == 0.0) return poison();
		return *this /= len;
	}
	Vec3 normalized() const;
	double dot(const Vec3 &v) const { return c_[0] * v.c_[0] + c_[1] * v.c_[1] + c_[2] * v.c_[2]; }
	Vec3 cross(const Vec3 &v) const {
		return Vec3(c_[1] * v.c_[2] - c_[2] * v.c_[1], c_[2] * v.c_[0] - c_[0] * v.c_[2], c_[0] * v.c_[1] - c_[1] * v.c_[0]);
	}
	bool near_zero() const {
		constexpr double tiny = 1e-8;
		return std::fabs(c_[0]) < tiny && std::fabs(c_[1]) < tiny && std::fabs(c_[2]) < tiny;
	}

	static Vec3 invalid() {
		const double q = std::numeric_limits<double>::quiet_NaN();
		return Vec3(q, q, q);
	}

private:
	Vec3 &poison() { return *this = invalid(); }
	double c_[3];
};

inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return Vec3(a.x() + b.x(), a.y() + b.y(), a.z() + b.z()); }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return Vec3(a.x() - b.x(), a.y() - b.y(), a.z() - b.z()); }
inline Vec3 operator*(const Vec3 &a, const Vec3 &b) { return Vec3(a.x() * b.x(), a.y() * b.y(), a.z() * b.z()); }
inline Vec3 operator*(double t, const Vec3 &v) { return Vec3(t * v.x(), t * v.y(), t * v.z()); }
inline Vec3 operator*(const Vec3 &v, double t) { return t * v; }
inline Vec3 operator/(const Vec3 &v, double t) { return t == 0.0 ? Vec3::invalid() : (1 / t) * v; }
inline Vec3 Vec3::normalized() const {
	const double len = length();
	return len == 0.0 ? invalid() : *this / len;
}

/// Mirror v about the plane with unit normal n.
inline Vec3 reflect(const Vec3 &v, const Vec3 &n) { return v - 2 * v.dot(n) * n; }

/// Snell refraction of the unit vector uv at a surface with unit normal n, eta = eta_incident / eta_transmitted.
inline Vec3 refract(const Vec3 &uv, const Vec3 &n, double etai_over_etat) {
	const double cos_theta = std::fmin((-uv).dot(n), 1.0);
	const Vec3 across = etai_over_etat * (uv + cos_theta * n);
	const Vec3 along = -std::sqrt(std::fabs(1.0 - across.length_squared())) * n;
	const double guard = along.length();
	if (std::isnan(guard) || std::isinf(guard)) return Vec3::invalid();
	return across + along;
}

using Point3 = Vec3;
using Color3 = Vec3;

}  // namespace are
