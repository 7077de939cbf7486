// <material/reflective.h> — are::Reflective(reflectivity): a mirror-like surface.
// reflect() mirrors the eye point across the viewport plane, o' = o - 2 (n·o + d)/|n|^2 n, and declines for a
// degenerate plane (|n|^2 < GEOMETRY_EPSILON) — reference include/material/reflective.h:11-25,
// src/material/reflective.cpp:9-26.  reflectivity_ is a public data member there and stays one here.
// Path tracer: with probability reflectivity_ a perfect mirror bounce tinted by tint_, else the diffuse lobe.
#pragma once

#include <basic/math.h>
#include <material/material.h>

namespace are {

class Reflective : public Material {
public:
	Reflective() = delete;
	explicit Reflective(double reflectivity) : reflectivity_(reflectivity) {}
	Reflective(double reflectivity, const Color3 &tint) : reflectivity_(reflectivity), tint_(tint) {}  // additive
	~Reflective() override = default;

	bool reflect(const Plane &viewport_plane, const Point3 &viewport_origin, Point3 &new_viewport_origin) const override {
		const double nn = viewport_plane.normal.length_squared();
		if (nn < GEOMETRY_EPSILON) return false;
		const double signed_dist = viewport_plane.normal.dot(viewport_origin) + viewport_plane.d;
		const double k = 2.0 * signed_dist / nn;
		new_viewport_origin = viewport_origin - k * viewport_plane.normal;
		return true;
	}

	int kind() const override { return REFLECTIVE; }
	void describe(double p[8]) const override {
		Material::describe(p);
		p[0] = reflectivity_;
		p[1] = tint_.x();
		p[2] = tint_.y();
		p[3] = tint_.z();
	}

	double reflectivity_;
	Color3 tint_ = Color3(1, 1, 1);
};

}  // namespace are
