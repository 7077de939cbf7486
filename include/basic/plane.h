// <basic/plane.h> — are::Plane, the set { P : normal·P + d = 0 } (reference include/basic/plane.h:10-21,
// src/basic/plane.cpp:7-27).  The point-normal constructor normalises; the ray test rejects near-parallel rays and
// hits closer than GEOMETRY_EPSILON.
#pragma once

#include <basic/math.h>
#include <basic/ray.h>
#include <basic/vec3.h>

#include <cmath>

namespace are {

struct Plane {
	Vec3 normal;
	double d;

	Plane() = default;
	Plane(Vec3 normal_, double d_) : normal(normal_), d(d_) {}
	Plane(const Point3 &p, const Vec3 &n) : normal(n.normalized()), d(-normal.dot(p)) {}

	bool intersect_ray(const Ray &ray, Point3 &intersection) const {
		const double facing = normal.dot(ray.D);
		if (std::abs(facing) < GEOMETRY_EPSILON) return false;
		const double t = -(normal.dot(ray.Q) + d) / facing;
		if (t < GEOMETRY_EPSILON) return false;
		intersection = ray.Q + t * ray.D;
		return true;
	}
};

}  // namespace are
