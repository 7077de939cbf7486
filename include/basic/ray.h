// <basic/ray.h> — are::Ray: origin Q, direction D.  As in the reference (include/basic/ray.h:8-18,
// src/basic/ray.cpp:5-10) the two-argument constructor stores the NORMALISED direction, and at(t) normalises again,
// so t is always a geometric distance.
#pragma once

#include <basic/vec3.h>

namespace are {

struct Ray {
	Point3 Q;
	Vec3 D;

	Ray() = default;
	Ray(const Point3 &origin, const Vec3 &direction) : Q(origin), D(direction.normalized()) {}

	Point3 at(double t) { return Q + D.normalized() * t; }
};

}  // namespace are
