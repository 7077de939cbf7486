// <material/diffuse.h> — are::Diffuse: a matte surface.  reflect() declines ("do not recurse": reference
// include/material/diffuse.h:11-21, src/material/diffuse.cpp:5-7); the path tracer renders it as a Lambertian
// surface coloured by the primitive's texture.
#pragma once

#include <material/material.h>

namespace are {

class Diffuse : public Material {
public:
	Diffuse() = default;
	~Diffuse() override = default;
	bool reflect(const Plane &, const Point3 &, Point3 &) const override { return false; }
	int kind() const override { return DIFFUSE; }
};

}  // namespace are
