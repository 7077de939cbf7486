// <material/material.h> — are::Material, the polymorphic base of all surface materials.
//
// Reference surface (include/material/material.h:10-29, src/material/material.cpp): protected default constructor,
// non-copyable, virtual destructor, and ONE virtual operation, reflect(): given the plane of a viewport polygon and
// the current eye point, produce the mirrored eye point (true) or decline (false).  The base declines.
//
// Additive: kind()/describe() tell the CUDA shim how a material scatters rays in the path tracer
// (values = are_material_kind, parameters as documented in include/are_cuda.h).
#pragma once

#include <basic/plane.h>
#include <basic/vec3.h>

namespace are {

class Texture;  // <texture.h>

class Material {
protected:
	Material() = default;

public:
	Material(const Material &) = delete;
	Material &operator=(const Material &) = delete;
	Material(Material &) = delete;
	Material &operator=(Material &) = delete;
	virtual ~Material() = default;

	virtual bool reflect(const Plane & /*viewport_plane*/, const Point3 & /*viewport_origin*/, Point3 & /*new_viewport_origin*/) const { return false; }

	// ---- path-tracer description (additive) ----
	enum Kind { DIFFUSE = 0, REFLECTIVE = 1, LAMBERTIAN = 2, METAL = 3, DIELECTRIC = 4, DIFFUSE_LIGHT = 5 };
	virtual int kind() const { return DIFFUSE; }
	virtual void describe(double p[8]) const {
		for (int i = 0; i < 8; ++i) p[i] = 0.0;
	}
	/// texture that replaces the primitive's own one for this material's colour (nullptr: use the primitive's)
	virtual const Texture *texture_override() const { return nullptr; }
};

}  // namespace are
