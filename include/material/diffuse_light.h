// <material/diffuse_light.h> — are::DiffuseLight(scale) (NEW): an emitter.  Radiance = scale * texture colour
// (the primitive's texture, or `emit` when given), emitted from both faces; it scatters nothing.
#pragma once

#include <material/material.h>
#include <texture.h>

namespace are {

class DiffuseLight : public Material {
public:
	explicit DiffuseLight(double scale = 1.0, const Texture *emit = nullptr) : scale_(scale), emit_(emit) {}
	int kind() const override { return DIFFUSE_LIGHT; }
	void describe(double p[8]) const override {
		Material::describe(p);
		p[1] = scale_;
	}
	const Texture *texture_override() const override { return emit_; }
	double scale_;

private:
	const Texture *emit_;
};

}  // namespace are
