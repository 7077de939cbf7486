// <material/metal.h> — are::Metal(fuzz) (NEW): specular reflection are::reflect(d, n) perturbed by fuzz times a
// uniform unit vector; the ray is absorbed when the perturbed direction points into the surface.
#pragma once

#include <material/material.h>
#include <texture.h>

namespace are {

class Metal : public Material {
public:
	explicit Metal(double fuzz, const Texture *albedo = nullptr) : fuzz_(fuzz < 1.0 ? (fuzz < 0.0 ? 0.0 : fuzz) : 1.0), albedo_(albedo) {}
	int kind() const override { return METAL; }
	void describe(double p[8]) const override {
		Material::describe(p);
		p[0] = fuzz_;
	}
	const Texture *texture_override() const override { return albedo_; }
	double fuzz_;

private:
	const Texture *albedo_;
};

}  // namespace are
