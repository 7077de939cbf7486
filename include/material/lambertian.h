// <material/lambertian.h> — are::Lambertian (NEW; no counterpart in the reference): ideal diffuse reflector.
// Albedo = the primitive's texture, or `albedo` when one is given.  Sampling = the cosine-weighted gather of the
// reference's only Monte-Carlo prototype (experiments/rt.cpp:285-289 with rotateToHemisphere :50-55).
#pragma once

#include <material/material.h>
#include <texture.h>

namespace are {

class Lambertian : public Material {
public:
	Lambertian() = default;
	explicit Lambertian(const Texture *albedo) : albedo_(albedo) {}
	int kind() const override { return LAMBERTIAN; }
	const Texture *texture_override() const override { return albedo_; }

private:
	const Texture *albedo_ = nullptr;  // non-owning, like every pointer in this API
};

}  // namespace are
