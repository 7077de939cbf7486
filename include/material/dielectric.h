// <material/dielectric.h> — are::Dielectric(index_of_refraction) (NEW): clear glass.  Refraction uses are::refract;
// because that routine never reports total internal reflection (it takes |1 - |r_perp|^2| under the root), the test
// eta * sin(theta) > 1 is made by the material, together with Schlick's reflectance.
#pragma once

#include <material/material.h>

#include <stdexcept>

namespace are {

class Dielectric : public Material {
public:
	explicit Dielectric(double index_of_refraction) : ior_(index_of_refraction) {
		if (!(index_of_refraction > 0.0)) throw std::invalid_argument("Index of refraction must be positive");
	}
	int kind() const override { return DIELECTRIC; }
	void describe(double p[8]) const override {
		Material::describe(p);
		p[0] = ior_;
	}
	double ior_;
};

}  // namespace are
