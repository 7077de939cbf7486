// <object/object_set.h> — are::ObjectSet: the scene, a bag of non-owning primitive pointers.
// The reference holds triangles only (include/object/object_set.h:10-12); quads and spheres are additive.
#pragma once

#include <object/object.h>
#include <object/quad.h>
#include <object/sphere.h>
#include <object/triangle.h>

#include <vector>

namespace are {

struct ObjectSet {
	std::vector<Triangle *> triangles;
	std::vector<Quad *> quads;      // additive
	std::vector<Sphere *> spheres;  // additive
};

}  // namespace are
