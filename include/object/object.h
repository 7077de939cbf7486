// <object/object.h> — are::Object, the polymorphic base of all geometry.
// Reference surface (include/object/object.h:14-39, src/object/object.cpp:7-17): protected default constructor,
// non-copyable, virtual destructor; the three virtual queries default to "no" / an empty texture.
#pragma once

#include <basic/ray.h>
#include <basic/vec3.h>
#include <texture.h>

namespace are {

struct ObjectSet;  // <object/object_set.h>

class Object {
protected:
	Object() = default;

public:
	Object(const Object &) = delete;
	Object &operator=(const Object &) = delete;
	Object(Object &) = delete;
	Object &operator=(Object &) = delete;
	virtual ~Object() = default;

	/// Is `point` inside the (planar) primitive?
	virtual bool point_in(const Point3 & /*point*/) const { return false; }
	/// Nearest intersection of `ray` with the primitive along its positive direction; reports the hit POINT.
	virtual bool intersect_ray(const Ray & /*ray*/, Point3 & /*hit_point*/) const { return false; }
	/// What the primitive shows when looked at from `viewport_origin_point`.
	virtual Texture trace_texture(const ObjectSet & /*object_set*/, const Point3 & /*viewport_origin_point*/) const { return Texture(); }
};

}  // namespace are
