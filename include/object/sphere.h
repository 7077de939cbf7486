// <object/sphere.h> — are::Sphere(center, radius, material, texture) (NEW).
// Ray test: oc = C - Q, h = D·oc, c = |oc|^2 - r^2, disc = h^2 - |D|^2 c, nearest root beyond GEOMETRY_EPSILON.
// Texture coordinates: u = (atan2(-n.z, n.x) + pi) / 2 pi, v = acos(-n.y) / pi for the outward normal n.
#pragma once

#include <basic/math.h>
#include <material/material.h>
#include <object/object.h>
#include <texture.h>

#include <cmath>
#include <stdexcept>

namespace are {

class Sphere : public Object {
public:
	Sphere() = delete;
	Sphere(const Point3 &center, double radius, Material *material, Texture *texture)
		: material_(material), texture_(texture), center_(center), radius_(radius) {
		if (!material_) throw std::invalid_argument("Material pointer cannot be null");
		if (!texture_) throw std::invalid_argument("Texture pointer cannot be null");
		if (!(radius > 0.0)) throw std::invalid_argument("Sphere radius must be positive");
	}
	bool point_in(const Point3 &point) const override { return (point - center_).length_squared() <= radius_ * radius_; }
	bool intersect_ray(const Ray &ray, Point3 &hit_point) const override {
		const Vec3 oc = center_ - ray.Q;
		const double a = ray.D.length_squared(), h = ray.D.dot(oc), c = oc.length_squared() - radius_ * radius_;
		const double disc = h * h - a * c;
		if (disc < 0.0) return false;
		const double root = std::sqrt(disc);
		double t = (h - root) / a;
		if (!(t > GEOMETRY_EPSILON)) {
			t = (h + root) / a;
			if (!(t > GEOMETRY_EPSILON)) return false;
		}
		hit_point = ray.Q + t * ray.D;
		return true;
	}
	Texture trace_texture(const ObjectSet &, const Point3 &) const override { return *texture_; }

	Material *material() const { return material_; }
	Texture *texture() const { return texture_; }
	const Point3 &center() const { return center_; }
	double radius() const { return radius_; }

private:
	Material *material_;
	Texture *texture_;
	Point3 center_;
	double radius_;
};

}  // namespace are
