// <object/quad.h> — are::Quad(Q, u, v, material, texture) (NEW): the parallelogram Q + a u + b v, a, b in [0,1].
// Validation as are::Triangle.  The ray test is are::Plane(Q, u x v).intersect_ray (reference
// src/basic/plane.cpp:13-27) followed by the planar coordinates a = w·(p x v), b = w·(u x p), w = n / (n·n).
#pragma once

#include <basic/plane.h>
#include <material/material.h>
#include <object/object.h>
#include <texture.h>

#include <stdexcept>
#include <vector>

namespace are {

class Quad : public Object {
public:
	Quad() = delete;
	Quad(const Point3 &Q_, const Vec3 &u_, const Vec3 &v_, Material *material, Texture *texture)
		: material_(material), texture_(texture), Q(Q_), u(u_), v(v_) {
		if (!material_) throw std::invalid_argument("Material pointer cannot be null");
		if (!texture_) throw std::invalid_argument("Texture pointer cannot be null");
		if (u.near_zero()) throw std::invalid_argument("Edge vector u cannot be zero vector");
		if (v.near_zero()) throw std::invalid_argument("Edge vector v cannot be zero vector");
		if (u.cross(v).near_zero()) throw std::invalid_argument("Edge vectors u and v cannot be collinear");
		vertices_ = { Q, Q + u, Q + u + v, Q + v };
	}

	bool planar_coords(const Point3 &p, double &a, double &b) const {
		const Vec3 n = u.cross(v);
		const Vec3 w = n / n.dot(n);
		const Vec3 rel = p - Q;
		a = w.dot(rel.cross(v));
		b = w.dot(u.cross(rel));
		return a >= 0.0 && a <= 1.0 && b >= 0.0 && b <= 1.0;
	}
	bool point_in(const Point3 &point) const override {
		double a, b;
		return planar_coords(point, a, b);
	}
	bool intersect_ray(const Ray &ray, Point3 &hit_point) const override {
		Point3 p;
		if (!Plane(Q, u.cross(v)).intersect_ray(ray, p)) return false;
		double a, b;
		if (!planar_coords(p, a, b)) return false;
		hit_point = p;
		return true;
	}
	Texture trace_texture(const ObjectSet &, const Point3 &) const override { return *texture_; }

	const std::vector<Point3> &get_vertices() const { return vertices_; }
	Material *material() const { return material_; }
	Texture *texture() const { return texture_; }
	const Point3 &origin() const { return Q; }
	const Vec3 &edge_u() const { return u; }
	const Vec3 &edge_v() const { return v; }

private:
	Material *material_;
	Texture *texture_;
	Point3 Q;
	Vec3 u, v;
	std::vector<Point3> vertices_;
};

}  // namespace are
