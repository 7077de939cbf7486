// <object/triangle.h> — are::Triangle(Q, u, v, material, texture): vertices Q, Q+u, Q+v.
//
// Reference surface and behaviour (include/object/triangle.h:15-46, src/object/triangle.cpp:9-128):
//   * the constructor throws std::invalid_argument for a null material / texture, for near-zero u or v, and for
//     collinear u, v (u x v near zero) — same messages;
//   * point_in: barycentric containment with +-GEOMETRY_EPSILON slack;
//   * intersect_ray: Moeller-Trumbore, two-sided, |a| < eps rejects, alpha/beta slack eps, accepts t > eps, returns
//     the hit point Q_ray + t D;
//   * get_vertices(): {Q, Q+u, Q+v}.
// trace_texture is declared by the reference but never defined there (nothing that instantiates a Triangle links
// against it); here it returns a copy of the triangle's own texture, which is what the reference's specification
// prescribes for a surface that does not recurse (experiments/Request.md:14).
// Additive: material() / texture() / origin() / edge_u() / edge_v() getters and optional per-vertex texture
// coordinates, needed to hand the triangle to the GPU.
#pragma once

#include <basic/math.h>
#include <basic/plane.h>
#include <basic/ray.h>
#include <basic/vec3.h>
#include <material/material.h>
#include <object/object.h>
#include <texture.h>

#include <cmath>
#include <stdexcept>
#include <vector>

namespace are {

class Triangle : public Object {
public:
	Triangle() = delete;
	Triangle(const Point3 &Q_, const Vec3 &u_, const Vec3 &v_, Material *material, Texture *texture)
		: material_(material), texture_(texture), Q(Q_), u(u_), v(v_) {
		if (!material_) throw std::invalid_argument("Material pointer cannot be null");
		if (!texture_) throw std::invalid_argument("Texture pointer cannot be null");
		if (u.near_zero()) throw std::invalid_argument("Edge vector u cannot be zero vector");
		if (v.near_zero()) throw std::invalid_argument("Edge vector v cannot be zero vector");
		if (u.cross(v).near_zero()) throw std::invalid_argument("Edge vectors u and v cannot be collinear");
		vertices_ = { Q, Q + u, Q + v };
	}
	~Triangle() override = default;

	bool point_in(const Point3 &point) const override {
		const Vec3 rel = point - Q;
		const double vv = v.dot(v), vu = v.dot(u), vr = v.dot(rel), uu = u.dot(u), ur = u.dot(rel);
		double det = vv * uu - vu * vu;
		if (std::abs(det) < GEOMETRY_EPSILON) return false;
		det = 1.0 / det;
		const double alpha = (uu * vr - vu * ur) * det, beta = (vv * ur - vu * vr) * det;
		return alpha >= -GEOMETRY_EPSILON && beta >= -GEOMETRY_EPSILON && alpha + beta <= 1.0 + GEOMETRY_EPSILON;
	}

	bool intersect_ray(const Ray &ray, Point3 &hit_point) const override {
		const Vec3 h = ray.D.cross(v);
		const double a = u.dot(h);
		if (std::abs(a) < GEOMETRY_EPSILON) return false;
		const double f = 1.0 / a;
		const Vec3 s = ray.Q - Q;
		const double alpha = f * s.dot(h);
		if (alpha < -GEOMETRY_EPSILON || alpha > 1.0 + GEOMETRY_EPSILON) return false;
		const Vec3 q = s.cross(u);
		const double beta = f * ray.D.dot(q);
		if (beta < -GEOMETRY_EPSILON || alpha + beta > 1.0 + GEOMETRY_EPSILON) return false;
		const double t = f * v.dot(q);
		if (!(t > GEOMETRY_EPSILON)) return false;
		hit_point = ray.Q + t * ray.D;
		return true;
	}

	Texture trace_texture(const ObjectSet &, const Point3 &) const override { return *texture_; }

	const std::vector<Point3> &get_vertices() const { return vertices_; }

	// ---- additive ----
	Material *material() const { return material_; }
	Texture *texture() const { return texture_; }
	const Point3 &origin() const { return Q; }
	const Vec3 &edge_u() const { return u; }
	const Vec3 &edge_v() const { return v; }
	/// texture coordinates of Q, Q+u, Q+v (default (0,0), (1,0), (0,1))
	void set_uv(double u0, double v0, double u1, double v1, double u2, double v2) {
		const double t[6] = { u0, v0, u1, v1, u2, v2 };
		for (int i = 0; i < 6; ++i) uv_[i] = t[i];
		has_uv_ = true;
	}
	bool has_uv() const { return has_uv_; }
	const double *uv() const { return uv_; }

private:
	Material *material_;
	Texture *texture_;
	Point3 Q;
	Vec3 u, v;
	std::vector<Point3> vertices_;
	double uv_[6] = { 0, 0, 1, 0, 0, 1 };
	bool has_uv_ = false;
};

}  // namespace are
