// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A flat extern "C" face over the *real* reference library (NanoEra/aurora-rendering-engine),
// compiled from the sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libare_ref.so.  Nothing of the reference is copied into this repository: this file
// only #includes the reference headers at build time and calls the reference's own routines, so the
// C restatement in oracle/are_oracle.c (and, through it, the CUDA kernels) can be pinned against
// the genuine implementation.
//
// Reference entry points exercised (all paths under /root/reference):
//   are::Vec3 ops, reflect, refract          src/basic/vec3.cpp:149-199
//   are::Ray ctor (normalises D), at()       src/basic/ray.cpp:5-10
//   are::Plane ctor / intersect_ray          src/basic/plane.cpp:7-27
//   are::Triangle ctor / point_in / hit      src/object/triangle.cpp:9-121
//   are::Diffuse / are::Reflective::reflect  src/material/diffuse.cpp:5-7, reflective.cpp:9-26
//   are::Texture load / pixel / save         src/texture.cpp:9-83, 362-395
//
// The reference never defines are::Triangle::trace_texture (decl include/object/triangle.h:31),
// so this TU supplies the one missing definition needed to link (returns an empty Texture, which
// is what the base class does, src/object/object.cpp:15-17).

#include <basic/math.h>
#include <basic/plane.h>
#include <basic/ray.h>
#include <basic/vec3.h>
#include <material/diffuse.h>
#include <material/reflective.h>
#include <object/object_set.h>
#include <object/triangle.h>
#include <texture.h>

#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace are {
Texture Triangle::trace_texture(const ObjectSet &, const Point3 &) const {
	return Texture();
}
} // namespace are

namespace {

inline are::Vec3 ld(const double *p) {
	return are::Vec3(p[0], p[1], p[2]);
}
inline void st(double *p, const are::Vec3 &v) {
	p[0] = v.x();
	p[1] = v.y();
	p[2] = v.z();
}

struct TriSet {
	are::Diffuse material;
	are::Texture texture { 1, 1, are::Color3(1, 1, 1) };
	std::vector<std::unique_ptr<are::Triangle>> tris;
	are::ObjectSet set;
};

} // namespace

extern "C" {

// sizeof of the reference's ABI-visible types: Vec3, Ray, Plane, Triangle, Texture.
void ref_sizeof(int out[5]) {
	out[0] = (int)sizeof(are::Vec3);
	out[1] = (int)sizeof(are::Ray);
	out[2] = (int)sizeof(are::Plane);
	out[3] = (int)sizeof(are::Triangle);
	out[4] = (int)sizeof(are::Texture);
}

double ref_geometry_epsilon(void) {
	return are::GEOMETRY_EPSILON;
}

// op: 0 add, 1 sub, 2 mul(component), 3 cross, 4 a*s, 5 a/s, 6 normalized(a), 7 -a
void ref_vec3_binary(int op, int n, const double *a, const double *b, const double *s, double *out) {
	for (int i = 0; i < n; ++i) {
		are::Vec3 A = ld(a + 3 * i), B = b ? ld(b + 3 * i) : are::Vec3();
		double k = s ? s[i] : 0.0;
		are::Vec3 r;
		switch (op) {
		case 0: r = A + B; break;
		case 1: r = A - B; break;
		case 2: r = A * B; break;
		case 3: r = A.cross(B); break;
		case 4: r = A * k; break;
		case 5: r = A / k; break;
		case 6: r = A.normalized(); break;
		default: r = -A; break;
		}
		st(out + 3 * i, r);
	}
}

// op: 0 dot, 1 length, 2 length_squared, 3 near_zero (0/1)
void ref_vec3_scalar(int op, int n, const double *a, const double *b, double *out) {
	for (int i = 0; i < n; ++i) {
		are::Vec3 A = ld(a + 3 * i), B = b ? ld(b + 3 * i) : are::Vec3();
		switch (op) {
		case 0: out[i] = A.dot(B); break;
		case 1: out[i] = A.length(); break;
		case 2: out[i] = A.length_squared(); break;
		default: out[i] = A.near_zero() ? 1.0 : 0.0; break;
		}
	}
}

void ref_reflect(int n, const double *v, const double *nrm, double *out) {
	for (int i = 0; i < n; ++i)
		st(out + 3 * i, are::reflect(ld(v + 3 * i), ld(nrm + 3 * i)));
}

void ref_refract(int n, const double *uv, const double *nrm, const double *eta, double *out) {
	for (int i = 0; i < n; ++i)
		st(out + 3 * i, are::refract(ld(uv + 3 * i), ld(nrm + 3 * i), eta[i]));
}

// Ray ctor stores a normalised direction; at(t) = Q + normalize(D) * t.
void ref_ray(int n, const double *Q, const double *D, const double *t, double *outD, double *outAt) {
	for (int i = 0; i < n; ++i) {
		are::Ray r(ld(Q + 3 * i), ld(D + 3 * i));
		st(outD + 3 * i, r.D);
		if (outAt)
			st(outAt + 3 * i, r.at(t ? t[i] : 0.0));
	}
}

// Plane(point, normal): out4 = (nx, ny, nz, d)
void ref_plane_from_point_normal(int n, const double *p, const double *nrm, double *out4) {
	for (int i = 0; i < n; ++i) {
		are::Plane pl(ld(p + 3 * i), ld(nrm + 3 * i));
		st(out4 + 4 * i, pl.normal);
		out4[4 * i + 3] = pl.d;
	}
}

// Plane(normal, d).intersect_ray(Ray(Q, D)) — one plane per ray.
void ref_plane_intersect(int n, const double *plane4, const double *Q, const double *D, int *hit, double *P) {
	for (int i = 0; i < n; ++i) {
		are::Plane pl(ld(plane4 + 4 * i), plane4[4 * i + 3]);
		are::Ray r(ld(Q + 3 * i), ld(D + 3 * i));
		are::Point3 x(NAN, NAN, NAN);
		hit[i] = pl.intersect_ray(r, x) ? 1 : 0;
		st(P + 3 * i, x);
	}
}

// Material::reflect (viewport-origin mirroring). kind 0 = Diffuse, 1 = Reflective(reflectivity).
void ref_material_reflect(int kind, double reflectivity, int n, const double *plane4, const double *origin, int *ok, double *out) {
	are::Diffuse dif;
	are::Reflective refl(reflectivity);
	const are::Material *m = kind == 0 ? static_cast<const are::Material *>(&dif) : &refl;
	for (int i = 0; i < n; ++i) {
		are::Plane pl(ld(plane4 + 4 * i), plane4[4 * i + 3]);
		are::Point3 o(NAN, NAN, NAN);
		ok[i] = m->reflect(pl, ld(origin + 3 * i), o) ? 1 : 0;
		st(out + 3 * i, o);
	}
}

// Triangle ctor validation. Returns 0 ok, 1 invalid_argument, 2 other exception.
// flags bit0: pass null material, bit1: pass null texture.
int ref_triangle_ctor(const double *Q, const double *u, const double *v, int flags, double *verts9) {
	are::Diffuse mat;
	are::Texture tex(1, 1, are::Color3(0, 0, 0));
	try {
		are::Triangle t(ld(Q), ld(u), ld(v), (flags & 1) ? nullptr : &mat, (flags & 2) ? nullptr : &tex);
		if (verts9) {
			const auto &vs = t.get_vertices();
			for (int k = 0; k < 3; ++k)
				st(verts9 + 3 * k, vs[k]);
		}
		return 0;
	} catch (const std::invalid_argument &) {
		return 1;
	} catch (...) {
		return 2;
	}
}

// Opaque triangle set built from (Q,u,v) triples; pushes Triangle* into an are::ObjectSet.
void *ref_triset_create(int ntri, const double *Q, const double *u, const double *v) {
	auto *ts = new TriSet();
	try {
		for (int i = 0; i < ntri; ++i) {
			ts->tris.emplace_back(new are::Triangle(ld(Q + 3 * i), ld(u + 3 * i), ld(v + 3 * i), &ts->material, &ts->texture));
			ts->set.triangles.push_back(ts->tris.back().get());
		}
	} catch (...) {
		delete ts;
		return nullptr;
	}
	return ts;
}

void ref_triset_destroy(void *h) {
	delete static_cast<TriSet *>(h);
}

// Every (ray, triangle) pair: hit[r*ntri+k], P[(r*ntri+k)*3..] straight from Triangle::intersect_ray
// through the Object* virtual (src/object/triangle.cpp:82-121).
void ref_triset_hit_matrix(void *h, int nrays, const double *Q, const double *D, int *hit, double *P) {
	auto *ts = static_cast<TriSet *>(h);
	const int nt = (int)ts->set.triangles.size();
	for (int r = 0; r < nrays; ++r) {
		are::Ray ray(ld(Q + 3 * r), ld(D + 3 * r));
		for (int k = 0; k < nt; ++k) {
			const are::Object *o = ts->set.triangles[k];
			are::Point3 x(NAN, NAN, NAN);
			bool ok = o->intersect_ray(ray, x);
			hit[(size_t)r * nt + k] = ok ? 1 : 0;
			st(P + ((size_t)r * nt + k) * 3, x);
		}
	}
}

// Closest hit over the ObjectSet as the renderer drives it: the reference returns a hit
// *point*, so t is recovered as (P - Q)·D with D the Ray's normalised direction; ties keep the lower index.
// prim = -1 on miss. Returns the number of rays that hit something.
long ref_triset_closest_hit(void *h, int nrays, const double *Q, const double *D, int *prim, double *t, double *P) {
	auto *ts = static_cast<TriSet *>(h);
	const int nt = (int)ts->set.triangles.size();
	long nhit = 0;
	for (int r = 0; r < nrays; ++r) {
		are::Ray ray(ld(Q + 3 * r), ld(D + 3 * r));
		int best = -1;
		double bt = std::numeric_limits<double>::infinity();
		are::Point3 bp(NAN, NAN, NAN);
		for (int k = 0; k < nt; ++k) {
			const are::Object *o = ts->set.triangles[k];
			are::Point3 x;
			if (o->intersect_ray(ray, x)) {
				double tk = (x - ray.Q).dot(ray.D);
				if (tk < bt) {
					bt = tk;
					best = k;
					bp = x;
				}
			}
		}
		prim[r] = best;
		t[r] = best >= 0 ? bt : NAN;
		st(P + 3 * r, bp);
		nhit += best >= 0;
	}
	return nhit;
}

void ref_triset_point_in(void *h, int tri, int n, const double *pts, int *inside) {
	auto *ts = static_cast<TriSet *>(h);
	const are::Object *o = ts->set.triangles[tri];
	for (int i = 0; i < n; ++i)
		inside[i] = o->point_in(ld(pts + 3 * i)) ? 1 : 0;
}

// Texture(path): returns 0 ok (w,h filled, rgb = w*h*3 doubles row-major if non-null), 1 runtime_error, 2 other.
int ref_texture_load(const char *path, int *w, int *h, double *rgb, long cap) {
	try {
		std::unique_ptr<are::Texture> t(new are::Texture(std::string(path)));
		*w = t->width_;
		*h = t->height_;
		if (rgb) {
			if ((long)t->width_ * t->height_ * 3 > cap)
				return 3;
			for (int y = 0; y < t->height_; ++y)
				for (int x = 0; x < t->width_; ++x)
					st(rgb + ((size_t)y * t->width_ + x) * 3, t->pixel(x, y));
		}
		return 0;
	} catch (const std::runtime_error &) {
		return 1;
	} catch (...) {
		return 2;
	}
}

// Texture(w,h,fill) + pixel(x,y) = rgb + save_texture(path). Returns 1 saved, 0 save_texture said false, <0 exception.
int ref_texture_save(const char *path, int w, int h, const double *rgb) {
	try {
		std::unique_ptr<are::Texture> t(new are::Texture(w, h, are::Color3(0, 0, 0)));
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x)
				t->pixel(x, y) = ld(rgb + ((size_t)y * w + x) * 3);
		return t->save_texture(std::string(path)) ? 1 : 0;
	} catch (const std::runtime_error &) {
		return -1;
	} catch (...) {
		return -2;
	}
}

// Texture::paste (src/texture.cpp:85-360) on images given as w*h*3 doubles; dst is updated in place.
// corners = { lt.x, lt.y, rt.x, rt.y, lb.x, lb.y, rb.x, rb.y }.  Returns 0, or 1 on runtime_error.
int ref_texture_paste(double *dst, int dw, int dh, const double *src, int sw, int sh, const int *corners) {
	try {
		std::unique_ptr<are::Texture> d(new are::Texture(dw, dh, are::Color3(0, 0, 0)));
		std::unique_ptr<are::Texture> s(new are::Texture(sw, sh, are::Color3(0, 0, 0)));
		for (int y = 0; y < dh; ++y)
			for (int x = 0; x < dw; ++x) d->pixel(x, y) = ld(dst + ((size_t)y * dw + x) * 3);
		for (int y = 0; y < sh; ++y)
			for (int x = 0; x < sw; ++x) s->pixel(x, y) = ld(src + ((size_t)y * sw + x) * 3);
		d->paste(*s, { corners[0], corners[1] }, { corners[2], corners[3] }, { corners[4], corners[5] }, { corners[6], corners[7] });
		for (int y = 0; y < dh; ++y)
			for (int x = 0; x < dw; ++x) st(dst + ((size_t)y * dw + x) * 3, d->pixel(x, y));
		return 0;
	} catch (const std::runtime_error &) {
		return 1;
	}
}

// Texture(w,h,fill) ctor error behaviour: 0 ok, 1 runtime_error.
int ref_texture_fill_ctor(int w, int h) {
	try {
		std::unique_ptr<are::Texture> t(new are::Texture(w, h, are::Color3(0.5, 0.5, 0.5)));
		return 0;
	} catch (const std::runtime_error &) {
		return 1;
	}
}

} // extern "C"
