// host_api_probe.cpp — a host program written ONLY against the reference's public API (namespace are, headers
// <basic/...>, <material/...>, <object/...>, <texture.h>).  tests/test_cpp_api.py compiles this one source twice:
//   (1) against include/ of this repository (header-only mirror), and
//   (2) against /root/reference/include + /root/reference/src/*.cpp (the real library; -DARE_REFERENCE_BUILD supplies
//       the one definition the reference forgets, Triangle::trace_texture),
// runs both and requires byte-identical output: same values to the last bit (%.17g), same exception types and
// messages, same files written.  Where the reference tree is absent (GPU box) the output is compared with the
// committed tests/golden/host_api_probe.txt, which was produced by build (2).
#include <basic/math.h>
#include <basic/plane.h>
#include <basic/ray.h>
#include <basic/vec3.h>
#include <material/diffuse.h>
#include <material/reflective.h>
#include <object/object_set.h>
#include <object/triangle.h>
#include <texture.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef ARE_REFERENCE_BUILD
namespace are {
Texture Triangle::trace_texture(const ObjectSet &, const Point3 &) const { return Texture(); }
}  // namespace are
#endif

using namespace are;

static std::uint64_t g_state = 0x9E3779B97F4A7C15ull;
static double rnd() {  // splitmix64 -> [0,1)
	std::uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
static Vec3 rvec(double lo, double hi) {
	double x = lo + (hi - lo) * rnd(), y = lo + (hi - lo) * rnd(), z = lo + (hi - lo) * rnd();
	return Vec3(x, y, z);
}
static void show(const char *tag, const Vec3 &v) { std::printf("%s %.17g %.17g %.17g\n", tag, v.x(), v.y(), v.z()); }

int main(int argc, char **argv) {
	const std::string tmp = argc > 1 ? argv[1] : ".";
	std::printf("sizeof %zu %zu %zu eps %.17g\n", sizeof(Vec3), sizeof(Ray), sizeof(Plane), GEOMETRY_EPSILON);
	// ---- Vec3 ----
	for (int i = 0; i < 64; ++i) {
		Vec3 a = rvec(-3, 3), b = rvec(-3, 3);
		double s = 4 * rnd() - 2;
		show("add", a + b); show("sub", a - b); show("mul", a * b); show("scl", a * s); show("lcs", s * a); show("div", a / s);
		show("crs", a.cross(b)); show("neg", -a); show("nrm", a.normalized());
		std::printf("dot %.17g len %.17g len2 %.17g nz %d\n", a.dot(b), a.length(), a.length_squared(), (int)a.near_zero());
		Vec3 c = a; c += b; c -= a * 0.5; c *= s; c /= 3.0; show("cmp", c);
		Vec3 n = b.normalized();
		show("rfl", reflect(a.normalized(), n));
		double eta = 0.3 + 3 * rnd();
		show("rfr", refract(a.normalized(), n, eta));
	}
	show("nan_div", Vec3(1, 2, 3) / 0.0);
	show("nan_nrm", Vec3(0, 0, 0).normalized());
	{ Vec3 z(0, 0, 0); z.normalize(); show("nan_nrm2", z); Vec3 w(1, 1, 1); w /= 0.0; show("nan_div2", w); }
	std::printf("near_zero %d %d\n", (int)Vec3(1e-9, -1e-9, 5e-9).near_zero(), (int)Vec3(1e-9, 1e-7, 0).near_zero());
	show("kat_reflect", reflect(Vec3(1, -1, 0), Vec3(0, 1, 0)));
	show("kat_refract", refract(Vec3(1, -1, 0).normalized(), Vec3(0, 1, 0), 1 / 1.5));
	show("kat_tir", refract(Vec3(1, -1, 0).normalized(), Vec3(0, 1, 0), 3.0));
	// ---- Ray / Plane ----
	for (int i = 0; i < 32; ++i) {
		Vec3 ro = rvec(-2, 2), rd = rvec(-2, 2);
		Ray r(ro, rd);
		double t = 5 * rnd();
		show("rayD", r.D); show("rayAt", r.at(t));
		Vec3 pp = rvec(-2, 2), pn = rvec(-1, 1);
		Plane pl(pp, pn);
		std::printf("plane %.17g %.17g %.17g %.17g\n", pl.normal.x(), pl.normal.y(), pl.normal.z(), pl.d);
		Point3 x(-7, -7, -7);
		bool hit = pl.intersect_ray(r, x);
		std::printf("phit %d ", (int)hit); show("", x);
	}
	{ Plane pl(Point3(0, 0, 0), Vec3(0, 0, 2)); Point3 x(9, 9, 9); Ray par(Point3(0, 0, 1), Vec3(1, 0, 0));
	  bool h1 = pl.intersect_ray(par, x), h2 = pl.intersect_ray(Ray(Point3(0, 0, 1), Vec3(0, 0, 1)), x);
	  std::printf("parallel %d behind %d\n", (int)h1, (int)h2); }
	// ---- materials ----
	{
		Diffuse dif; Reflective refl(0.9);
		Plane pl(Point3(0, 0, 0), Vec3(0, 0, 2));
		Point3 o(7, 7, 7);
		bool r1 = dif.reflect(pl, Point3(1, 2, 3), o);
		std::printf("diffuse %d ", (int)r1); show("", o);
		bool r2 = refl.reflect(pl, Point3(1, 2, 3), o);
		std::printf("reflective %d ", (int)r2); show("", o);
		bool r3 = refl.reflect(Plane(Vec3(0, 0, 0), 1.0), Point3(1, 2, 3), o);
		std::printf("degenerate %d reflectivity %.17g\n", (int)r3, refl.reflectivity_);
		const Material *base = &refl;
		for (int i = 0; i < 16; ++i) {
			Vec3 pp = rvec(-2, 2), pn = rvec(-1, 1);
			Plane q(pp, pn);
			q.normal = q.normal * (0.5 + rnd());  // non-unit normals are legal input
			Vec3 eye = rvec(-3, 3);
			bool ok = base->reflect(q, eye, o);
			std::printf("mirror %d ", (int)ok); show("", o);
		}
	}
	// ---- Triangle ----
	Diffuse mat; Texture tex(2, 2, Color3(0.25, 0.5, 0.75));
	struct Case { Vec3 Q, u, v; Material *m; Texture *t; };
	const Case bad[] = { { Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 1, 0), nullptr, &tex }, { Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 1, 0), &mat, nullptr },
		{ Vec3(0, 0, 0), Vec3(0, 0, 0), Vec3(0, 1, 0), &mat, &tex }, { Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(1e-9, 0, 0), &mat, &tex },
		{ Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(2, 0, 0), &mat, &tex }, { Vec3(0, 0, 0), Vec3(1e-5, 0, 0), Vec3(0, 1e-5, 0), &mat, &tex } };
	for (const Case &c : bad) {
		try { Triangle t(c.Q, c.u, c.v, c.m, c.t); std::printf("ctor ok\n"); }
		catch (const std::invalid_argument &e) { std::printf("ctor invalid_argument: %s\n", e.what()); }
		catch (const std::exception &e) { std::printf("ctor other: %s\n", e.what()); }
	}
	std::vector<Triangle *> owned;
	ObjectSet set;
	for (int i = 0; i < 24; ++i) {
		Vec3 q = rvec(-1, 1), u = rvec(-1, 1), v = rvec(-1, 1);
		owned.push_back(new Triangle(q, u, v, &mat, &tex));
		set.triangles.push_back(owned.back());
	}
	owned.push_back(new Triangle(Point3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 1, 0), &mat, &tex));
	set.triangles.push_back(owned.back());
	for (const Point3 &v : owned.back()->get_vertices()) show("vert", v);
	{ Point3 h(8, 8, 8); const Object *o = owned.back();
	  bool k1 = o->intersect_ray(Ray(Point3(.25, .25, 1), Vec3(0, 0, -2)), h);
	  std::printf("kat_hit %d ", (int)k1); show("", h);
	  bool k2 = o->intersect_ray(Ray(Point3(.75, .75, 1), Vec3(0, 0, -2)), h), k3 = o->intersect_ray(Ray(Point3(.25, .25, -1), Vec3(0, 0, -2)), h);
	  std::printf("kat_miss %d %d in %d %d\n", (int)k2, (int)k3, (int)o->point_in(Point3(.2, .2, 0)), (int)o->point_in(Point3(.8, .8, 0))); }
	long hits = 0;
	for (int i = 0; i < 4000; ++i) {
		Vec3 ro = rvec(-1.5, 1.5), rd = rvec(-1, 1);
		Ray r(ro, rd);
		int best = -1; double bt = 1e300; Point3 bp(0, 0, 0);
		for (size_t k = 0; k < set.triangles.size(); ++k) {
			Point3 x;
			if (set.triangles[k]->intersect_ray(r, x)) {
				double t = (x - r.Q).dot(r.D);
				if (t < bt) { bt = t; best = (int)k; bp = x; }
			}
		}
		if (best >= 0) { ++hits; if (i % 16 == 0) { std::printf("hit %d %d %.17g ", i, best, bt); show("", bp); } }
	}
	std::printf("hits %ld\n", hits);
	for (int i = 0; i < 200; ++i) {
		const Triangle *t = set.triangles[i % set.triangles.size()];
		const std::vector<Point3> &vs = t->get_vertices();
		double a = 1.4 * rnd() - 0.2, b = 1.4 * rnd() - 0.2;
		Point3 p = vs[0] + a * (vs[1] - vs[0]) + b * (vs[2] - vs[0]);
		std::printf("%d", (int)static_cast<const Object *>(t)->point_in(p));
	}
	std::printf("\n");
	// ---- Texture ----
	{
		Texture img(5, 3, Color3(0, 0, 0));
		for (int y = 0; y < 3; ++y)
			for (int x = 0; x < 5; ++x) {
				double r = 1.3 * rnd() - 0.1, g = 1.3 * rnd() - 0.1, b = 1.3 * rnd() - 0.1;
				img.pixel(x, y) = Color3(r, g, b);
			}
		img.pixel(0, 0) = Color3(0.0, 1.0, 254.0 / 255.0);
		bool s1 = img.save_texture(tmp + "/probe.ppm"), s2 = img.save_texture(tmp + "/probe.png");
		std::printf("save %d bad_suffix %d\n", (int)s1, (int)s2);
		Texture back(tmp + "/probe.ppm");
		std::printf("loaded %d %d\n", back.width_, back.height_);
		for (int y = 0; y < 3; ++y)
			for (int x = 0; x < 5; ++x) show("px", back.pixel(x, y));
		FILE *f = std::fopen((tmp + "/probe.ppm").c_str(), "rb");
		int c, n = 0; unsigned sum = 0;
		while ((c = std::fgetc(f)) != EOF) { sum = sum * 31u + (unsigned)c; ++n; }
		std::fclose(f);
		std::printf("file %d bytes checksum %u\n", n, sum);
		try { Texture t{ std::string("/definitely/not/here.ppm") }; std::printf("load ok\n"); }
		catch (const std::runtime_error &e) { std::printf("load runtime_error: %s\n", e.what()); }
		f = std::fopen((tmp + "/p3.ppm").c_str(), "wb"); std::fputs("P3\n1 1\n255\n0 0 0\n", f); std::fclose(f);
		try { Texture t{ tmp + "/p3.ppm" }; std::printf("p3 ok\n"); } catch (const std::runtime_error &) { std::printf("p3 runtime_error\n"); }
		f = std::fopen((tmp + "/short.ppm").c_str(), "wb"); std::fputs("P6\n2 2\n255\n", f); std::fputc(1, f); std::fclose(f);
		try { Texture t{ tmp + "/short.ppm" }; std::printf("short ok\n"); } catch (const std::runtime_error &) { std::printf("short runtime_error\n"); }
		f = std::fopen((tmp + "/max.ppm").c_str(), "wb"); std::fputs("P6\n1 1\n65535\n", f); std::fclose(f);
		try { Texture t{ tmp + "/max.ppm" }; std::printf("max ok\n"); } catch (const std::runtime_error &) { std::printf("max runtime_error\n"); }
		try { Texture t(0, 4, Color3(0, 0, 0)); std::printf("fill ok\n"); } catch (const std::runtime_error &e) { std::printf("fill runtime_error: %s\n", e.what()); }
		Texture empty;
		try { empty.pixel(0, 0); std::printf("pixel ok\n"); } catch (const std::runtime_error &e) { std::printf("pixel runtime_error: %s\n", e.what()); }
		try { empty.save_texture(tmp + "/e.ppm"); std::printf("esave ok\n"); } catch (const std::runtime_error &e) { std::printf("esave runtime_error: %s\n", e.what()); }
	}
	// ---- Texture::paste ----
	{
		Texture src(7, 5, Color3(0, 0, 0));
		for (int y = 0; y < 5; ++y)
			for (int x = 0; x < 7; ++x) {
				double r = rnd(), g = rnd(), b = rnd();
				src.pixel(x, y) = Color3(r, g, b);
			}
		const int cases[5][8] = { { 2, 3, 27, 1, 1, 19, 29, 17 }, { -6, -4, 40, 3, 2, 30, 31, 23 }, { 5, 5, 5, 5, 5, 20, 5, 20 },
			{ 20, 2, 3, 4, 25, 21, 1, 18 }, { 0, 0, 31, 0, 0, 23, 31, 23 } };
		for (const int(&c)[8] : cases) {
			Texture dst(32, 24, Color3(0.5, 0.25, 0.125));
			dst.paste(src, { c[0], c[1] }, { c[2], c[3] }, { c[4], c[5] }, { c[6], c[7] });
			int changed = 0;
			double sum = 0;
			for (int y = 0; y < 24; ++y)
				for (int x = 0; x < 32; ++x) {
					const Color3 &q = dst.pixel(x, y);
					if (q.x() != 0.5 || q.y() != 0.25 || q.z() != 0.125) ++changed;
					sum += q.x() * (1 + x) + q.y() * (1 + y) + q.z();
				}
			std::printf("paste changed %d checksum %.17g ", changed, sum); show("centre", dst.pixel(15, 11));
		}
		Texture empty;
		try { empty.paste(src, { 0, 0 }, { 1, 0 }, { 0, 1 }, { 1, 1 }); std::printf("epaste ok\n"); } catch (const std::runtime_error &e) { std::printf("epaste runtime_error: %s\n", e.what()); }
	}
	for (Triangle *t : owned) delete t;
	return 0;
}
