// cornell_host.cpp — a reference-style host program rendering through the GPU core.
//
// The scene is built exactly the way a user of the reference builds one: are::Texture and are::Material objects owned
// by the host, are::Triangle(Q, u, v, Material*, Texture*) for every surface, raw pointers pushed into an
// are::ObjectSet.  The only new line is the hand-off to are::cuda::Renderer, which replaces the per-pixel CPU loop of
// the reference's prototypes (experiments/rt.cpp:353-371).
//
//   g++ -std=c++17 -O2 -Iinclude examples/cornell_host.cpp -Laurora_rendering_engine_b200/lib -lare_b200 [continued]
//       -Wl,-rpath,$PWD/aurora_rendering_engine_b200/lib -o cornell_host
//   ./cornell_host out.ppm [width height spp [raw_float_dump [n_gpus]]]      n_gpus: 0 = every visible GPU (default 1)
#include <are_cuda.hpp>
#include <camera.h>
#include <material/diffuse_light.h>
#include <material/lambertian.h>
#include <object/object_set.h>
#include <texture.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

using namespace are;

struct Scene {
	std::vector<std::unique_ptr<Texture>> textures;
	std::vector<std::unique_ptr<Material>> materials;
	std::vector<std::unique_ptr<Triangle>> triangles;
	ObjectSet set;

	Texture *solid(double r, double g, double b) {
		textures.emplace_back(new Texture(Texture::solid(Color3(r, g, b))));
		return textures.back().get();
	}
	// two triangles (p0,p1,p2), (p0,p2,p3), texture coordinates as rt.cpp's pushQuad assigns them
	void quad(const Point3 &p0, const Point3 &p1, const Point3 &p2, const Point3 &p3, Material *m, Texture *t) {
		triangles.emplace_back(new Triangle(p0, p1 - p0, p2 - p0, m, t));
		triangles.back()->set_uv(0, 0, 1, 0, 1, 1);
		set.triangles.push_back(triangles.back().get());
		triangles.emplace_back(new Triangle(p0, p2 - p0, p3 - p0, m, t));
		triangles.back()->set_uv(0, 0, 1, 1, 0, 1);
		set.triangles.push_back(triangles.back().get());
	}
	void flat(const Point3 &Q, const Vec3 &u, const Vec3 &v, Material *m, Texture *t) { quad(Q, Q + u, Q + u + v, Q + v, m, t); }
	void box(const Vec3 &size, double rot_y_deg, const Vec3 &offset, Material *m, Texture *t) {
		const double th = rot_y_deg * 3.14159265358979323846 / 180.0, c = std::cos(th), s = std::sin(th);
		auto xf = [&](const Point3 &p) { return Point3(c * p.x() + s * p.z(), p.y(), -s * p.x() + c * p.z()) + offset; };
		const Vec3 dx(size.x(), 0, 0), dy(0, size.y(), 0), dz(0, 0, size.z());
		const Point3 o(0, 0, 0);
		struct F { Point3 Q; Vec3 u, v; };
		const F faces[6] = { { o + dz, dx, dy }, { o + dx + dz, -dz, dy }, { o + dx, -dx, dy }, { o, dz, dy }, { o + dy + dz, dx, -dz }, { o, dx, dz } };
		for (const F &f : faces) quad(xf(f.Q), xf(f.Q + f.u), xf(f.Q + f.u + f.v), xf(f.Q + f.v), m, t);
	}
};

int main(int argc, char **argv) {
	const char *out_path = argc > 1 ? argv[1] : "cornell_host.ppm";
	const int W = argc > 3 ? std::atoi(argv[2]) : 256, H = argc > 3 ? std::atoi(argv[3]) : 256;
	const int spp = argc > 4 ? std::atoi(argv[4]) : 64;

	Scene sc;
	Texture *red = sc.solid(.65, .05, .05), *white = sc.solid(.73, .73, .73), *green = sc.solid(.12, .45, .15), *bright = sc.solid(15, 15, 15);
	sc.materials.emplace_back(new Lambertian());
	sc.materials.emplace_back(new DiffuseLight(1.0));
	Material *matte = sc.materials[0].get(), *lamp = sc.materials[1].get();
	sc.flat(Point3(555, 0, 0), Vec3(0, 555, 0), Vec3(0, 0, 555), matte, green);
	sc.flat(Point3(0, 0, 0), Vec3(0, 555, 0), Vec3(0, 0, 555), matte, red);
	sc.flat(Point3(343, 554, 332), Vec3(-130, 0, 0), Vec3(0, 0, -105), lamp, bright);
	sc.flat(Point3(0, 0, 0), Vec3(555, 0, 0), Vec3(0, 0, 555), matte, white);
	sc.flat(Point3(555, 555, 555), Vec3(-555, 0, 0), Vec3(0, 0, -555), matte, white);
	sc.flat(Point3(0, 0, 555), Vec3(555, 0, 0), Vec3(0, 555, 0), matte, white);
	sc.box(Vec3(165, 330, 165), 15.0, Vec3(265, 0, 295), matte, white);
	sc.box(Vec3(165, 165, 165), -18.0, Vec3(130, 0, 65), matte, white);

	Camera cam(Point3(278, 278, -800), Point3(278, 278, 0), Vec3(0, 1, 0), 40.0);
	cam.focus_dist = 10.0;

	try {
		const int want = argc > 6 ? std::atoi(argv[6]) : 1;
		std::vector<int> devices = cuda::Renderer::all_devices();
		if (want > 0 && static_cast<size_t>(want) < devices.size()) devices.resize(static_cast<size_t>(want));
		cuda::Renderer gpu(devices);  // one renderer over all of them: samples sharded, sums reduced over NVLink peer memory
		gpu.upload(sc.set);
		cuda::Settings s;
		s.spp = spp;
		s.background_bottom = s.background_top = Color3(0, 0, 0);
		if (argc > 5) {  // raw sample sums for the parity test
			std::vector<float> sums(static_cast<size_t>(W) * H * 3);
			gpu.render_sums(cam, W, H, s, sums.data());
			FILE *f = std::fopen(argv[5], "wb");
			std::fwrite(sums.data(), sizeof(float), sums.size(), f);
			std::fclose(f);
		}
		Texture image = gpu.render(cam, W, H, s);
		if (!image.save_texture(out_path)) { std::fprintf(stderr, "cannot write %s\n", out_path); return 2; }
		const are_render_stats &st = gpu.stats();
		std::printf("%d triangles, %dx%d, %d spp on %d GPU(s): %llu rays in %.2f ms (%.1f Mrays/s), wrote %s\n", (int)sc.set.triangles.size(), W, H, spp,
			gpu.device_count(), (unsigned long long)st.rays, st.kernel_ms, st.rays / (st.kernel_ms * 1e3), out_path);
	} catch (const std::exception &e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
