// patch_host.cpp — the reference's own rendering algorithm, driven through its class API.
//
// Builds the Cornell room of the reference's patch renderer (experiments/rt10.cpp:832-925: white / red / green walls, a
// blue and a yellow metal box) from are::Triangle / are::Diffuse / are::Reflective / are::Texture objects in an
// are::ObjectSet, then lets the GPU do what are::Object::trace_texture is declared for: are::cuda::patch_render paints
// the scene seen through the two viewport triangles, are::cuda::trace_texture returns one mirror triangle's texture.
//
//   g++ -std=c++17 -O2 -Iinclude examples/patch_host.cpp -Laurora_rendering_engine_b200/lib -lare_b200 [continued]
//       -Wl,-rpath,$PWD/aurora_rendering_engine_b200/lib -o patch_host
//   ./patch_host out.ppm [width height [raw_double_dump [texture_dump]]]
#include <are_cuda.hpp>
#include <material/diffuse.h>
#include <material/reflective.h>
#include <object/object_set.h>
#include <texture.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

using namespace are;

struct Room {
	std::vector<std::unique_ptr<Texture>> textures;
	std::vector<std::unique_ptr<Material>> materials;
	std::vector<std::unique_ptr<Triangle>> triangles;
	ObjectSet set;

	Texture *colour(double r, double g, double b) {
		textures.emplace_back(new Texture(Texture::solid(Color3(r, g, b))));
		return textures.back().get();
	}
	template <typename M, typename... A>
	Material *material(A... a) {
		materials.emplace_back(new M(a...));
		return materials.back().get();
	}
	void tri(const Point3 &a, const Point3 &b, const Point3 &c, double u0, double v0, double u1, double v1, double u2, double v2, Material *m, Texture *t) {
		triangles.emplace_back(new Triangle(a, b - a, c - a, m, t));
		triangles.back()->set_uv(u0, v0, u1, v1, u2, v2);
		set.triangles.push_back(triangles.back().get());
	}
	// rt10.cpp:778-800: (p00, p10, p01) and (p10, p11, p01), uv (0,0)-(1,1) per quad
	void quad(const Point3 &p00, const Point3 &p10, const Point3 &p11, const Point3 &p01, Material *m, Texture *t) {
		tri(p00, p10, p01, 0, 0, 1, 0, 0, 1, m, t);
		tri(p10, p11, p01, 1, 0, 1, 1, 0, 1, m, t);
	}
	void box(const Point3 &lo, const Point3 &hi, Material *m, Texture *t) {  // rt10.cpp:802-830
		const Point3 p000(lo.x(), lo.y(), lo.z()), p001(lo.x(), lo.y(), hi.z()), p010(lo.x(), hi.y(), lo.z()), p011(lo.x(), hi.y(), hi.z());
		const Point3 p100(hi.x(), lo.y(), lo.z()), p101(hi.x(), lo.y(), hi.z()), p110(hi.x(), hi.y(), lo.z()), p111(hi.x(), hi.y(), hi.z());
		quad(p000, p001, p011, p010, m, t);
		quad(p100, p110, p111, p101, m, t);
		quad(p000, p100, p101, p001, m, t);
		quad(p010, p011, p111, p110, m, t);
		quad(p000, p010, p110, p100, m, t);
		quad(p001, p101, p111, p011, m, t);
	}
};

int main(int argc, char **argv) {
	const char *ppm = argc > 1 ? argv[1] : "patch.ppm";
	const int W = argc > 3 ? std::atoi(argv[2]) : 900, H = argc > 3 ? std::atoi(argv[3]) : 650;
	try {
		Room room;
		Material *diffuse = room.material<Diffuse>();
		Material *blue_metal = room.material<Reflective>(0.90), *yellow_metal = room.material<Reflective>(0.92);
		Texture *white = room.colour(0.85, 0.85, 0.85), *red = room.colour(0.85, 0.25, 0.25), *green = room.colour(0.25, 0.85, 0.25);
		Texture *blue = room.colour(0.25, 0.45, 1.0), *yellow = room.colour(1.0, 0.92, 0.20);
		room.quad(Point3(-1, 0, 0), Point3(1, 0, 0), Point3(1, 0, 2), Point3(-1, 0, 2), diffuse, white);
		room.quad(Point3(-1, 2, 0), Point3(-1, 2, 2), Point3(1, 2, 2), Point3(1, 2, 0), diffuse, white);
		room.quad(Point3(-1, 0, 2), Point3(1, 0, 2), Point3(1, 2, 2), Point3(-1, 2, 2), diffuse, white);
		room.quad(Point3(-1, 0, 0), Point3(-1, 0, 2), Point3(-1, 2, 2), Point3(-1, 2, 0), diffuse, red);
		room.quad(Point3(1, 0, 0), Point3(1, 2, 0), Point3(1, 2, 2), Point3(1, 0, 2), diffuse, green);
		room.box(Point3(-0.70, 0.0, 0.80), Point3(-0.15, 0.60, 1.30), blue_metal, blue);
		room.box(Point3(0.15, 0.0, 1.00), Point3(0.70, 1.10, 1.65), yellow_metal, yellow);

		// camera of rt10.cpp:895-921
		const Point3 eye(0.0, 1.0, -3.0);
		const double aspect = (double)W / (double)H, vp_h = 1.6, vp_w = vp_h * aspect, z = -2.0;
		const Point3 TL(0.0 - vp_w * 0.5, 1.0 - vp_h * 0.5, z), TR(0.0 + vp_w * 0.5, 1.0 - vp_h * 0.5, z);
		const Point3 BL(0.0 - vp_w * 0.5, 1.0 + vp_h * 0.5, z), BR(0.0 + vp_w * 0.5, 1.0 + vp_h * 0.5, z);
		const cuda::PatchViewport A = { { TL, TR, BL }, { { 0, 0 }, { 1, 0 }, { 0, 1 } } }, B = { { TR, BR, BL }, { { 1, 0 }, { 1, 1 }, { 0, 1 } } };
		cuda::PatchSettings cfg;
		cfg.max_depth = 4; cfg.min_area_px = 6.0; cfg.max_tex_res = 256; cfg.min_tex_res = 16;
		cfg.env = Color3(0.06, 0.07, 0.09);

		cuda::Renderer gpu;
		std::vector<std::uint8_t> rgb8;
		are_patch_stats st;
		Texture image = cuda::patch_render(gpu, room.set, eye, A, B, W, H, cfg, &rgb8, &st);
		if (are_cuda_write_ppm(ppm, W, H, rgb8.data()) != ARE_OK) throw std::runtime_error("cannot write the PPM file");
		std::printf("patch render %dx%d: %llu node textures (%llu texels), %llu warp triangles, %llu launches, plan %.3f ms, kernels %.3f ms\n", W, H,
			(unsigned long long)st.nodes, (unsigned long long)st.node_texels, (unsigned long long)st.ops, (unsigned long long)st.launches, st.plan_ms, st.kernel_ms);
		if (argc > 4) {
			FILE *f = std::fopen(argv[4], "wb");
			if (f) { std::fwrite(image.data(), sizeof(double), (size_t)W * H * 3, f); std::fclose(f); }
		}
		if (argc > 5) {  // Object::trace_texture of the blue box's front face (-Z face, first triangle) seen from the eye
			const Triangle &front = *room.triangles[10 + 8];
			Texture t = cuda::trace_texture(gpu, room.set, front, eye, 200, 200, cfg);
			FILE *f = std::fopen(argv[5], "wb");
			if (f) { std::fwrite(t.data(), sizeof(double), (size_t)t.width_ * t.height_ * 3, f); std::fclose(f); }
			std::printf("trace_texture: %dx%d\n", t.width_, t.height_);
		}
	} catch (const std::exception &e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
