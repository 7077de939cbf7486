// oracle/ref_patch_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A flat extern "C" face over the REAL reference patch renderer, experiments/rt10.cpp (the cleanest of the
// reference's "patch-as-viewport" prototypes; its shipped image experiments/output_rt10.ppm reproduces bit-exactly,
// SURVEY.md §4).  oracle/Makefile compiles this file with -I/root/reference so that the reference translation unit
// is #included where it lies (its main() renamed by the preprocessor); nothing of it is copied into the repository.
// Output: oracle/_ref/librt10_ref.so.  Used to pin oracle/patch_oracle.cpp on arbitrary scenes and to record
// tests/golden/patch_vectors.npz.
//
// Reference routines reached (experiments/rt10.cpp):
//   Camera::render / renderViewportTriangle   :677-772
//   renderTriangleWithTriangle                :551-664
//   collectCandidateTriangles                 :462-485
//   clipPolygonToTriangle / rasterizeWarp*    :348-460
//   Image::sampleBilinear / writePPM          :98-143
#define main are_rt10_reference_main
#include <experiments/rt10.cpp>
#undef main

#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct RefScene {
	std::vector<Material> mats;
	std::vector<Triangle> tris;
};

void build_scene(RefScene &s, int n_tri, const double *P, const double *UV, const int *material, int n_mat, const int *mat_type,
	const double *mat_albedo, const double *mat_metalness) {
	s.mats.resize(n_mat);
	for (int i = 0; i < n_mat; ++i) {
		s.mats[i].type = mat_type[i] ? MaterialType::Reflective : MaterialType::Diffuse;
		s.mats[i].albedo = Vec3(mat_albedo[3 * i], mat_albedo[3 * i + 1], mat_albedo[3 * i + 2]);
		s.mats[i].metalness = mat_metalness[i];
	}
	s.tris.resize(n_tri);
	for (int i = 0; i < n_tri; ++i) {
		Triangle &t = s.tris[i];
		t.id = i + 1;  // rt10.cpp:780 numbers scene triangles from 1
		for (int k = 0; k < 3; ++k) {
			t.p[k] = Vec3(P[9 * i + 3 * k], P[9 * i + 3 * k + 1], P[9 * i + 3 * k + 2]);
			t.uv[k] = Vec2(UV[6 * i + 2 * k], UV[6 * i + 2 * k + 1]);
		}
		t.mat = (material[i] >= 0 && material[i] < n_mat) ? &s.mats[material[i]] : nullptr;
	}
}

Triangle viewport_tri(int id, const double *P, const double *UV) {
	Triangle t;
	t.id = id;
	for (int k = 0; k < 3; ++k) {
		t.p[k] = Vec3(P[3 * k], P[3 * k + 1], P[3 * k + 2]);
		t.uv[k] = Vec2(UV[2 * k], UV[2 * k + 1]);
	}
	t.mat = nullptr;
	return t;
}

RenderConfig make_cfg(const double *cfg) {
	RenderConfig c;
	c.maxDepth = (int)cfg[0];
	c.minAreaPxToRecurse = cfg[1];
	c.maxTriTexRes = (int)cfg[2];
	c.minTriTexRes = (int)cfg[3];
	c.envColor = Vec3(cfg[4], cfg[5], cfg[6]);
	return c;
}

// Image::writePPM's per-pixel arithmetic (rt10.cpp:118-143) without the file: the harness cannot call writePPM into
// memory, so the P6 payload is produced by writing to a temporary file through the reference's own routine.
void image_to_rgb8(const Image &img, double gamma, uint8_t *out) {
	char path[] = "/tmp/are_rt10_XXXXXX";
	int fd = mkstemp(path);
	if (fd >= 0) close(fd);
	img.writePPM(path, gamma);
	FILE *f = std::fopen(path, "rb");
	if (f) {
		int w = 0, h = 0, maxv = 0;
		if (std::fscanf(f, "P6 %d %d %d", &w, &h, &maxv) == 3 && std::fgetc(f) == '\n') {
			size_t got = std::fread(out, 1, (size_t)w * h * 3, f);
			(void)got;
		}
		std::fclose(f);
	}
	std::remove(path);
}

}  // namespace

extern "C" {

// cfg = { maxDepth, minAreaPxToRecurse, maxTriTexRes, minTriTexRes, env.r, env.g, env.b, gamma }
int ref_patch_render(int n_tri, const double *P, const double *UV, const int *material, int n_mat, const int *mat_type, const double *mat_albedo,
	const double *mat_metalness, const double *origin, const double *vp_P, const double *vp_UV, int width, int height, const double *cfg,
	double *out_rgb, uint8_t *out_rgb8) {
	try {
		RefScene s;
		build_scene(s, n_tri, P, UV, material, n_mat, mat_type, mat_albedo, mat_metalness);
		Camera cam;
		cam.origin = Vec3(origin[0], origin[1], origin[2]);
		cam.width = width;
		cam.height = height;
		cam.vpA = viewport_tri(-100, vp_P, vp_UV);
		cam.vpB = viewport_tri(-101, vp_P + 9, vp_UV + 6);
		Image img = cam.render(s.tris, make_cfg(cfg));
		if (out_rgb)
			for (size_t i = 0; i < img.pix.size(); ++i) {
				out_rgb[3 * i] = img.pix[i].x;
				out_rgb[3 * i + 1] = img.pix[i].y;
				out_rgb[3 * i + 2] = img.pix[i].z;
			}
		if (out_rgb8) image_to_rgb8(img, cfg[7], out_rgb8);
		return 0;
	} catch (...) {
		return -1;
	}
}

// renderTriangleWithTriangle (rt10.cpp:551-664) for scene triangle `current` seen from `origin`, as a root call
// (depth 0, empty recursion stack).  out_tex = clamped tex_w x tex_h x 3 doubles; returns the clamped size in out_wh.
int ref_patch_trace_texture(int n_tri, const double *P, const double *UV, const int *material, int n_mat, const int *mat_type,
	const double *mat_albedo, const double *mat_metalness, const double *origin, int current, int tex_w, int tex_h, double est_area_px,
	const double *cfg, double *out_tex, int *out_wh) {
	try {
		RefScene s;
		build_scene(s, n_tri, P, UV, material, n_mat, mat_type, mat_albedo, mat_metalness);
		std::vector<int> stack;
		Image img = renderTriangleWithTriangle(s.tris, Vec3(origin[0], origin[1], origin[2]), s.tris[current], tex_w, tex_h, 0, est_area_px, stack,
			make_cfg(cfg));
		out_wh[0] = img.w;
		out_wh[1] = img.h;
		for (size_t i = 0; i < img.pix.size(); ++i) {
			out_tex[3 * i] = img.pix[i].x;
			out_tex[3 * i + 1] = img.pix[i].y;
			out_tex[3 * i + 2] = img.pix[i].z;
		}
		return 0;
	} catch (...) {
		return -1;
	}
}

// The reference program itself: renders its built-in Cornell box and writes ./output_rt10.ppm (rt10.cpp:832-933).
int ref_rt10_main(void) { return are_rt10_reference_main(); }

}  // extern "C"
