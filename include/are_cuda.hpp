// <are_cuda.hpp> — header-only C++ shim between the are:: class API and the C ABI of <are_cuda.h>.
//
// A host program written in the reference's style builds are::Texture / are::Material / are::Triangle objects, pushes
// pointers into an are::ObjectSet, and then — instead of looping over pixels on the CPU — hands the set to
// are::cuda::Renderer:
//
//     are::cuda::Renderer gpu;                     // one CUDA device, throws std::runtime_error if there is none
//     gpu.upload(scene);                           // walks ObjectSet::triangles (+ quads, spheres), commits to the GPU
//     are::Texture image = gpu.render(camera, 512, 512, settings);
//     image.save_texture("out.ppm");               // the reference's own writer semantics
//
// Errors follow the reference's convention: what the C ABI reports as ARE_ERR_INVALID_ARGUMENT is re-thrown as
// std::invalid_argument (are::Triangle's constructor checks), everything else as std::runtime_error.
#pragma once

#include <are_cuda.h>
#include <camera.h>
#include <material/reflective.h>
#include <object/object_set.h>
#include <texture.h>

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace are {
namespace cuda {

struct Settings {
	int spp = 64;
	int max_depth = 50;
	std::uint64_t seed = 1;
	double t_min = 1e-3;
	Color3 background_bottom = Color3(1, 1, 1);
	Color3 background_top = Color3(0.5, 0.7, 1.0);
	int integrator = ARE_INTEGRATOR_PATH;  // ARE_INTEGRATOR_RT_AO = the shading of experiments/rt.cpp
	int traversal = ARE_TRAVERSAL_AUTO;
	int ao_samples = 32;
	int sample_begin = 0;  // first global sample index (sample-sharded multi-GPU jobs)
};

class Renderer {
public:
	explicit Renderer(int device = 0) {
		if (are_cuda_create(&ctx_, device) != ARE_OK) throw std::runtime_error(std::string("are_cuda_create: ") + are_cuda_last_error(nullptr));
	}
	/// Several GPUs of one box behind one renderer: every render() / render_sums() is sharded by sample range over the
	/// devices and summed on devices[0] over NVLink peer memory (are_cuda_create_multi).
	explicit Renderer(const std::vector<int> &devices) {
		if (are_cuda_create_multi(&ctx_, devices.data(), static_cast<int>(devices.size())) != ARE_OK)
			throw std::runtime_error(std::string("are_cuda_create_multi: ") + are_cuda_last_error(nullptr));
	}
	/// All visible GPUs.
	static std::vector<int> all_devices() {
		std::vector<int> d(static_cast<size_t>(are_cuda_device_count() > 0 ? are_cuda_device_count() : 0));
		for (size_t i = 0; i < d.size(); ++i) d[i] = static_cast<int>(i);
		return d;
	}
	int device_count() const {
		int n = 1;
		are_cuda_group_info(ctx_, &n, nullptr);
		return n;
	}
	~Renderer() { are_cuda_destroy(ctx_); }
	Renderer(const Renderer &) = delete;
	Renderer &operator=(const Renderer &) = delete;

	/// Flatten the scene and upload it. May be called again after the scene changed.
	void upload(const ObjectSet &set) {
		check(are_cuda_clear(ctx_));
		textures_.clear();
		materials_.clear();
		for (const Triangle *t : set.triangles) {
			const int mat = material_id(t->material()), tex = texture_id(t->texture());
			const int id = check(are_cuda_add_triangle(ctx_, t->origin().e(), t->edge_u().e(), t->edge_v().e(), mat, tex));
			if (t->has_uv()) check(are_cuda_set_triangle_uv(ctx_, id, t->uv()));
		}
		for (const Quad *q : set.quads)
			check(are_cuda_add_quad(ctx_, q->origin().e(), q->edge_u().e(), q->edge_v().e(), material_id(q->material()), texture_id(q->texture())));
		for (const Sphere *s : set.spheres)
			check(are_cuda_add_sphere(ctx_, s->center().e(), s->radius(), material_id(s->material()), texture_id(s->texture())));
		std::uint64_t bytes = 0;
		check(are_cuda_commit(ctx_, &bytes));
		uploaded_bytes_ = bytes;
	}

	/// Render spp samples per pixel; the returned texture holds the mean linear radiance per pixel.
	Texture render(const Camera &camera, int width, int height, const Settings &s = Settings()) {
		std::vector<float> sums(static_cast<size_t>(width) * height * 3);
		render_sums(camera, width, height, s, sums.data());
		Texture out(width, height, Color3(0, 0, 0));
		const double inv = s.spp > 0 ? 1.0 / s.spp : 0.0;
		for (int y = 0; y < height; ++y)
			for (int x = 0; x < width; ++x) {
				const float *p = &sums[(static_cast<size_t>(y) * width + x) * 3];
				out.pixel(x, y) = Color3(p[0] * inv, p[1] * inv, p[2] * inv);
			}
		return out;
	}

	/// Raw per-pixel sample SUMS (W*H*3 floats) — what a multi-GPU driver adds up across devices.
	void render_sums(const Camera &camera, int width, int height, const Settings &s, float *sums) {
		are_camera c = to_c(camera);
		are_render_params p = to_c(width, height, s);
		check(are_cuda_render(ctx_, &c, &p, sums, &stats_));
	}

	/// Who builds the hierarchy at the next upload(): ARE_BVH_BUILDER_HOST_SAH (default) or ARE_BVH_BUILDER_DEVICE_LBVH
	/// (milliseconds per million primitives; for scenes that change every frame).
	void set_bvh_builder(int builder) { check(are_cuda_set_bvh_builder(ctx_, builder)); }
	/// Context options (are_option in are_cuda.h: kernel choice, scene-compiler switches).
	void set_option(int option, int value) { check(are_cuda_set_option(ctx_, option, value)); }
	/// Triangles of the uploaded set moved (same objects, new geometry): hand the ones that changed to move(), then refit().
	/// refit() returns false when the hierarchy cannot be refitted (host-built tree, fused parallelograms / boxes) — call
	/// upload() again in that case.  Triangle ids are positions in ObjectSet::triangles as uploaded.
	void move(int triangle_id, const Triangle &t) {
		check(are_cuda_update_triangles(ctx_, 1, &triangle_id, t.origin().e(), t.edge_u().e(), t.edge_v().e()));
	}
	bool refit(double *device_ms = nullptr) {
		const int st = are_cuda_refit(ctx_, device_ms);
		if (st == ARE_ERR_RUNTIME) return false;
		check(st);
		return true;
	}
	are_commit_info commit_info() {
		are_commit_info info;
		check(are_cuda_get_commit_info(ctx_, &info));
		return info;
	}

	const are_render_stats &stats() const { return stats_; }
	std::uint64_t uploaded_bytes() const { return uploaded_bytes_; }
	are_cuda_ctx *context() { return ctx_; }

	static are_camera to_c(const Camera &cam) {
		are_camera c;
		for (int i = 0; i < 3; ++i) {
			c.pos[i] = cam.pos[i];
			c.target[i] = cam.target[i];
			c.up[i] = cam.up[i];
		}
		c.vfov_deg = cam.vfov_deg;
		c.focus_dist = cam.focus_dist;
		c.defocus_angle_deg = cam.defocus_angle_deg;
		c.jitter = cam.jitter ? 1 : 0;
		c.pad_ = 0;
		return c;
	}
	static are_render_params to_c(int width, int height, const Settings &s) {
		are_render_params p;
		p.width = width;
		p.height = height;
		p.sample_begin = s.sample_begin;
		p.sample_count = s.spp;
		p.max_depth = s.max_depth;
		p.integrator = s.integrator;
		p.traversal = s.traversal;
		p.ao_samples = s.ao_samples;
		p.seed = s.seed;
		p.t_min = s.t_min;
		for (int i = 0; i < 3; ++i) {
			p.background_bottom[i] = s.background_bottom[i];
			p.background_top[i] = s.background_top[i];
		}
		return p;
	}

private:
	int check(int status) {
		if (status >= 0) return status;
		const std::string msg = are_cuda_last_error(ctx_);
		if (status == ARE_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
		throw std::runtime_error(msg);
	}
	int texture_id(const Texture *t) {
		if (!t) throw std::invalid_argument("Texture pointer cannot be null");
		auto it = textures_.find(t);
		if (it != textures_.end()) return it->second;
		const int kind = t->kind();
		const int id = kind == ARE_TEX_IMAGE ? check(are_cuda_add_texture(ctx_, kind, t->params(), t->data(), t->width_, t->height_))
											 : check(are_cuda_add_texture(ctx_, kind, t->params(), nullptr, 0, 0));
		textures_[t] = id;
		return id;
	}
	int material_id(const Material *m) {
		if (!m) throw std::invalid_argument("Material pointer cannot be null");
		auto it = materials_.find(m);
		if (it != materials_.end()) return it->second;
		double p[8];
		m->describe(p);
		const int kind = m->kind();
		const Texture *over = m->texture_override();
		const double over_id = over ? texture_id(over) : -1.0;
		if (kind == ARE_MAT_LAMBERTIAN || kind == ARE_MAT_DIFFUSE_LIGHT) p[0] = over_id;
		else if (kind == ARE_MAT_METAL) p[1] = over_id;
		const int id = check(are_cuda_add_material(ctx_, kind, p));
		materials_[m] = id;
		return id;
	}

	are_cuda_ctx *ctx_ = nullptr;
	std::map<const Texture *, int> textures_;
	std::map<const Material *, int> materials_;
	are_render_stats stats_{};
	std::uint64_t uploaded_bytes_ = 0;
};

/// are::Texture::paste on the GPU (bit-identical to the CPU member function): warps `src` into the quad of `dst`.
inline void paste(Renderer &gpu, Texture &dst, const Texture &src, const std::pair<int, int> &left_top, const std::pair<int, int> &right_top,
	const std::pair<int, int> &left_bottom, const std::pair<int, int> &right_bottom) {
	if (dst.width_ <= 0 || dst.height_ <= 0 || !dst.data()) throw std::runtime_error("Texture is not initialized.");
	const int corners[8] = { left_top.first, left_top.second, right_top.first, right_top.second, left_bottom.first, left_bottom.second,
		right_bottom.first, right_bottom.second };
	double *d = dst.pixel(0, 0).e();  // contiguous w*h*3 doubles
	const int st = are_cuda_texture_paste(gpu.context(), d, dst.width_, dst.height_, src.data(), src.width_, src.height_, corners);
	if (st < 0) throw std::runtime_error(are_cuda_last_error(gpu.context()));
}

// ---- the reference's own algorithm: patch-as-viewport rendering ------------------------------------------------
// are::Object::trace_texture(object_set, viewport_origin) is what the reference declares for it
// (include/object/object.h:37-38) and never defines; experiments/rt10.cpp is its prototype.  These wrappers map the
// class API onto are_cuda_patch_*: a triangle's vertices are get_vertices(), its uv are Triangle::uv(), its material is
// Diffuse (type 0) or Reflective (type 1, metalness = reflectivity_), its albedo the colour of its (solid) Texture.
struct PatchSettings {  // RenderConfig, rt10.cpp:536-542
	int max_depth = 4;
	double min_area_px = 4.0;
	int max_tex_res = 256, min_tex_res = 16;
	Color3 env = Color3(0.08, 0.08, 0.10);
	double gamma = 2.2;
};
struct PatchViewport {  // one viewport triangle of the camera (rt10.cpp:899-915)
	Point3 p[3];
	double uv[3][2];
};

class PatchScene {
public:
	explicit PatchScene(const ObjectSet &set) {
		std::map<std::pair<const Material *, const Texture *>, int> ids;
		for (const Triangle *t : set.triangles) {
			const std::vector<Point3> &v = t->get_vertices();
			for (int k = 0; k < 3; ++k)
				for (int c = 0; c < 3; ++c) P_.push_back(v[k][c]);
			UV_.insert(UV_.end(), t->uv(), t->uv() + 6);
			const auto key = std::make_pair(static_cast<const Material *>(t->material()), static_cast<const Texture *>(t->texture()));
			auto it = ids.find(key);
			if (it == ids.end()) {
				it = ids.insert({ key, static_cast<int>(type_.size()) }).first;
				const Reflective *mirror = dynamic_cast<const Reflective *>(t->material());
				type_.push_back(mirror ? 1 : 0);
				metal_.push_back(mirror ? mirror->reflectivity_ : 0.0);
				const Texture *tex = t->texture();
				const Color3 c = tex->data() ? static_cast<const Texture &>(*tex).pixel(0, 0) : Color3(tex->params()[0], tex->params()[1], tex->params()[2]);
				for (int k = 0; k < 3; ++k) albedo_.push_back(c[k]);
			}
			material_.push_back(it->second);
			index_[t] = static_cast<int>(material_.size()) - 1;
		}
		c_.n_tri = static_cast<int32_t>(material_.size());
		c_.n_mat = static_cast<int32_t>(type_.size());
		c_.P = P_.data(); c_.UV = UV_.data(); c_.material = material_.data();
		c_.mat_type = type_.data(); c_.mat_albedo = albedo_.data(); c_.mat_metalness = metal_.data();
	}
	PatchScene(const PatchScene &) = delete;
	PatchScene &operator=(const PatchScene &) = delete;
	const are_patch_scene *c() const { return &c_; }
	int index_of(const Triangle &t) const {
		auto it = index_.find(&t);
		if (it == index_.end()) throw std::invalid_argument("triangle is not part of the object set");
		return it->second;
	}
	static are_patch_config to_c(const PatchSettings &s) {
		are_patch_config c;
		c.max_depth = s.max_depth; c.max_tex_res = s.max_tex_res; c.min_tex_res = s.min_tex_res; c.pad_ = 0;
		c.min_area_px = s.min_area_px;
		for (int k = 0; k < 3; ++k) c.env[k] = s.env[k];
		c.gamma = s.gamma;
		return c;
	}

private:
	std::vector<double> P_, UV_, albedo_, metal_;
	std::vector<int32_t> material_, type_;
	std::map<const Triangle *, int> index_;
	are_patch_scene c_{};
};

/// Object::trace_texture on the GPU: the texture of `current` (a triangle of `set`) seen from `viewport_origin`, with
/// everything visible through it painted in (renderTriangleWithTriangle, rt10.cpp:551-664).  The size is clamped to
/// [min_tex_res, max_tex_res] as the reference clamps it.
inline Texture trace_texture(Renderer &gpu, const ObjectSet &set, const Triangle &current, const Point3 &viewport_origin, int tex_w, int tex_h,
	const PatchSettings &s = PatchSettings()) {
	PatchScene scene(set);
	const are_patch_config cfg = PatchScene::to_c(s);
	std::vector<double> tex(static_cast<size_t>(s.max_tex_res) * s.max_tex_res * 3);
	int wh[2] = { 0, 0 };
	const int st = are_cuda_patch_trace_texture(gpu.context(), scene.c(), viewport_origin.e(), scene.index_of(current), tex_w, tex_h, 0.0, &cfg, tex.data(),
		wh, nullptr);
	if (st == ARE_ERR_INVALID_ARGUMENT) throw std::invalid_argument(are_cuda_last_error(gpu.context()));
	if (st < 0) throw std::runtime_error(are_cuda_last_error(gpu.context()));
	Texture out(wh[0], wh[1], Color3(0, 0, 0));
	for (int y = 0; y < wh[1]; ++y)
		for (int x = 0; x < wh[0]; ++x) {
			const double *p = &tex[(static_cast<size_t>(y) * wh[0] + x) * 3];
			out.pixel(x, y) = Color3(p[0], p[1], p[2]);
		}
	return out;
}

/// Camera::render of rt10.cpp:755-772: the scene seen from `origin` through the two viewport triangles.  Returns the
/// linear image; rgb8 (optional, W*H*3) receives the gamma-encoded P6 payload of Image::writePPM (rt10.cpp:118-143).
inline Texture patch_render(Renderer &gpu, const ObjectSet &set, const Point3 &origin, const PatchViewport &a, const PatchViewport &b, int width, int height,
	const PatchSettings &s = PatchSettings(), std::vector<std::uint8_t> *rgb8 = nullptr, are_patch_stats *stats = nullptr) {
	PatchScene scene(set);
	const are_patch_config cfg = PatchScene::to_c(s);
	double vp_P[18], vp_UV[12];
	const PatchViewport *vp[2] = { &a, &b };
	for (int v = 0; v < 2; ++v)
		for (int k = 0; k < 3; ++k) {
			for (int c = 0; c < 3; ++c) vp_P[9 * v + 3 * k + c] = vp[v]->p[k][c];
			vp_UV[6 * v + 2 * k] = vp[v]->uv[k][0];
			vp_UV[6 * v + 2 * k + 1] = vp[v]->uv[k][1];
		}
	std::vector<double> rgb(static_cast<size_t>(width) * height * 3);
	if (rgb8) rgb8->resize(rgb.size());
	const int st = are_cuda_patch_render(gpu.context(), scene.c(), origin.e(), vp_P, vp_UV, width, height, &cfg, rgb.data(), rgb8 ? rgb8->data() : nullptr, stats);
	if (st == ARE_ERR_INVALID_ARGUMENT) throw std::invalid_argument(are_cuda_last_error(gpu.context()));
	if (st < 0) throw std::runtime_error(are_cuda_last_error(gpu.context()));
	Texture out(width, height, Color3(0, 0, 0));
	for (int y = 0; y < height; ++y)
		for (int x = 0; x < width; ++x) {
			const double *p = &rgb[(static_cast<size_t>(y) * width + x) * 3];
			out.pixel(x, y) = Color3(p[0], p[1], p[2]);
		}
	return out;
}

}  // namespace cuda
}  // namespace are
