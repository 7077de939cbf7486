// <texture.h> — are::Texture: a W x H image of Color3 (fp64), row-major image_[y][x].
//
// Public surface of the reference class (reference include/texture.h:10-32, src/texture.cpp): construct from a binary
// PPM (P6, maxval 255 only; channel = byte / 255.0) or by filling with a colour, pixel(x, y) hands out a mutable
// reference, save_texture() writes P6 with (unsigned char)clamp(c * 255, 0, 255) — truncation, no gamma — and refuses
// anything that does not end in ".ppm".  Failures throw std::runtime_error exactly where the reference does.
//
// Differences, all additive:
//   * value semantics are safe (the reference returns Texture by value but has no copy constructor and leaks the row
//     table in its destructor, src/texture.cpp:68-75); storage here is one contiguous vector.
//   * procedural kinds for the path tracer (solid / uv checker / spatial checker / Perlin noise) are described by a tag
//     and up to eight parameters; they carry no pixels and are evaluated on the GPU.
//   * Texture::paste (the homography warp of the reference's patch renderer, src/texture.cpp:85-360) is not part of
//     the path-tracing hot path and is not provided.
#pragma once

#include <basic/vec3.h>

#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace are {

class Texture {
public:
	enum Kind { IMAGE = 4, SOLID = 0, CHECKER_UV = 1, CHECKER_3D = 2, NOISE = 3 };  // values = are_texture_kind

	Texture() = default;

	explicit Texture(const std::string &image_path) {
		FILE *fp = std::fopen(image_path.c_str(), "rb");
		if (!fp) throw std::runtime_error("Failed to open texture file: " + image_path);
		char magic[3] = { 0, 0, 0 };
		if (std::fscanf(fp, "%2s", magic) != 1 || std::string(magic) != "P6") {
			std::fclose(fp);
			throw std::runtime_error("Unsupported PPM format in file: " + image_path);
		}
		int w = 0, h = 0, maxval = 0;
		if (std::fscanf(fp, "%d %d %d", &w, &h, &maxval) != 3 || maxval != 255) {
			std::fclose(fp);
			throw std::runtime_error("Invalid PPM header in file: " + image_path);
		}
		std::fgetc(fp);  // the single whitespace byte after maxval
		width_ = w;
		height_ = h;
		texels_.resize(static_cast<size_t>(w) * h);
		for (Color3 &px : texels_) {
			unsigned char rgb[3];
			if (std::fread(rgb, 1, 3, fp) != 3) {
				std::fclose(fp);
				throw std::runtime_error("Unexpected end of file while reading pixel data in file: " + image_path);
			}
			px = Color3(rgb[0] / 255.0, rgb[1] / 255.0, rgb[2] / 255.0);
		}
		std::fclose(fp);
	}

	Texture(int width, int height, const Color3 &fill_color) : width_(width), height_(height) {
		if (width <= 0 || height <= 0) throw std::runtime_error("Texture width and height must be positive.");
		texels_.assign(static_cast<size_t>(width) * height, fill_color);
		// a filled image is also usable as a solid colour by the path tracer
		params_[0] = fill_color.x();
		params_[1] = fill_color.y();
		params_[2] = fill_color.z();
		uniform_fill_ = true;
	}

	// ---- procedural kinds (additive) -------------------------------------------------------------------------
	static Texture solid(const Color3 &c) { return Texture(1, 1, c); }
	/// (floor(u*scale) + floor(v*scale)) even -> even colour, on the surface's (u,v)   [experiments/rt.cpp:93-102]
	static Texture checker_uv(double scale, const Color3 &even, const Color3 &odd) { return procedural(CHECKER_UV, scale, even, odd); }
	/// (floor(x/scale) + floor(y/scale) + floor(z/scale)) even -> even colour, on the hit point
	static Texture checker_3d(double scale, const Color3 &even, const Color3 &odd) { return procedural(CHECKER_3D, scale, even, odd); }
	/// 0.5 * (1 + sin(scale * z + 10 * turbulence(p, 7))) marble; tables drawn from Philox(seed)
	static Texture noise(double scale, unsigned seed) {
		Texture t;
		t.kind_ = NOISE;
		t.params_[0] = scale;
		t.params_[1] = seed;
		return t;
	}

	Color3 &pixel(int x, int y) {
		if (texels_.empty()) throw std::runtime_error("Texture is not initialized.");
		uniform_fill_ = false;  // the caller may write through the reference
		return texels_[static_cast<size_t>(y) * width_ + x];
	}
	const Color3 &pixel(int x, int y) const {
		if (texels_.empty()) throw std::runtime_error("Texture is not initialized.");
		return texels_[static_cast<size_t>(y) * width_ + x];
	}

	bool save_texture(const std::string &file_path) const {
		if (texels_.empty()) throw std::runtime_error("Texture is not initialized.");
		if (file_path.size() < 4 || file_path.compare(file_path.size() - 4, 4, ".ppm") != 0) return false;
		FILE *fp = std::fopen(file_path.c_str(), "wb");
		if (!fp) return false;
		std::fprintf(fp, "P6\n%d %d\n255\n", width_, height_);
		std::vector<unsigned char> row(static_cast<size_t>(width_) * 3);
		for (int y = 0; y < height_; ++y) {
			for (int x = 0; x < width_; ++x) {
				const Color3 &c = texels_[static_cast<size_t>(y) * width_ + x];
				for (int k = 0; k < 3; ++k) row[3 * x + k] = static_cast<unsigned char>(std::clamp(c[k] * 255.0, 0.0, 255.0));
			}
			std::fwrite(row.data(), 1, row.size(), fp);
		}
		std::fclose(fp);
		return true;
	}

	int width_ = 0, height_ = 0;

	// ---- read access for the CUDA shim (additive) ------------------------------------------------------------
	/// are_texture_kind this texture uploads as: a uniformly filled image is a SOLID colour, any other image IMAGE.
	int kind() const { return kind_ == IMAGE ? (uniform_fill_ ? SOLID : IMAGE) : kind_; }
	const double *params() const { return params_; }
	/// width_*height_*3 doubles, row-major — the layout are_cuda_add_texture expects.
	const double *data() const { return texels_.empty() ? nullptr : texels_.front().e(); }

private:
	static Texture procedural(Kind k, double scale, const Color3 &even, const Color3 &odd) {
		Texture t;
		t.kind_ = k;
		t.params_[0] = scale;
		for (int i = 0; i < 3; ++i) {
			t.params_[1 + i] = even[i];
			t.params_[4 + i] = odd[i];
		}
		return t;
	}

	std::vector<Color3> texels_;
	Kind kind_ = IMAGE;
	double params_[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	bool uniform_fill_ = false;
};

}  // namespace are
