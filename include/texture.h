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
//   * Texture::paste (the homography warp of the reference's patch renderer, src/texture.cpp:85-360) is provided with the
//     same arithmetic (bit-identical texels); its GPU form is are_cuda_texture_paste / are::cuda::paste.
#pragma once

#include <basic/vec3.h>

#include <algorithm>
#include <cmath>
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

	/// Warp src_texture into the quadrilateral (left_top, right_top, left_bottom, right_bottom) of this image: the
	/// homography that maps src's corner pixel centres (0,0), (w-1,0), (0,h-1), (w-1,h-1) onto the four points; a
	/// destination pixel is overwritten when its centre lies inside the quad (even-odd rule) and its pre-image lies
	/// inside src; the value is the bilinear blend of the four surrounding source texels.
	void paste(const Texture &src_texture, const std::pair<int, int> &left_top, const std::pair<int, int> &right_top,
		const std::pair<int, int> &left_bottom, const std::pair<int, int> &right_bottom) {
		if (texels_.empty()) throw std::runtime_error("Texture is not initialized.");
		const int sw_i = src_texture.width_, sh_i = src_texture.height_;
		if (sw_i <= 0 || sh_i <= 0) return;
		uniform_fill_ = false;
		constexpr double eps = 1e-12;  // GEOMETRY_EPSILON
		// walk order for the inside test: LT, RT, RB, LB
		const double qx[4] = { double(left_top.first), double(right_top.first), double(right_bottom.first), double(left_bottom.first) };
		const double qy[4] = { double(left_top.second), double(right_top.second), double(right_bottom.second), double(left_bottom.second) };
		const double lo_x = std::min({ qx[0], qx[1], qx[2], qx[3] }), hi_x = std::max({ qx[0], qx[1], qx[2], qx[3] });
		const double lo_y = std::min({ qy[0], qy[1], qy[2], qy[3] }), hi_y = std::max({ qy[0], qy[1], qy[2], qy[3] });
		if (hi_x < 0.0 || hi_y < 0.0 || lo_x > double(width_ - 1) || lo_y > double(height_ - 1)) return;
		const int x_first = std::max(0, int(std::floor(lo_x))), x_last = std::min(width_ - 1, int(std::ceil(hi_x)));
		const int y_first = std::max(0, int(std::floor(lo_y))), y_last = std::min(height_ - 1, int(std::ceil(hi_y)));
		// forward homography: 8 unknowns from 4 point pairs, augmented system reduced by Gauss-Jordan with partial pivoting
		const double sw = double(sw_i), sh = double(sh_i);
		const double from[4][2] = { { 0.0, 0.0 }, { sw - 1.0, 0.0 }, { 0.0, sh - 1.0 }, { sw - 1.0, sh - 1.0 } };
		const double to[4][2] = { { qx[0], qy[0] }, { qx[1], qy[1] }, { qx[3], qy[3] }, { qx[2], qy[2] } };
		double m[8][9];
		for (int k = 0; k < 4; ++k) {
			const double x = from[k][0], y = from[k][1], u = to[k][0], v = to[k][1];
			const double row_u[9] = { x, y, 1.0, 0.0, 0.0, 0.0, -u * x, -u * y, u };
			const double row_v[9] = { 0.0, 0.0, 0.0, x, y, 1.0, -v * x, -v * y, v };
			std::copy(row_u, row_u + 9, m[2 * k]);
			std::copy(row_v, row_v + 9, m[2 * k + 1]);
		}
		for (int col = 0; col < 8; ++col) {
			int lead = col;
			for (int r = col + 1; r < 8; ++r)
				if (std::fabs(m[r][col]) > std::fabs(m[lead][col])) lead = r;
			if (std::fabs(m[lead][col]) < eps) return;  // degenerate corner configuration
			if (lead != col) std::swap_ranges(m[col] + col, m[col] + 9, m[lead] + col);
			const double scale = m[col][col];
			for (int c = col; c < 9; ++c) m[col][c] /= scale;
			for (int r = 0; r < 8; ++r) {
				const double k = m[r][col];
				if (r == col || std::fabs(k) < eps) continue;
				for (int c = col; c < 9; ++c) m[r][c] -= k * m[col][c];
			}
		}
		// inverse of [[a b c][d e f][g h 1]] by cofactors
		const double a = m[0][8], b = m[1][8], c = m[2][8], d = m[3][8], e = m[4][8], f = m[5][8], g = m[6][8], h = m[7][8], one = 1.0;
		const double c00 = (e * one - f * h), c01 = -(d * one - f * g), c02 = (d * h - e * g);
		const double c10 = -(b * one - c * h), c11 = (a * one - c * g), c12 = -(a * h - b * g);
		const double c20 = (b * f - c * e), c21 = -(a * f - c * d), c22 = (a * e - b * d);
		const double det = a * c00 + b * c01 + c * c02;
		if (std::fabs(det) < eps) return;
		const double rdet = 1.0 / det;
		const double inv[3][3] = { { c00 * rdet, c10 * rdet, c20 * rdet }, { c01 * rdet, c11 * rdet, c21 * rdet }, { c02 * rdet, c12 * rdet, c22 * rdet } };
		for (int y = y_first; y <= y_last; ++y)
			for (int x = x_first; x <= x_last; ++x) {
				const double px = double(x) + 0.5, py = double(y) + 0.5;
				bool inside = false;
				for (int i = 0, j = 3; i < 4; j = i++) {
					const double rise = qy[j] - qy[i];
					if (((qy[i] > py) != (qy[j] > py)) && (px < (qx[j] - qx[i]) * (py - qy[i]) / (rise == 0.0 ? 1e-30 : rise) + qx[i])) inside = !inside;
				}
				if (!inside) continue;
				const double w = inv[2][0] * px + inv[2][1] * py + inv[2][2];
				if (std::fabs(w) < eps) continue;
				double u = (inv[0][0] * px + inv[0][1] * py + inv[0][2]) / w, v = (inv[1][0] * px + inv[1][1] * py + inv[1][2]) / w;
				if (u < 0.0 || v < 0.0 || u > sw - 1.0 || v > sh - 1.0) continue;
				u = std::clamp(u, 0.0, sw - 1.0);
				v = std::clamp(v, 0.0, sh - 1.0);
				const int iu = int(std::floor(u)), iv = int(std::floor(v));
				const int iu1 = std::min(iu + 1, sw_i - 1), iv1 = std::min(iv + 1, sh_i - 1);
				const double fu = u - double(iu), fv = v - double(iv);
				texels_[size_t(y) * width_ + x] = src_texture.pixel(iu, iv) * ((1.0 - fu) * (1.0 - fv)) + src_texture.pixel(iu1, iv) * (fu * (1.0 - fv))
					+ src_texture.pixel(iu, iv1) * ((1.0 - fu) * fv) + src_texture.pixel(iu1, iv1) * (fu * fv);
			}
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
