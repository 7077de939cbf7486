// bake.cpp — generator, NVRTC compilation and launch of the scene-specialised render kernel (bake.h).
#include "bake.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

// The device headers, embedded by `ld -r -b binary` (Makefile): NVRTC gets them as in-memory include files.
#define EMBEDDED(sym) extern "C" const char _binary_##sym##_start[], _binary_##sym##_end[];
EMBEDDED(rtc_compat_h) EMBEDDED(dev_types_h) EMBEDDED(vec_cuh) EMBEDDED(philox_cuh) EMBEDDED(intersect_cuh) EMBEDDED(shade_cuh)
EMBEDDED(wide_cuh) EMBEDDED(render_args_h) EMBEDDED(render_path_cuh)
#undef EMBEDDED

namespace areb {

namespace {

// ---- source generation ------------------------------------------------------------------------------------
std::string flit(float v) {  // exact literal of a float
	char b[64];
	if (v == 0.0f) return std::signbit(v) ? "-0.0f" : "0.0f";
	snprintf(b, sizeof b, "%af", (double)v);
	return b;
}
// c[0]*v[0] + c[1]*v[1] + c[2]*v[2] + tail as the fmaf nest the precompiled kernels evaluate (innermost term z), with the
// terms whose coefficient is exactly zero left out: fma(+-0, x, acc) == acc for every finite x, so the value is the same.
std::string dot3(const float c[3], const char *const v[3], const std::string &tail, bool have_tail) {
	std::string e = tail;
	bool have = have_tail;
	for (int k = 2; k >= 0; --k) {
		if (c[k] == 0.0f) continue;
		if (!have) e = "(" + flit(c[k]) + " * " + v[k] + ")";
		else e = "fmaf(" + flit(c[k]) + ", " + v[k] + ", " + e + ")";
		have = true;
	}
	return have ? e : std::string("0.0f");
}
int nonzeros(const float c[3]) { return (c[0] != 0.0f) + (c[1] != 0.0f) + (c[2] != 0.0f); }

const char *const D[3] = { "d.x", "d.y", "d.z" }, *const O[3] = { "o.x", "o.y", "o.z" }, *const P[3] = { "px", "py", "pz" };
const char *const DP[3] = { "Dx", "Dy", "Dz" };

void emit_box(std::string &s, const HotPrim &r, const HotPrim &r2, int open_class, int idx, bool packed, bool &need_pairs) {
	const f4 ax[3] = { r.r0, r.r1, r.r2 };
	char buf[256];
	s += "\t{\n";
	for (int i = 0; i < 3; ++i) {
		const float n[3] = { ax[i].x, ax[i].y, ax[i].z }, nn[3] = { -ax[i].x, -ax[i].y, -ax[i].z };
		if (packed && nonzeros(n) >= 2) {
			// (s_i, e_i) = n·(d, -o) + (1e-30, c): one FFMA2 per non-zero component
			need_pairs = true;
			std::string e = "make_float2(1e-30f, " + flit(ax[i].w) + ")";
			for (int k = 2; k >= 0; --k)
				if (n[k] != 0.0f) e = "__ffma2_rn(make_float2(" + flit(n[k]) + ", " + flit(n[k]) + "), " + DP[k] + ", " + e + ")";
			snprintf(buf, sizeof buf, "\t\tconst float2 se%d = ", i);
			s += buf + e + ";\n";
			snprintf(buf, sizeof buf, "\t\tconst float s%d = se%d.x, e%d = se%d.y;\n", i, i, i, i);
			s += buf;
		} else {
			snprintf(buf, sizeof buf, "\t\tconst float s%d = ", i);
			s += buf + dot3(n, D, "1e-30f", true) + ";\n";
			snprintf(buf, sizeof buf, "\t\tconst float e%d = ", i);
			s += buf + dot3(nn, O, flit(ax[i].w), true) + ";\n";
		}
	}
	snprintf(buf, sizeof buf, "\t\ttest_box_se<%d>(s0, s1, s2, e0, e1, e2, %s, %s, %s, %s, tmin, %d, h);\n\t}\n", open_class, flit(r2.r0.x).c_str(),
		flit(r2.r0.y).c_str(), flit(r2.r0.z).c_str(), open_class == 2 ? "true" : "false", idx);
	s += buf;
}

void emit_plane(std::string &s, const HotPrim &r, bool quad, int idx) {
	const float n[3] = { r.r0.x, r.r0.y, r.r0.z }, nn[3] = { -r.r0.x, -r.r0.y, -r.r0.z };
	const float A[3] = { r.r1.x, r.r1.y, r.r1.z }, B[3] = { r.r2.x, r.r2.y, r.r2.z };
	char buf[256];
	s += "\t{\n\t\tconst float denom = " + dot3(n, D, "", false) + ";\n";
	s += "\t\tconst float num = " + dot3(nn, O, flit(r.r0.w), true) + ";\n";
	s += "\t\tconst float t = num * rcp_fast(denom);\n";
	s += "\t\tconst float px = fmaf(t, d.x, o.x), py = fmaf(t, d.y, o.y), pz = fmaf(t, d.z, o.z);\n";
	s += "\t\tconst float a = " + dot3(A, P, flit(-r.r1.w), true) + ";\n";
	s += "\t\tconst float b = " + dot3(B, P, flit(-r.r2.w), true) + ";\n";
	snprintf(buf, sizeof buf, "\t\tplane_accept<%s>(t, a, b, tmin, %d, h);\n\t}\n", quad ? "true" : "false", idx);
	s += buf;
}

uint64_t fnv1a(const std::string &s) {
	uint64_t h = 0xcbf29ce484222325ull;
	for (unsigned char c : s) h = (h ^ c) * 0x100000001b3ull;
	return h;
}

void emit_sphere(std::string &s, const HotPrim &r, int idx) {
	char buf[512];
	snprintf(buf, sizeof buf, "\ttest_sphere(make_float4(%s, %s, %s, %s), make_float4(%s, 0.0f, 0.0f, 0.0f), o, d, tmin, %d, h);\n", flit(r.r0.x).c_str(),
		flit(r.r0.y).c_str(), flit(r.r0.z).c_str(), flit(r.r0.w).c_str(), flit(r.r1.x).c_str(), idx);
	s += buf;
}

}  // namespace

std::string bake_source(const CompiledScene &cs, bool lean, bool packed, int min_blocks) {
	if (lean ? !cs.lean_ok : (cs.brute.empty() || cs.n_hot > BAKE_MAX_SLOTS)) return std::string();
	const HotRange &br = cs.brute_range;
	std::string body;
	bool need_pairs = false;
	char buf[256];
	int slot = br.first;
	for (int i = 0; i < br.nb; ++i, slot += 2) {
		// the lean list has its open boxes first; the generic list keeps the record's own flag
		const bool open = cs.brute[(size_t)slot + 1].r0.w != 0.0f;
		snprintf(buf, sizeof buf, "\t// box %d (%s), hot slots %d-%d\n", i, open ? "one face absent" : "closed", slot, slot + 1);
		body += buf;
		emit_box(body, cs.brute[(size_t)slot], cs.brute[(size_t)slot + 1], open ? 2 : 1, slot, packed, need_pairs);
	}
	for (int j = 0; j < br.nq; ++j, ++slot) {
		snprintf(buf, sizeof buf, "\t// parallelogram, hot slot %d\n", slot);
		body += buf;
		emit_plane(body, cs.brute[(size_t)slot], true, slot);
	}
	for (int j = 0; j < br.nt; ++j, ++slot) {
		snprintf(buf, sizeof buf, "\t// triangle, hot slot %d\n", slot);
		body += buf;
		emit_plane(body, cs.brute[(size_t)slot], false, slot);
	}
	for (int j = 0; j < br.ns; ++j, ++slot) {
		snprintf(buf, sizeof buf, "\t// sphere, hot slot %d\n", slot);
		body += buf;
		emit_sphere(body, cs.brute[(size_t)slot], slot);
	}
	bool noise = false;
	for (const TextureRec &t : cs.texs) noise = noise || t.kind == TK_NOISE;
	std::string s;
	s += "// generated by bake.cpp from the committed scene: the closest-hit tests of its brute-force list as straight-line code\n";
	s += "#define ARE_BAKED 1\n";
	// __launch_bounds__ CTAs per SM.  The lean kernel needs 55 registers whatever the bound (9 CTAs fit), but ptxas schedules it
	// differently: measured on the Cornell box 5 / 6 / 7 / 8 / 9 / 10 -> 11 274 / 11 275 / 11 112 / 11 045 / 10 773 / 10 742 Msamples/s
	if (lean) s += "#define RENDER_MIN_BLOCKS_LEAN " + std::to_string(min_blocks > 0 ? min_blocks : 6) + "\n";
	else s += "#define RENDER_MIN_BLOCKS " + std::to_string(min_blocks > 0 ? min_blocks : 5) + "\n";  // textured scene, 4 / 5 / 6 / 7: 14 639 / 14 597 / 13 893 / 13 964 Msamples/s
	s += "#include \"intersect.cuh\"\nnamespace areb {\n";
	s += "__device__ __forceinline__ void intersect_baked(V3<float> o, V3<float> d, float tmin, Hit &h) {\n";
	if (need_pairs) s += "\tconst float2 Dx = make_float2(d.x, -o.x), Dy = make_float2(d.y, -o.y), Dz = make_float2(d.z, -o.z);\n";
	s += body;
	s += "}\n}  // namespace areb\n#include \"render_path.cuh\"\n";
	if (lean) {
		s += "extern \"C\" __global__ void __launch_bounds__(RENDER_THREADS, RENDER_MIN_BLOCKS_LEAN) k_render_baked(const __grid_constant__ areb::RenderArgs A) {\n";
		s += "\tareb::render_path_body<0, false, false, true, true>(A);\n}\n";
	} else {
		s += "extern \"C\" __global__ void __launch_bounds__(RENDER_THREADS, RENDER_MIN_BLOCKS) k_render_baked(const __grid_constant__ areb::RenderArgs A) {\n";
		s += std::string("\tareb::render_path_body<0, false, false, false, true, ") + (noise ? "true" : "false") + ">(A);\n}\n";
	}
	return s;
}

// ---- NVRTC + driver API, loaded on first use ----------------------------------------------------------------
namespace {

struct Rtc {
	void *lib = nullptr;
	nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
	nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
	nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
	nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
	nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
	nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
	nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
	const char *(*GetErrorString)(nvrtcResult) = nullptr;
	std::string why;
	bool ok = false;
};
struct Drv {
	void *lib = nullptr;
	CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
	CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
	CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
	CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
	std::string why;
	bool ok = false;
};

template <typename F>
bool sym(void *lib, const char *name, F &fn, std::string &why) {
	fn = reinterpret_cast<F>(dlsym(lib, name));
	if (!fn) why = std::string("missing symbol ") + name;
	return fn != nullptr;
}

Rtc &rtc() {
	static Rtc r;
	static std::once_flag once;
	std::call_once(once, [] {
		// the toolkit's own NVRTC first: a bare soname would resolve to whatever libnvrtc.so.12 the process has already
		// loaded (PyTorch bundles one of a different minor version), and the CUBIN should not depend on import order
		const char *names[] = { "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so" };
		for (const char *n : names)
			if ((r.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
		if (!r.lib) { r.why = "libnvrtc.so.12 not found"; return; }
		r.ok = sym(r.lib, "nvrtcCreateProgram", r.CreateProgram, r.why) && sym(r.lib, "nvrtcCompileProgram", r.CompileProgram, r.why) &&
			sym(r.lib, "nvrtcGetCUBINSize", r.GetCUBINSize, r.why) && sym(r.lib, "nvrtcGetCUBIN", r.GetCUBIN, r.why) &&
			sym(r.lib, "nvrtcGetProgramLogSize", r.GetProgramLogSize, r.why) && sym(r.lib, "nvrtcGetProgramLog", r.GetProgramLog, r.why) &&
			sym(r.lib, "nvrtcDestroyProgram", r.DestroyProgram, r.why) && sym(r.lib, "nvrtcGetErrorString", r.GetErrorString, r.why);
	});
	return r;
}
Drv &drv() {
	static Drv d;
	static std::once_flag once;
	std::call_once(once, [] {
		d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
		if (!d.lib) { d.why = "libcuda.so.1 not found"; return; }
		d.ok = sym(d.lib, "cuModuleLoadData", d.ModuleLoadData, d.why) && sym(d.lib, "cuModuleGetFunction", d.ModuleGetFunction, d.why) &&
			sym(d.lib, "cuLaunchKernel", d.LaunchKernel, d.why) && sym(d.lib, "cuGetErrorString", d.GetErrorString, d.why);
	});
	return d;
}

struct Embedded {
	const char *name, *begin, *end;
};
#define EMB(file, sym) { file, _binary_##sym##_start, _binary_##sym##_end }
const Embedded kHeaders[] = { EMB("rtc_compat.h", rtc_compat_h), EMB("dev_types.h", dev_types_h), EMB("vec.cuh", vec_cuh), EMB("philox.cuh", philox_cuh),
	EMB("intersect.cuh", intersect_cuh), EMB("shade.cuh", shade_cuh), EMB("wide.cuh", wide_cuh), EMB("render_args.h", render_args_h),
	EMB("render_path.cuh", render_path_cuh) };
#undef EMB

}  // namespace

bool bake_compile_cubin(const std::string &src, std::string &cubin, std::string &log) {
	Rtc &r = rtc();
	if (!r.ok) { log = "NVRTC unavailable: " + r.why; return false; }
	constexpr int NH = (int)(sizeof kHeaders / sizeof kHeaders[0]);
	std::vector<std::string> texts(NH);
	const char *hsrc[NH], *hname[NH];
	for (int i = 0; i < NH; ++i) {
		texts[i].assign(kHeaders[i].begin, kHeaders[i].end);
		hsrc[i] = texts[i].c_str();
		hname[i] = kHeaders[i].name;
	}
	nvrtcProgram prog;
	nvrtcResult st = r.CreateProgram(&prog, src.c_str(), "are_baked_scene.cu", NH, hsrc, hname);
	if (st != NVRTC_SUCCESS) { log = std::string("nvrtcCreateProgram: ") + r.GetErrorString(st); return false; }
	// the flags render.o is built with (Makefile): approximate division / square root, flushed denormals, FMA contraction
	const char *opts[] = { "--gpu-architecture=sm_100a", "--std=c++17", "--prec-div=false", "--prec-sqrt=false", "--ftz=true", "-lineinfo",
		"-default-device" };
	st = r.CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
	size_t n = 0;
	if (r.GetProgramLogSize(prog, &n) == NVRTC_SUCCESS && n > 1) {
		log.resize(n);
		r.GetProgramLog(prog, &log[0]);
	}
	if (st != NVRTC_SUCCESS) {
		log = std::string("nvrtcCompileProgram: ") + r.GetErrorString(st) + "\n" + log;
		r.DestroyProgram(&prog);
		return false;
	}
	st = r.GetCUBINSize(prog, &n);
	if (st == NVRTC_SUCCESS) {
		cubin.resize(n);
		st = r.GetCUBIN(prog, &cubin[0]);
	}
	r.DestroyProgram(&prog);
	if (st != NVRTC_SUCCESS) { log = std::string("nvrtcGetCUBIN: ") + r.GetErrorString(st); return false; }
	return true;
}

struct BakedKernel {
	CUmodule mod = nullptr;
	CUfunction fn = nullptr;
	const std::string *cubin = nullptr;  // the image it was loaded from (owned by g_cubins)
};

namespace {
std::mutex g_mu;
std::map<std::pair<uint64_t, int>, BakedKernel> g_cache;   // (source hash, device) -> loaded kernel; lives as long as the process
std::map<uint64_t, std::string> g_cubins;                  // source hash -> CUBIN (shared by all devices)
std::map<uint64_t, std::string> g_failed;                  // source hash -> why it does not compile (not retried)

std::string cu_err(Drv &d, CUresult e) {
	const char *m = nullptr;
	if (d.GetErrorString(e, &m) == CUDA_SUCCESS && m) return m;
	return "CUDA driver error " + std::to_string((int)e);
}
}  // namespace

const BakedKernel *bake_get(const CompiledScene &cs, bool lean, bool packed, int min_blocks, int device, std::string &err, double *compile_ms) {
	if (compile_ms) *compile_ms = 0.0;
	const std::string src = bake_source(cs, lean, packed, min_blocks);
	if (src.empty()) { err = lean ? "scene has no lean form" : "scene has no brute-force list of at most 16 slots"; return nullptr; }
	Drv &d = drv();
	if (!d.ok) { err = "CUDA driver API unavailable: " + d.why; return nullptr; }
	const uint64_t key = fnv1a(src);
	std::lock_guard<std::mutex> lock(g_mu);
	auto hit = g_cache.find({ key, device });
	if (hit != g_cache.end()) return &hit->second;
	auto bad = g_failed.find(key);
	if (bad != g_failed.end()) { err = bad->second; return nullptr; }
	const auto t0 = std::chrono::steady_clock::now();
	auto cb = g_cubins.find(key);
	if (cb == g_cubins.end()) {
		std::string cubin, log;
		if (!bake_compile_cubin(src, cubin, log)) { g_failed[key] = log; err = log; return nullptr; }
		cb = g_cubins.emplace(key, std::move(cubin)).first;
	}
	BakedKernel k;
	k.cubin = &cb->second;
	CUresult e = d.ModuleLoadData(&k.mod, cb->second.data());
	if (e == CUDA_SUCCESS) e = d.ModuleGetFunction(&k.fn, k.mod, "k_render_baked");
	if (e != CUDA_SUCCESS) { err = "loading the baked kernel: " + cu_err(d, e); return nullptr; }
	if (compile_ms) *compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return &g_cache.emplace(std::make_pair(key, device), k).first->second;
}

const std::string *bake_cubin(const BakedKernel *k) { return k ? k->cubin : nullptr; }

int bake_launch(const BakedKernel *k, const RenderArgs &a, int blocks, int threads, size_t smem, cudaStream_t s, std::string &err) {
	Drv &d = drv();
	if (!k || !d.ok) { err = "no baked kernel"; return -1; }
	void *params[] = { const_cast<RenderArgs *>(&a) };
	CUresult e = d.LaunchKernel(k->fn, (unsigned)blocks, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, reinterpret_cast<CUstream>(s), params, nullptr);
	if (e != CUDA_SUCCESS) { err = "launching the baked kernel: " + cu_err(d, e); return -1; }
	return 1;
}

}  // namespace areb
