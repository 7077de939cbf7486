// api.cu — implementation of the C ABI declared in include/are_cuda.h.
//
// Error model: nothing throws across the boundary; every entry point returns an are_status and leaves a message
// retrievable with are_cuda_last_error (the C++ shim turns ARE_ERR_INVALID_ARGUMENT / ARE_ERR_RUNTIME back into
// std::invalid_argument / std::runtime_error, matching /root/reference/src/object/triangle.cpp:13-36 and
// src/texture.cpp:13-79).  There is no CPU implementation behind any of these calls.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise
#define ARE_NVTX 1
#endif

#include "../../include/are_cuda.h"
#include "bake.h"
#include "dev_types.h"
#include "kernels.h"
#include "patch.h"
#include "scene.h"

using namespace areb;

struct are_cuda_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	HostScene scene;
	CompiledScene cs;
	DevScene dev;
	bool committed = false;
	std::vector<void *> scene_allocs;
	unsigned long long *d_counters = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	float *own_accum = nullptr;
	size_t own_accum_elems = 0;
	int sm_count = 0;
	CompileOptions opt;
	PatchWorkspace patch_ws;
	float *d_gamma_thr = nullptr;  // thresholds of the rt.cpp gamma-2.2 encode (built on first use)
	uint8_t *d_rgb8 = nullptr;     // tone-mapped frame, kept between calls (a cudaMalloc / cudaFree pair per call costs more than the encode)
	size_t d_rgb8_bytes = 0;
	int bvh_builder = ARE_BVH_BUILDER_HOST_SAH;
	cudaMemPool_t pool = nullptr;  // scene arrays come from a private stream-ordered pool that keeps freed blocks: a re-commit reuses them
	void *lbvh_ws = nullptr;  // scratch of the device BVH builder, grown on demand and kept between commits (a refit reads it)
	size_t lbvh_ws_bytes = 0;
	int lbvh_items = 0;       // items of the device-built hierarchy the workspace describes (0: none — no refit possible)
	std::vector<int> dirty;   // user ids of primitives moved by are_cuda_update_* since the last commit / refit
	std::vector<char> dirty_flag;
	are_commit_info commit_info = {};
	int wide_min_nodes = 0x7fffffff;  // AUTO never picks the compressed 8-wide BVH (measured slower); ARE_OPT_WIDE_MIN_NODES overrides
	// options (are_cuda_set_option)
	bool opt_lean = true;            // lean brute-force kernel for scenes that have a lean form
	bool opt_bake = true;            // scene-specialised kernel (NVRTC at commit) for the same scenes
	bool opt_bake_packed = false;    // its slab products as fma.rn.f32x2 pairs
	int opt_bake_min_blocks = 0;     // its __launch_bounds__ CTAs per SM (0: as the lean kernel)
	int opt_builder_override = -1;   // -1: are_cuda_set_bvh_builder decides
	int opt_lbvh_max_height = ARE_BVH_STACK;
	int opt_l2_persist = 0;          // BVH renders: mark the node array as L2-persisting (cudaAccessPolicyWindow) for the launch
	bool opt_quant_nodes = true;     // big BVH2 hierarchies get (commit) and use (render) a quantised 32-byte-node copy
	QGrid qgrid_host = {};           // its grid, for the launch-time guard on the camera's distance
	size_t l2_persist_max = 0, l2_window_max = 0;
	const BakedKernel *baked = nullptr;  // owned by the process-wide cache in bake.cpp
	bool baked_lean = false;             // it is the lean kernel's baked form (else the generic brute-force kernel's)
	std::string bake_note;
	// multi-device context (are_cuda_create_multi): this context drives devices[0]; `peers` are full single-device contexts
	// on the other devices, owned here.  They share this context's compiled scene (csp) and render sample shards.
	const CompiledScene *csp = nullptr;   // the compiled scene the committed device arrays were made from (&cs, or the parent's)
	std::vector<are_cuda_ctx *> peers;
	are_cuda_ctx *parent = nullptr;
	cudaEvent_t ev_done = nullptr;        // multi-device: this device's render of the current round is on its stream up to here
	cudaEvent_t ev_red = nullptr;         // ... and its reduce kernel
	void *wf_ws = nullptr;                // ray queues of the wavefront integrator, grown on demand
	size_t wf_ws_bytes = 0;
	float *stage = nullptr;               // first device, groups without full peer mapping: one frame of staging
	size_t stage_elems = 0;
	cudaStream_t own_stream = nullptr;    // peers run on their own non-blocking stream
	bool p2p_all = false;                 // every device of the group can map every other device's memory
};

static std::string g_create_error;

namespace {

int fail(are_cuda_ctx *c, int status, const std::string &msg) {
	if (c) c->err = msg;
	else g_create_error = msg;
	return status;
}
#define CK(call)                                                                                              \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                              \
		if (e_ != cudaSuccess) return fail(ctx, ARE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

struct Range {  // NVTX range around one C-ABI call (tracing hooks)
#ifdef ARE_NVTX
	explicit Range(const char *name) { nvtxRangePushA(name); }
	~Range() { nvtxRangePop(); }
#else
	explicit Range(const char *) {}
#endif
};

struct Bind {  // make the context's device current for the duration of a call
	int prev = -1;
	explicit Bind(are_cuda_ctx *c) {
		cudaGetDevice(&prev);
		if (prev != c->device) cudaSetDevice(c->device);
		else prev = -1;
	}
	~Bind() {
		if (prev >= 0) cudaSetDevice(prev);
	}
};

cudaError_t scene_malloc(are_cuda_ctx *ctx, void **p, size_t bytes) { return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, ctx->pool, ctx->stream); }

void free_scene_allocs(are_cuda_ctx *ctx) {
	for (void *p : ctx->scene_allocs) cudaFreeAsync(p, ctx->stream);
	ctx->scene_allocs.clear();
	std::memset(&ctx->dev, 0, sizeof ctx->dev);
	ctx->committed = false;
}

template <typename T>
int upload(are_cuda_ctx *ctx, const std::vector<T> &v, const T **out, uint64_t &bytes) {
	*out = nullptr;
	if (v.empty()) return ARE_OK;
	void *d = nullptr;
	CK(scene_malloc(ctx, &d, v.size() * sizeof(T)));
	ctx->scene_allocs.push_back(d);
	CK(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	bytes += v.size() * sizeof(T);
	*out = static_cast<const T *>(d);
	return ARE_OK;
}

inline size_t nz(size_t b) { return b ? b : 1; }
// scoped temporary device buffer
struct Tmp {
	void *p = nullptr;
	~Tmp() {
		if (p) cudaFree(p);
	}
	template <typename T>
	T *as() { return static_cast<T *>(p); }
};
#define TMP_IN(tmp, host, bytes)                                                              \
	do {                                                                                      \
		CK(cudaMalloc(&(tmp).p, nz(bytes)));                                      \
		CK(cudaMemcpyAsync((tmp).p, (host), (bytes), cudaMemcpyHostToDevice, ctx->stream));   \
	} while (0)
#define TMP_OUT(tmp, bytes) CK(cudaMalloc(&(tmp).p, nz(bytes)))
#define GET_OUT(host, tmp, bytes)                                                                          \
	do {                                                                                                   \
		if (host) CK(cudaMemcpyAsync((host), (tmp).p, (bytes), cudaMemcpyDeviceToHost, ctx->stream));      \
	} while (0)

// -> kernel mode: 0 brute force (shared memory), 1 BVH2, 2 compressed 8-wide BVH
bool resolve_traversal(are_cuda_ctx *ctx, int traversal, int &mode) {
	const bool brute_ok = ctx->dev.brute != nullptr, wide_ok = ctx->dev.wnodes != nullptr;
	if (traversal == ARE_TRAVERSAL_BRUTE) {
		if (!brute_ok) return false;
		mode = 0;
	} else if (traversal == ARE_TRAVERSAL_BVH) mode = 1;
	else if (traversal == ARE_TRAVERSAL_WIDE) mode = wide_ok ? 2 : 1;  // a single-primitive scene has no wide hierarchy
	else if (traversal == ARE_TRAVERSAL_BVH4) mode = ctx->dev.nodes4 ? 3 : 1;  // not built (option off, device-built tree, too deep): BVH2
	else if (traversal == ARE_TRAVERSAL_AUTO) {
		// Brute force (baked, from shared memory) against the BVH2, measured: a room where every ray hits something — textured
		// scene, 6 slots: 14.5 against 8.9 Gsamples/s; Cornell, 4 fused items: 11.1 against 4.9 — and a sparse cloud where most rays
		// miss everything, 4 / 8 / 16 / 32 slots: 85 / 65 / 41 / 24 against 71 / 66 / 54 / 50.  16 = BAKE_MAX_SLOTS.
		if (brute_ok && ctx->csp->n_hot <= 16) mode = 0;
		else mode = (wide_ok && (int)ctx->csp->nodes.size() > ctx->wide_min_nodes) ? 2 : 1;
	} else return false;
	return true;
}

int need_commit(are_cuda_ctx *ctx) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (!ctx->committed) return fail(ctx, ARE_ERR_NOT_COMMITTED, "scene not committed: call are_cuda_commit first");
	return ARE_OK;
}

int add_prim(are_cuda_ctx *ctx, int type, const double Q[3], const double u[3], const double v[3], int mat, int tex) {
	if (!ctx || !Q || !u || !v) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	// are::Triangle's ctor rejects null Material* / Texture* (triangle.cpp:13-20); ids play that role here
	if (mat < 0 || mat >= (int)ctx->scene.materials.size()) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "Material pointer cannot be null");
	if (tex < 0 || tex >= (int)ctx->scene.textures.size()) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "Texture pointer cannot be null");
	if (const char *m = validate_edges(u, v)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, m);
	HostPrim p;
	p.type = type; p.mat = mat; p.tex = tex;
	std::memcpy(p.Q, Q, sizeof p.Q);
	std::memcpy(p.u, u, sizeof p.u);
	std::memcpy(p.v, v, sizeof p.v);
	ctx->scene.prims.push_back(p);
	ctx->committed = false;
	return (int)ctx->scene.prims.size() - 1;
}

}  // namespace

extern "C" {

int are_cuda_abi_version(void) { return ARE_CUDA_ABI_VERSION; }

int are_cuda_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int are_cuda_create(are_cuda_ctx **out, int device) {
	are_cuda_ctx *ctx = nullptr;
	if (!out) return fail(nullptr, ARE_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		return fail(nullptr, ARE_ERR_NO_DEVICE, "no CUDA device available: this library has no CPU path");
	}
	if (device < 0 || device >= n) return fail(nullptr, ARE_ERR_INVALID_ARGUMENT, "device index out of range");
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) return fail(nullptr, ARE_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; kernels are built for sm_100a only");
	ctx = new are_cuda_ctx();
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
	ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
	std::memset(&ctx->dev, 0, sizeof ctx->dev);
	Bind b(ctx);
	if (ctx->l2_persist_max > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max) != cudaSuccess) { cudaGetLastError(); ctx->l2_persist_max = 0; }
	cudaError_t e1 = cudaMalloc((void **)&ctx->d_counters, CNT_N * sizeof(unsigned long long));
	cudaError_t e2 = cudaEventCreate(&ctx->ev0), e3 = cudaEventCreate(&ctx->ev1);
	if (e1 == cudaSuccess) {
		cudaMemPoolProps props = {};
		props.allocType = cudaMemAllocationTypePinned;
		props.handleTypes = cudaMemHandleTypeNone;
		props.location.type = cudaMemLocationTypeDevice;
		props.location.id = device;
		e1 = cudaMemPoolCreate(&ctx->pool, &props);
		if (e1 == cudaSuccess) {
			uint64_t keep = ~0ull;  // never hand freed blocks back to the driver while the context lives
			e1 = cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
	}
	if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
		std::string m = std::string("context set-up failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
		delete ctx;
		return fail(nullptr, ARE_ERR_CUDA, m);
	}
	*out = ctx;
	return ARE_OK;
}

int are_cuda_create_multi(are_cuda_ctx **out, const int *devices, int n_devices) {
	if (!out) return fail(nullptr, ARE_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	if (!devices || n_devices < 1 || n_devices > ARE_MAX_GROUP_DEVICES) return fail(nullptr, ARE_ERR_INVALID_ARGUMENT, "1 to 16 devices expected");
	for (int i = 0; i < n_devices; ++i)
		for (int j = 0; j < i; ++j)
			if (devices[i] == devices[j]) return fail(nullptr, ARE_ERR_INVALID_ARGUMENT, "a device is listed twice");
	are_cuda_ctx *root = nullptr;
	int st = are_cuda_create(&root, devices[0]);
	if (st != ARE_OK) return st;
	std::vector<are_cuda_ctx *> all(1, root);
	for (int i = 1; i < n_devices && st == ARE_OK; ++i) {
		are_cuda_ctx *c = nullptr;
		st = are_cuda_create(&c, devices[i]);
		if (st != ARE_OK) break;
		c->parent = root;
		root->peers.push_back(c);
		all.push_back(c);
		Bind b(c);
		if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { st = fail(nullptr, ARE_ERR_CUDA, "peer stream creation failed"); break; }
		c->stream = c->own_stream;
	}
	bool p2p = true;
	for (size_t i = 0; st == ARE_OK && i < all.size(); ++i) {
		Bind b(all[i]);
		if (cudaEventCreateWithFlags(&all[i]->ev_done, cudaEventDisableTiming) != cudaSuccess ||
			cudaEventCreateWithFlags(&all[i]->ev_red, cudaEventDisableTiming) != cudaSuccess) { st = fail(nullptr, ARE_ERR_CUDA, "event creation failed"); break; }
		for (size_t j = 0; j < all.size(); ++j) {
			if (i == j) continue;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, all[i]->device, all[j]->device);
			if (can) {
				cudaError_t e = cudaDeviceEnablePeerAccess(all[j]->device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
				cudaGetLastError();
			}
			p2p = p2p && can;
		}
	}
	if (st != ARE_OK) { are_cuda_destroy(root); return st; }
	root->p2p_all = p2p && all.size() > 1;
	*out = root;
	return ARE_OK;
}

int are_cuda_group_info(are_cuda_ctx *ctx, int *n_devices, int *peer_mapped) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (n_devices) *n_devices = 1 + (int)ctx->peers.size();
	if (peer_mapped) *peer_mapped = ctx->p2p_all ? 1 : 0;
	return ARE_OK;
}

void are_cuda_destroy(are_cuda_ctx *ctx) {
	if (!ctx) return;
	for (are_cuda_ctx *p : ctx->peers) {  // a group owns its peer contexts
		p->parent = nullptr;
		are_cuda_destroy(p);
	}
	ctx->peers.clear();
	Bind b(ctx);
	cudaStreamSynchronize(ctx->stream);
	free_scene_allocs(ctx);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
	ctx->patch_ws.release();
	if (ctx->lbvh_ws) cudaFree(ctx->lbvh_ws);
	if (ctx->d_gamma_thr) cudaFree(ctx->d_gamma_thr);
	if (ctx->d_rgb8) cudaFree(ctx->d_rgb8);
	if (ctx->own_accum) cudaFree(ctx->own_accum);
	if (ctx->d_counters) cudaFree(ctx->d_counters);
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
	if (ctx->ev_red) cudaEventDestroy(ctx->ev_red);
	if (ctx->stage) cudaFree(ctx->stage);
	if (ctx->wf_ws) cudaFree(ctx->wf_ws);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

const char *are_cuda_last_error(are_cuda_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int are_cuda_set_stream(are_cuda_ctx *ctx, void *cuda_stream) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	ctx->stream = static_cast<cudaStream_t>(cuda_stream);
	return ARE_OK;
}

int are_cuda_set_bvh_builder(are_cuda_ctx *ctx, int builder) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (builder != ARE_BVH_BUILDER_HOST_SAH && builder != ARE_BVH_BUILDER_DEVICE_LBVH) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown BVH builder");
	ctx->bvh_builder = builder;
	return ARE_OK;
}
void are_cuda_set_build_threads(int n) { set_build_threads(n); }

int are_cuda_get_baked_cubin(are_cuda_ctx *ctx, void *out, uint64_t cap, uint64_t *size) {
	int st = need_commit(ctx);
	if (st) return st;
	const std::string *img = bake_cubin(ctx->baked);
	if (size) *size = img ? img->size() : 0;
	if (!img) return fail(ctx, ARE_ERR_RUNTIME, ctx->bake_note.empty() ? "the committed scene has no baked kernel" : ctx->bake_note);
	if (out && cap >= img->size()) std::memcpy(out, img->data(), img->size());
	return ARE_OK;
}

int are_cuda_set_option(are_cuda_ctx *ctx, int option, int value) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	switch (option) {
	case ARE_OPT_LEAN_KERNEL: ctx->opt_lean = value != 0; return ARE_OK;  // takes effect at the next render
	case ARE_OPT_BAKED_KERNEL: ctx->opt_bake = value != 0; return ARE_OK;  // off: at once; on: from the next commit
	case ARE_OPT_BAKED_PACKED: ctx->opt_bake_packed = value != 0; return ARE_OK;
	case ARE_OPT_FUSE_PARALLELOGRAMS: ctx->opt.fuse_parallelograms = value != 0; return ARE_OK;
	case ARE_OPT_FUSE_BOXES: ctx->opt.fuse_boxes = value != 0; return ARE_OK;
	case ARE_OPT_BUILD_WIDE: ctx->opt.build_wide = value != 0; return ARE_OK;
	case ARE_OPT_WIDE_MIN_NODES: ctx->wide_min_nodes = value; return ARE_OK;
	case ARE_OPT_LBVH_MAX_HEIGHT: ctx->opt_lbvh_max_height = value > 0 ? value : ARE_BVH_STACK; return ARE_OK;
	case ARE_OPT_L2_PERSIST_NODES: ctx->opt_l2_persist = value; return ARE_OK;
	case ARE_OPT_BUILD_BVH4: ctx->opt.build_bvh4 = value != 0; return ARE_OK;
	case ARE_OPT_BAKED_MIN_BLOCKS: ctx->opt_bake_min_blocks = value > 0 && value <= 16 ? value : 0; return ARE_OK;
	case ARE_OPT_QUANTIZED_NODES: ctx->opt_quant_nodes = value != 0; return ARE_OK;  // off: at once; on: from the next commit
	default: return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown option");
	}
}
int are_cuda_get_commit_info(are_cuda_ctx *ctx, are_commit_info *out) {
	int st = need_commit(ctx);
	if (st) return st;
	if (!out) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	*out = ctx->commit_info;
	return ARE_OK;
}

int are_cuda_add_texture(are_cuda_ctx *ctx, int kind, const double params[8], const double *rgb, int w, int h) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (kind < ARE_TEX_SOLID || kind > ARE_TEX_IMAGE) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown texture kind");
	HostTexture t;
	t.kind = kind;
	if (params) std::memcpy(t.p, params, sizeof t.p);
	if (kind == ARE_TEX_IMAGE) {
		// are::Texture(w,h,fill) throws std::runtime_error on non-positive sizes (src/texture.cpp:53-55)
		if (w <= 0 || h <= 0) return fail(ctx, ARE_ERR_RUNTIME, "Texture width and height must be positive.");
		if (!rgb) return fail(ctx, ARE_ERR_RUNTIME, "Texture is not initialized.");
		t.w = w; t.h = h;
		t.rgb.assign(rgb, rgb + (size_t)w * h * 3);
	} else if ((kind == ARE_TEX_CHECKER_UV || kind == ARE_TEX_CHECKER_3D) && !(t.p[0] != 0.0)) {
		return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "checker scale must be non-zero");
	}
	ctx->scene.textures.push_back(std::move(t));
	ctx->committed = false;
	return (int)ctx->scene.textures.size() - 1;
}

int are_cuda_add_material(are_cuda_ctx *ctx, int kind, const double params[8]) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (kind < ARE_MAT_DIFFUSE || kind > ARE_MAT_DIFFUSE_LIGHT) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown material kind");
	HostMaterial m;
	m.kind = kind;
	if (params) std::memcpy(m.p, params, sizeof m.p);
	else {  // documented defaults: no texture override, unit radiance scale, untinted mirror, glass
		if (kind == ARE_MAT_LAMBERTIAN || kind == ARE_MAT_DIFFUSE_LIGHT) m.p[0] = -1.0;
		if (kind == ARE_MAT_METAL) m.p[1] = -1.0;
		if (kind == ARE_MAT_DIFFUSE_LIGHT) m.p[1] = 1.0;
		if (kind == ARE_MAT_REFLECTIVE) { m.p[1] = m.p[2] = m.p[3] = 1.0; }
		if (kind == ARE_MAT_DIELECTRIC) m.p[0] = 1.5;
	}
	if (kind == ARE_MAT_DIELECTRIC && !(m.p[0] > 0.0)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "index of refraction must be positive");
	{  // a texture override is -1 or the id of a texture that exists already
		const int slot = (kind == ARE_MAT_LAMBERTIAN || kind == ARE_MAT_DIFFUSE_LIGHT) ? 0 : (kind == ARE_MAT_METAL ? 1 : -1);
		if (slot >= 0) {
			const double o = m.p[slot];
			if (!(o == std::floor(o)) || o < -1.0 || o >= (double)ctx->scene.textures.size())
				return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "texture override must be -1 or the id of an existing texture");
		}
	}
	ctx->scene.materials.push_back(m);
	ctx->committed = false;
	return (int)ctx->scene.materials.size() - 1;
}

int are_cuda_add_triangle(are_cuda_ctx *ctx, const double Q[3], const double u[3], const double v[3], int material_id, int texture_id) {
	return add_prim(ctx, PT_TRIANGLE, Q, u, v, material_id, texture_id);
}

int are_cuda_set_triangle_uv(are_cuda_ctx *ctx, int prim_id, const double uv[6]) {
	if (!ctx || !uv) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (prim_id < 0 || prim_id >= (int)ctx->scene.prims.size() || ctx->scene.prims[prim_id].type != PT_TRIANGLE)
		return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "prim_id is not a triangle");
	std::memcpy(ctx->scene.prims[prim_id].uv, uv, 6 * sizeof(double));
	ctx->committed = false;
	return ARE_OK;
}

int are_cuda_add_quad(are_cuda_ctx *ctx, const double Q[3], const double u[3], const double v[3], int material_id, int texture_id) {
	return add_prim(ctx, PT_QUAD, Q, u, v, material_id, texture_id);
}

int are_cuda_add_sphere(are_cuda_ctx *ctx, const double center[3], double radius, int material_id, int texture_id) {
	if (!ctx || !center) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (!(radius > 0.0) || !std::isfinite(radius)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "sphere radius must be positive");
	if (material_id < 0 || material_id >= (int)ctx->scene.materials.size()) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "Material pointer cannot be null");
	if (texture_id < 0 || texture_id >= (int)ctx->scene.textures.size()) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "Texture pointer cannot be null");
	HostPrim p;
	p.type = PT_SPHERE; p.mat = material_id; p.tex = texture_id;
	std::memcpy(p.Q, center, sizeof p.Q);
	p.u[0] = radius; p.u[1] = p.u[2] = 0.0;
	p.v[0] = p.v[1] = p.v[2] = 0.0;
	ctx->scene.prims.push_back(p);
	ctx->committed = false;
	return (int)ctx->scene.prims.size() - 1;
}

int are_cuda_add_triangles(are_cuda_ctx *ctx, int n, const double *Q, const double *u, const double *v, const int *material_id, const int *texture_id) {
	if (!ctx || n < 0 || (n && (!Q || !u || !v || !material_id || !texture_id))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	const size_t before = ctx->scene.prims.size();
	ctx->scene.prims.reserve(before + n);
	for (int i = 0; i < n; ++i) {
		int r = add_prim(ctx, PT_TRIANGLE, Q + 3 * (size_t)i, u + 3 * (size_t)i, v + 3 * (size_t)i, material_id[i], texture_id[i]);
		if (r < 0) { ctx->scene.prims.resize(before); return r; }
	}
	return (int)before;
}

int are_cuda_add_spheres(are_cuda_ctx *ctx, int n, const double *center, const double *radius, const int *material_id, const int *texture_id) {
	if (!ctx || n < 0 || (n && (!center || !radius || !material_id || !texture_id))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	const size_t before = ctx->scene.prims.size();
	ctx->scene.prims.reserve(before + n);
	for (int i = 0; i < n; ++i) {
		int r = are_cuda_add_sphere(ctx, center + 3 * (size_t)i, radius[i], material_id[i], texture_id[i]);
		if (r < 0) { ctx->scene.prims.resize(before); return r; }
	}
	return (int)before;
}

int are_cuda_clear(are_cuda_ctx *ctx) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	Bind b(ctx);
	cudaStreamSynchronize(ctx->stream);
	ctx->scene = HostScene();
	free_scene_allocs(ctx);
	return ARE_OK;
}

int are_cuda_num_primitives(are_cuda_ctx *ctx) { return ctx ? (int)ctx->scene.prims.size() : ARE_ERR_INVALID_ARGUMENT; }

// Big BVH2 hierarchies: (re)build the quantised copy of d.nodes (dev_types.h: BvhNodeQ) on ctx's stream.
static int quantize_nodes(are_cuda_ctx *ctx, DevScene &d, are_commit_info &info) {
	const are_cuda_ctx *root = ctx->parent ? ctx->parent : ctx;
	if (!d.nodes_q) {  // commit: allocate
		if (!root->opt_quant_nodes || !d.nodes || d.n_nodes <= render_path_big_nodes()) return ARE_OK;
		void *p = nullptr;
		CK(scene_malloc(ctx, &p, (size_t)d.n_nodes * sizeof(BvhNodeQ))); ctx->scene_allocs.push_back(p); d.nodes_q = static_cast<const BvhNodeQ *>(p);
		CK(scene_malloc(ctx, &p, sizeof(QGrid))); ctx->scene_allocs.push_back(p); d.qgrid = static_cast<const QGrid *>(p);
	}
	const int launched = launch_quantize_nodes(d.nodes, d.n_nodes, const_cast<BvhNodeQ *>(d.nodes_q), const_cast<QGrid *>(d.qgrid), ctx->stream);
	if (launched < 0) return fail(ctx, ARE_ERR_CUDA, "node quantisation failed");
	CK(cudaMemcpyAsync(&ctx->qgrid_host, d.qgrid, sizeof(QGrid), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	info.device_bvh_launches += (uint64_t)launched;
	// A scene whose extent dwarfs its primitives (one huge sphere under thousands of small ones) has a grid too coarse for
	// its small boxes; the padding then costs more visits than the smaller nodes save.  Measured by surface area.
	info.quant_area_permille = ctx->qgrid_host.area32 > 0.f ? (int32_t)std::min(1e6, 1000.0 * (double)ctx->qgrid_host.area_q / (double)ctx->qgrid_host.area32) : 1000;
	if (info.quant_area_permille > 1250) { d.nodes_q = nullptr; d.qgrid = nullptr; }
	return ARE_OK;
}

// Device half of a commit: upload the compiled scene `cs` to ctx's GPU (building the hierarchy there when the device
// builder was selected).  Returns ARE_OK, a negative status, or 1 when the device-built tree is taller than the traversal
// stack (the caller recompiles with the host builder).  `cs` may belong to another context (multi-device groups).
static int commit_device(are_cuda_ctx *ctx, const CompiledScene &cs, bool want_device_bvh, int max_height, bool lean, bool bake, bool bake_packed,
	are_commit_info &info, uint64_t &bytes) {
	Bind b(ctx);
	CK(cudaStreamSynchronize(ctx->stream));
	free_scene_allocs(ctx);
	DevScene d;
	std::memset(&d, 0, sizeof d);
	bytes = 0;
	ctx->lbvh_items = 0;
	const bool device_built = want_device_bvh && !cs.lb_lo.empty();
	if (device_built) {
		// ---- device BVH build: inputs are temporaries, outputs belong to the scene ----
		int st;
		LbvhInput in;
		LbvhOutput out;
		std::memset(&in, 0, sizeof in);
		std::memset(&out, 0, sizeof out);
#define UPT(vec, field)                                                       \
	if ((st = upload(ctx, cs.vec, &in.field, bytes)) != ARE_OK) return st;
		UPT(lb_lo, item_lo) UPT(lb_hi, item_hi) UPT(lb_prims, item_prims) UPT(lb_ids, item_ids) UPT(lb_slot, item_slot)
#undef UPT
		const size_t n_temp = ctx->scene_allocs.size();
		in.n_items = (int)cs.lb_lo.size();
		in.n_slots = (int)cs.lb_prims.size();
		for (int k = 0; k < 3; ++k) { in.cmin[k] = cs.lb_cmin[k]; in.cmax[k] = cs.lb_cmax[k]; }
		void *p = nullptr;
		CK(scene_malloc(ctx, &p, (size_t)(in.n_items > 1 ? in.n_items - 1 : 0) * sizeof(BvhNode))); ctx->scene_allocs.push_back(p); out.nodes = static_cast<BvhNode *>(p);
		CK(scene_malloc(ctx, &p, (size_t)in.n_slots * sizeof(HotPrim))); ctx->scene_allocs.push_back(p); out.prims = static_cast<HotPrim *>(p);
		CK(scene_malloc(ctx, &p, (size_t)in.n_slots * sizeof(HotIds))); ctx->scene_allocs.push_back(p); out.ids = static_cast<HotIds *>(p);
		const size_t ws_need = lbvh_workspace_bytes(in.n_items);
		if (ctx->lbvh_ws_bytes < ws_need) {
			if (ctx->lbvh_ws) { CK(cudaFree(ctx->lbvh_ws)); ctx->lbvh_ws = nullptr; ctx->lbvh_ws_bytes = 0; }
			CK(cudaMalloc(&ctx->lbvh_ws, ws_need + ws_need / 4));
			ctx->lbvh_ws_bytes = ws_need + ws_need / 4;
		}
		CK(cudaEventRecord(ctx->ev0, ctx->stream));
		std::string berr;
		const int launched = lbvh_build(in, out, ctx->lbvh_ws, ctx->lbvh_ws_bytes, ctx->stream, berr);
		if (launched < 0) return fail(ctx, ARE_ERR_CUDA, "device BVH build: " + berr);
		CK(cudaEventRecord(ctx->ev1, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		// the five input temporaries go; the three outputs stay
		for (size_t i = 0; i < n_temp; ++i) cudaFreeAsync(ctx->scene_allocs[i], ctx->stream);
		ctx->scene_allocs.erase(ctx->scene_allocs.begin(), ctx->scene_allocs.begin() + n_temp);
		if (out.height > max_height) {  // deeper than the traversal stack (many coincident centres): the SAH builder bounds its depth
			free_scene_allocs(ctx);
			return 1;
		}
		d.nodes = out.nodes; d.bvh_prims = out.prims; d.bvh_ids = out.ids;
		d.n_nodes = out.n_nodes;
		d.root_leaf_meta = out.root_leaf_ref;
		info.device_bvh_ms = ms;
		info.device_bvh_launches = (uint64_t)launched;
		info.bvh_height = out.height;
		ctx->lbvh_items = in.n_items;
	}
	// the traversal kernels push at most height - 1 far children above the sentinel and do not test for overflow
	if (!device_built && cs.bvh_depth > ARE_BVH_STACK) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "BVH deeper than the traversal stack");
	const BvhNode *dev_nodes = d.nodes;
	const HotPrim *dev_bvh_prims = d.bvh_prims;
	const HotIds *dev_bvh_ids = d.bvh_ids;
	const int dev_n_nodes = d.n_nodes, dev_root_leaf = d.root_leaf_meta;
	int st;
#define UP(vec, field)                                                     \
	if ((st = upload(ctx, cs.vec, &d.field, bytes)) != ARE_OK) return st;
	UP(brute, brute) UP(brute_ids, brute_ids) UP(bvh_prims, bvh_prims) UP(bvh_ids, bvh_ids) UP(nodes, nodes)
	if (!device_built) { UP(nodes4, nodes4) }
	d.n_nodes4 = d.nodes4 ? (int)cs.nodes4.size() : 0;
	if (cs.wide_depth <= 32) { UP(wnodes, wnodes) UP(wide_prims, wide_prims) UP(wide_ids, wide_ids) UP(wide_kinds, wide_kinds) }  // 32 = ARE_WIDE_STACK
	UP(info, info) UP(prim_plane, prim_plane) UP(tri_uv, tri_uv) UP(tri64, tri64) UP(quad64, quad64) UP(sph64, sph64)
	UP(tri_uv64, tri_uv64) UP(mats, mats) UP(texs, texs) UP(tex_data, tex_data) UP(rt_tris, rt_tris) UP(box_faces, box_faces) UP(shade, shade)
	UP(lean_shade, lean_shade) UP(lean_sbase, lean_sbase)
#undef UP
	d.lean_ok = cs.lean_ok ? 1 : 0;
	d.n_lean_shade = (int)cs.lean_shade.size();
	d.lean_n_open = cs.lean_n_open;
	d.brute_range = cs.brute_range;
	d.n_nodes = (int)cs.nodes.size();
	d.root_leaf_meta = cs.root_leaf_meta;
	if (device_built) { d.nodes = dev_nodes; d.bvh_prims = dev_bvh_prims; d.bvh_ids = dev_bvh_ids; d.n_nodes = dev_n_nodes; d.root_leaf_meta = dev_root_leaf; }
	d.n_wnodes = d.wnodes ? (int)cs.wnodes.size() : 0;
	st = quantize_nodes(ctx, d, info);
	if (st != ARE_OK) return st;
	d.n_hot = cs.n_hot;
	d.n_tri = cs.n_tri; d.n_quad = cs.n_quad; d.n_sph = cs.n_sph;
	d.n_mat = (int)cs.mats.size(); d.n_tex = (int)cs.texs.size();
	d.has_noise = 0;
	for (const TextureRec &t : cs.texs) d.has_noise |= t.kind == TK_NOISE ? 1 : 0;
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->dev = d;
	ctx->csp = &cs;
	ctx->committed = true;
	// scene-specialised kernel for scenes that have a lean form: generated + NVRTC-compiled once per distinct scene
	ctx->baked = nullptr;
	ctx->bake_note.clear();
	ctx->opt_lean = lean; ctx->opt_bake = bake; ctx->opt_bake_packed = bake_packed;
	if (bake && !cs.brute.empty()) {
		// the lean form where the scene has one (and the lean kernel is allowed), else the whole brute-force list of a scene of
		// at most BAKE_MAX_SLOTS hot slots around the generic kernel (textures, spheres, any material)
		ctx->baked_lean = cs.lean_ok && lean;
		if (ctx->baked_lean || cs.n_hot <= BAKE_MAX_SLOTS) {
			ctx->baked = bake_get(cs, ctx->baked_lean, bake_packed, ctx->parent ? ctx->parent->opt_bake_min_blocks : ctx->opt_bake_min_blocks, ctx->device,
				ctx->bake_note, &info.bake_compile_ms);
			info.baked = ctx->baked ? (ctx->baked_lean ? 1 : 2) : 0;
		}
	}
	info.builder = device_built ? ARE_BVH_BUILDER_DEVICE_LBVH : ARE_BVH_BUILDER_HOST_SAH;
	info.bvh_nodes = d.n_nodes;
	if (!device_built) info.bvh_height = cs.bvh_depth;
	info.hot_slots = cs.n_hot;
	return ARE_OK;
}

int are_cuda_commit(are_cuda_ctx *ctx, uint64_t *h2d_bytes) {
	Range nvtx_range("are_cuda_commit (scene compile + upload)");
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (ctx->parent) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "commit the group's first context");
	std::string err;
	ctx->opt.brute_max = (int)brute_smem_limit_prims();
	int builder = ctx->bvh_builder;
	const int max_height = std::min((int)ARE_BVH_STACK, ctx->opt_lbvh_max_height);  // ARE_OPT_LBVH_MAX_HEIGHT: test hook for the fall-back below
	are_commit_info info = {};
	uint64_t bytes = 0;
	for (int attempt = 0; attempt < 2; ++attempt) {
		ctx->committed = false;
		for (are_cuda_ctx *p : ctx->peers) p->committed = false;
		ctx->opt.device_bvh = builder == ARE_BVH_BUILDER_DEVICE_LBVH;
		const auto t0 = std::chrono::steady_clock::now();
		if (!compile_scene(ctx->scene, ctx->opt, ctx->cs, err)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, err);  // host half, once per group
		const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		info = are_commit_info();
		int st = commit_device(ctx, ctx->cs, ctx->opt.device_bvh, max_height, ctx->opt_lean, ctx->opt_bake, ctx->opt_bake_packed, info, bytes);
		for (size_t i = 0; st == ARE_OK && i < ctx->peers.size(); ++i) {
			are_commit_info pinfo = {};
			uint64_t pbytes = 0;
			st = commit_device(ctx->peers[i], ctx->cs, ctx->opt.device_bvh, max_height, ctx->opt_lean, ctx->opt_bake, ctx->opt_bake_packed, pinfo, pbytes);
			if (st < 0) fail(ctx, st, std::string("device ") + std::to_string(ctx->peers[i]->device) + ": " + ctx->peers[i]->err);
			bytes += pbytes;
			info.device_bvh_ms = std::max(info.device_bvh_ms, pinfo.device_bvh_ms);
			info.bake_compile_ms += pinfo.bake_compile_ms;
		}
		info.host_compile_ms = host_ms;
		info.host_bvh_ms = ctx->cs.host_bvh_ms;
		if (st == 1 && attempt == 0) { builder = ARE_BVH_BUILDER_HOST_SAH; continue; }
		if (st != ARE_OK) return st < 0 ? st : fail(ctx, ARE_ERR_RUNTIME, "device-built BVH rejected twice");
		break;
	}
	ctx->commit_info = info;
	ctx->dirty.clear();
	ctx->dirty_flag.clear();
	if (h2d_bytes) *h2d_bytes = bytes;
	return ARE_OK;
}

// ---- moving primitives: update + refit ---------------------------------------------------------------------
static int mark_dirty(are_cuda_ctx *ctx, int id, int type) {
	if (id < 0 || id >= (int)ctx->scene.prims.size() || ctx->scene.prims[id].type != type)
		return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "prim_id out of range or of another primitive type");
	if (ctx->dirty_flag.size() != ctx->scene.prims.size()) ctx->dirty_flag.assign(ctx->scene.prims.size(), 0);
	if (!ctx->dirty_flag[id]) { ctx->dirty_flag[id] = 1; ctx->dirty.push_back(id); }
	return ARE_OK;
}

int are_cuda_update_triangles(are_cuda_ctx *ctx, int n, const int *prim_ids, const double *Q, const double *u, const double *v) {
	if (!ctx || n < 0 || (n && (!prim_ids || !Q || !u || !v))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	for (int k = 0; k < n; ++k) {
		if (const char *why = validate_edges(u + 3 * (size_t)k, v + 3 * (size_t)k)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, why);
		int st = mark_dirty(ctx, prim_ids[k], PT_TRIANGLE);
		if (st) return st;
		HostPrim &p = ctx->scene.prims[prim_ids[k]];
		std::memcpy(p.Q, Q + 3 * (size_t)k, 24); std::memcpy(p.u, u + 3 * (size_t)k, 24); std::memcpy(p.v, v + 3 * (size_t)k, 24);
	}
	return ARE_OK;
}

int are_cuda_update_spheres(are_cuda_ctx *ctx, int n, const int *prim_ids, const double *center, const double *radius) {
	if (!ctx || n < 0 || (n && (!prim_ids || !center || !radius))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	for (int k = 0; k < n; ++k) {
		if (!(radius[k] > 0.0) || !std::isfinite(radius[k])) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "sphere radius must be positive");
		int st = mark_dirty(ctx, prim_ids[k], PT_SPHERE);
		if (st) return st;
		HostPrim &p = ctx->scene.prims[prim_ids[k]];
		std::memcpy(p.Q, center + 3 * (size_t)k, 24);
		p.u[0] = radius[k];
	}
	return ARE_OK;
}

// Device half of a refit on one context of the group.
static int refit_device(are_cuda_ctx *ctx, const CompiledScene &cs, const std::vector<int> &item, const std::vector<int> &dp, const std::vector<HotPrim> &rec,
	const std::vector<f4> &lo, const std::vector<f4> &hi, const std::vector<double> &geo, const std::vector<f4> &rt, float &ms) {
	Bind b(ctx);
	const int m = (int)item.size();
	if (ctx->lbvh_items != (int)cs.lb_lo.size() || !ctx->lbvh_ws) return fail(ctx, ARE_ERR_RUNTIME, "no device-built hierarchy to refit");
	Tmp d_item, d_dp, d_rec, d_lo, d_hi, d_geo, d_rt;
	TMP_IN(d_item, item.data(), (size_t)m * sizeof(int));
	TMP_IN(d_dp, dp.data(), (size_t)m * sizeof(int));
	TMP_IN(d_rec, rec.data(), (size_t)m * sizeof(HotPrim));
	TMP_IN(d_lo, lo.data(), (size_t)m * sizeof(f4));
	TMP_IN(d_hi, hi.data(), (size_t)m * sizeof(f4));
	TMP_IN(d_geo, geo.data(), (size_t)m * 9 * sizeof(double));
	TMP_IN(d_rt, rt.data(), (size_t)m * 3 * sizeof(f4));
	CK(cudaEventRecord(ctx->ev0, ctx->stream));
	PrimScatterArgs ps;
	ps.m = m; ps.dp = d_dp.as<int>(); ps.rec = d_rec.as<HotPrim>(); ps.geo64 = d_geo.as<double>(); ps.rt = d_rt.as<f4>();
	ps.prim_plane = const_cast<HotPrim *>(ctx->dev.prim_plane); ps.shade = const_cast<ShadeRec *>(ctx->dev.shade);
	ps.tri64 = const_cast<double *>(ctx->dev.tri64); ps.quad64 = const_cast<double *>(ctx->dev.quad64); ps.sph64 = const_cast<double *>(ctx->dev.sph64);
	ps.rt_tris = const_cast<f4 *>(ctx->dev.rt_tris);
	ps.n_tri = ctx->dev.n_tri; ps.n_quad = ctx->dev.n_quad;
	launch_prim_scatter(ps, ctx->stream);
	CK(cudaGetLastError());
	std::string err;
	const int launched = lbvh_refit(ctx->lbvh_ws, ctx->lbvh_ws_bytes, ctx->lbvh_items, m, d_item.as<int>(), d_rec.as<HotPrim>(), d_lo.as<f4>(), d_hi.as<f4>(),
		const_cast<HotPrim *>(ctx->dev.bvh_prims), const_cast<BvhNode *>(ctx->dev.nodes), ctx->stream, err);
	if (launched < 0) return fail(ctx, ARE_ERR_CUDA, "refit: " + err);
	if (ctx->dev.nodes_q) {  // the quantised copy follows (new grid: the root box may have grown)
		are_commit_info dummy = {};
		const int qs = quantize_nodes(ctx, ctx->dev, dummy);
		if (qs != ARE_OK) return qs;
	}
	CK(cudaEventRecord(ctx->ev1, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
	return ARE_OK;
}

int are_cuda_refit(are_cuda_ctx *ctx, double *device_ms) {
	Range nvtx_range("are_cuda_refit");
	int st = need_commit(ctx);
	if (st) return st;
	if (ctx->parent) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "refit the group's first context");
	if (device_ms) *device_ms = 0.0;
	CompiledScene &cs = ctx->cs;
	const size_t np = ctx->scene.prims.size();
	// eligible: hierarchy built on the device over items that are the primitives themselves (nothing fused, no boxes)
	if (ctx->lbvh_items <= 1 || cs.lb_lo.size() != np || cs.info.size() != np)
		return fail(ctx, ARE_ERR_RUNTIME, "refit needs a device-built hierarchy (ARE_BVH_BUILDER_DEVICE_LBVH) over unfused primitives: commit instead");
	if (ctx->dirty.empty()) return ARE_OK;
	std::vector<int> dp_of_user(np), item_of_dp(np);
	for (size_t dp = 0; dp < np; ++dp) dp_of_user[(size_t)cs.info[dp].user_id] = (int)dp;
	for (size_t i = 0; i < np; ++i) {
		const HotIds id = cs.lb_ids[(size_t)cs.lb_slot[i]];
		if (id.a < 0 || id.b != -1) return fail(ctx, ARE_ERR_RUNTIME, "refit needs unfused primitives: commit instead");
		item_of_dp[(size_t)id.a] = (int)i;
	}
	const int m = (int)ctx->dirty.size();
	std::vector<int> item(m), dpv(m);
	std::vector<HotPrim> rec(m);
	std::vector<f4> lo(m), hi(m), rt(3 * (size_t)m);
	std::vector<double> geo(9 * (size_t)m, 0.0);
	for (int k = 0; k < m; ++k) {
		const HostPrim &p = ctx->scene.prims[(size_t)ctx->dirty[k]];
		PrimUpdate u;
		if (!prim_update_records(p, u)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "updated primitive is degenerate");
		const int dp = dp_of_user[(size_t)ctx->dirty[k]];
		dpv[k] = dp; item[k] = item_of_dp[(size_t)dp];
		rec[k] = u.rec; lo[k] = u.lo; hi[k] = u.hi;
		for (int i = 0; i < 3; ++i) rt[3 * (size_t)k + i] = u.rt[i];
		double *g = &geo[9 * (size_t)k];
		if (p.type == PT_SPHERE) { g[0] = p.Q[0]; g[1] = p.Q[1]; g[2] = p.Q[2]; g[3] = p.u[0]; }
		else { std::memcpy(g, p.Q, 24); std::memcpy(g + 3, p.u, 24); std::memcpy(g + 6, p.v, 24); }
		// keep the host copies in step (multi-device peers and later refits read them)
		cs.prim_plane[(size_t)dp] = u.rec;
		if (p.type != PT_SPHERE) { cs.shade[(size_t)dp].r0.x = u.rec.r0.x; cs.shade[(size_t)dp].r0.y = u.rec.r0.y; cs.shade[(size_t)dp].r0.z = u.rec.r0.z; }
		cs.lb_prims[(size_t)cs.lb_slot[(size_t)item[k]]] = u.rec;
		cs.lb_lo[(size_t)item[k]] = u.lo; cs.lb_hi[(size_t)item[k]] = u.hi;
	}
	float ms = 0.f, worst = 0.f;
	st = refit_device(ctx, cs, item, dpv, rec, lo, hi, geo, rt, ms);
	worst = ms;
	for (size_t i = 0; st == ARE_OK && i < ctx->peers.size(); ++i) {
		st = refit_device(ctx->peers[i], cs, item, dpv, rec, lo, hi, geo, rt, ms);
		if (st != ARE_OK) fail(ctx, st, ctx->peers[i]->err);
		worst = std::max(worst, ms);
	}
	if (st != ARE_OK) return st;
	for (int id : ctx->dirty) ctx->dirty_flag[(size_t)id] = 0;
	ctx->dirty.clear();
	if (device_ms) *device_ms = worst;
	return ARE_OK;
}

static uint64_t fnv1a(uint64_t h, const void *data, size_t bytes) {
	const unsigned char *p = static_cast<const unsigned char *>(data);
	for (size_t i = 0; i < bytes; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
	return h;
}

int are_cuda_compile_probe(int n_tri, const double *Q, const double *u, const double *v, int out[8]) {
	return are_cuda_compile_probe_digest(n_tri, Q, u, v, out, nullptr);
}

int are_cuda_compile_probe_digest(int n_tri, const double *Q, const double *u, const double *v, int out[8], uint64_t *digest) {
	if (n_tri < 0 || !out || (n_tri && (!Q || !u || !v))) return ARE_ERR_INVALID_ARGUMENT;
	HostScene hs;
	hs.textures.emplace_back();
	hs.materials.emplace_back();
	for (int i = 0; i < n_tri; ++i) {
		if (validate_edges(u + 3 * (size_t)i, v + 3 * (size_t)i)) return ARE_ERR_INVALID_ARGUMENT;
		HostPrim p;
		p.type = PT_TRIANGLE;
		std::memcpy(p.Q, Q + 3 * (size_t)i, sizeof p.Q);
		std::memcpy(p.u, u + 3 * (size_t)i, sizeof p.u);
		std::memcpy(p.v, v + 3 * (size_t)i, sizeof p.v);
		hs.prims.push_back(p);
	}
	CompileOptions opt;
	opt.brute_max = (int)brute_smem_limit_prims();
	CompiledScene cs;
	std::string err;
	if (!compile_scene(hs, opt, cs, err)) return ARE_ERR_INVALID_ARGUMENT;
	const int r[8] = { cs.n_hot, cs.n_fused_pairs, cs.n_boxes, (int)cs.nodes.size(), cs.bvh_depth, cs.brute_range.nq, cs.brute_range.nt, cs.brute_range.nb };
	std::memcpy(out, r, sizeof r);
	if (digest) {
		uint64_t h = 0xcbf29ce484222325ull;
		h = fnv1a(h, cs.nodes.data(), cs.nodes.size() * sizeof(BvhNode));
		h = fnv1a(h, cs.bvh_prims.data(), cs.bvh_prims.size() * sizeof(HotPrim));
		h = fnv1a(h, cs.bvh_ids.data(), cs.bvh_ids.size() * sizeof(HotIds));
		*digest = h;
	}
	return ARE_OK;
}

int are_cuda_compile_probe_forms(int n_tri, const double *Q, const double *u, const double *v, int out[8]) {
	if (n_tri < 0 || !out || (n_tri && (!Q || !u || !v))) return ARE_ERR_INVALID_ARGUMENT;
	HostScene hs;
	hs.textures.emplace_back();
	hs.materials.emplace_back();
	for (int i = 0; i < n_tri; ++i) {
		if (validate_edges(u + 3 * (size_t)i, v + 3 * (size_t)i)) return ARE_ERR_INVALID_ARGUMENT;
		HostPrim p;
		p.type = PT_TRIANGLE;
		std::memcpy(p.Q, Q + 3 * (size_t)i, sizeof p.Q);
		std::memcpy(p.u, u + 3 * (size_t)i, sizeof p.u);
		std::memcpy(p.v, v + 3 * (size_t)i, sizeof p.v);
		hs.prims.push_back(p);
	}
	CompileOptions opt;
	opt.brute_max = (int)brute_smem_limit_prims();
	opt.device_bvh = true;
	CompiledScene cs;
	std::string err;
	if (!compile_scene(hs, opt, cs, err)) return ARE_ERR_INVALID_ARGUMENT;
	// every item's fp32 box must contain the fp64 vertices of the triangles behind it
	int conservative = 0;
	for (size_t i = 0; i < cs.lb_lo.size(); ++i) {
		const HotIds id = cs.lb_ids[(size_t)cs.lb_slot[i]];
		bool ok = true;
		const int tris[2] = { id.a, id.b >= 0 ? id.b : (id.b <= -2 ? -2 - id.b : -1) };
		for (int t : tris) {
			if (t < 0 || t >= cs.n_tri) continue;  // box items (a < 0) are covered through their faces' own tests elsewhere
			const double *q = &cs.tri64[9 * (size_t)t];
			for (int k = 0; k < 3; ++k) {
				const double p[3] = { q[0] + (k == 1 ? q[3] : (k == 2 ? q[6] : 0.0)), q[1] + (k == 1 ? q[4] : (k == 2 ? q[7] : 0.0)), q[2] + (k == 1 ? q[5] : (k == 2 ? q[8] : 0.0)) };
				const float lo[3] = { cs.lb_lo[i].x, cs.lb_lo[i].y, cs.lb_lo[i].z }, hi[3] = { cs.lb_hi[i].x, cs.lb_hi[i].y, cs.lb_hi[i].z };
				for (int a = 0; a < 3; ++a) ok = ok && (double)lo[a] <= p[a] && p[a] <= (double)hi[a];
			}
		}
		conservative += ok ? 1 : 0;
	}
	const int r[8] = { cs.lean_ok ? 1 : 0, (int)cs.lean_shade.size(), cs.lean_n_open, (int)cs.lb_lo.size(), (int)cs.lb_prims.size(), conservative,
		(int)cs.nodes.size(), cs.brute_range.nb };
	std::memcpy(out, r, sizeof r);
	return ARE_OK;
}

int are_cuda_bake_probe(int n_tri, const double *Q, const double *u, const double *v, int packed, char *source_out, uint64_t source_cap,
	uint64_t *source_len, const char *cubin_path) {
	if (n_tri < 0 || (n_tri && (!Q || !u || !v))) return ARE_ERR_INVALID_ARGUMENT;
	HostScene hs;
	hs.textures.emplace_back();
	hs.materials.emplace_back();
	for (int i = 0; i < n_tri; ++i) {
		if (validate_edges(u + 3 * (size_t)i, v + 3 * (size_t)i)) return ARE_ERR_INVALID_ARGUMENT;
		HostPrim p;
		p.type = PT_TRIANGLE;
		std::memcpy(p.Q, Q + 3 * (size_t)i, sizeof p.Q);
		std::memcpy(p.u, u + 3 * (size_t)i, sizeof p.u);
		std::memcpy(p.v, v + 3 * (size_t)i, sizeof p.v);
		hs.prims.push_back(p);
	}
	CompileOptions opt;
	opt.brute_max = (int)brute_smem_limit_prims();
	CompiledScene cs;
	std::string err;
	if (!compile_scene(hs, opt, cs, err)) return ARE_ERR_INVALID_ARGUMENT;
	const std::string src = bake_source(cs, cs.lean_ok && !(packed & 2), (packed & 1) != 0);
	if (source_len) *source_len = src.size();
	if (src.empty()) return ARE_ERR_RUNTIME;  // no lean form
	if (source_out && source_cap) {
		const size_t n = std::min<size_t>(src.size(), (size_t)source_cap - 1);
		std::memcpy(source_out, src.data(), n);
		source_out[n] = 0;
	}
	if (cubin_path) {
		std::string cubin, log;
		if (!bake_compile_cubin(src, cubin, log)) { g_create_error = log; return ARE_ERR_RUNTIME; }
		FILE *f = fopen(cubin_path, "wb");
		if (!f) return ARE_ERR_IO;
		const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
		fclose(f);
		if (!ok) return ARE_ERR_IO;
	}
	return ARE_OK;
}

// ---- per-ray harness -------------------------------------------------------------------------------------
int are_cuda_hit_batch(are_cuda_ctx *ctx, int n, const double *ray_Q, const double *ray_D, double t_min, int precision, int traversal,
	int *prim, double *t, double *P, double *N, double *uv) {
	Range nvtx_range("are_cuda_hit_batch");
	int st = need_commit(ctx);
	if (st) return st;
	if (n < 0 || (n && (!ray_Q || !ray_D))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null ray arrays");
	if (precision != 32 && precision != 64) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "precision must be 32 or 64");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	const size_t n3 = (size_t)n * 3 * sizeof(double);
	Tmp dQ, dD, dprim, dt, dP, dN, duv;
	TMP_IN(dQ, ray_Q, n3);
	TMP_IN(dD, ray_D, n3);
	TMP_OUT(dprim, n * sizeof(int)); TMP_OUT(dt, n * sizeof(double)); TMP_OUT(dP, n3); TMP_OUT(dN, n3); TMP_OUT(duv, (size_t)n * 2 * sizeof(double));
	if (precision == 64) launch_hit64(ctx->dev, n, dQ.as<double>(), dD.as<double>(), t_min, dprim.as<int>(), dt.as<double>(), dP.as<double>(), dN.as<double>(), duv.as<double>(), ctx->stream);
	else {
		int mode;
		if (!resolve_traversal(ctx, traversal, mode)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown traversal, or scene too large for brute-force traversal");
		launch_hit32(ctx->dev, n, dQ.as<double>(), dD.as<double>(), t_min, mode, dprim.as<int>(), dt.as<double>(), dP.as<double>(), dN.as<double>(), duv.as<double>(), ctx->stream);
	}
	CK(cudaGetLastError());
	GET_OUT(prim, dprim, n * sizeof(int)); GET_OUT(t, dt, n * sizeof(double)); GET_OUT(P, dP, n3); GET_OUT(N, dN, n3); GET_OUT(uv, duv, (size_t)n * 2 * sizeof(double));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_scatter_batch(are_cuda_ctx *ctx, int n, const int *material_id, const int *texture_id, const double *wi, const double *N, const double *P,
	const double *uv, const double *rnd, int precision, double *wo, double *attenuation, double *emitted, int *alive) {
	int st = need_commit(ctx);
	if (st) return st;
	if (n < 0 || (n && (!material_id || !texture_id || !wi || !N || !P || !uv || !rnd))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null input arrays");
	if (precision != 32 && precision != 64) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "precision must be 32 or 64");
	for (int i = 0; i < n; ++i)
		if (material_id[i] < 0 || material_id[i] >= ctx->dev.n_mat || texture_id[i] < 0 || texture_id[i] >= ctx->dev.n_tex)
			return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "material / texture id out of range");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	const size_t n3 = (size_t)n * 3 * sizeof(double);
	Tmp dm, dx, dwi, dN, dP, duv, drnd, dwo, datt, dem, dal;
	TMP_IN(dm, material_id, n * sizeof(int)); TMP_IN(dx, texture_id, n * sizeof(int));
	TMP_IN(dwi, wi, n3); TMP_IN(dN, N, n3); TMP_IN(dP, P, n3);
	TMP_IN(duv, uv, (size_t)n * 2 * sizeof(double)); TMP_IN(drnd, rnd, (size_t)n * 4 * sizeof(double));
	TMP_OUT(dwo, n3); TMP_OUT(datt, n3); TMP_OUT(dem, n3); TMP_OUT(dal, n * sizeof(int));
	if (precision == 64) launch_scatter64(ctx->dev, n, dm.as<int>(), dx.as<int>(), dwi.as<double>(), dN.as<double>(), dP.as<double>(), duv.as<double>(), drnd.as<double>(), dwo.as<double>(), datt.as<double>(), dem.as<double>(), dal.as<int>(), ctx->stream);
	else launch_scatter32(ctx->dev, n, dm.as<int>(), dx.as<int>(), dwi.as<double>(), dN.as<double>(), dP.as<double>(), duv.as<double>(), drnd.as<double>(), dwo.as<double>(), datt.as<double>(), dem.as<double>(), dal.as<int>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(wo, dwo, n3); GET_OUT(attenuation, datt, n3); GET_OUT(emitted, dem, n3); GET_OUT(alive, dal, n * sizeof(int));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_texture_batch(are_cuda_ctx *ctx, int n, const int *texture_id, const double *uv, const double *P, int precision, double *rgb) {
	int st = need_commit(ctx);
	if (st) return st;
	if (n < 0 || (n && (!texture_id || !uv || !P || !rgb))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null arrays");
	if (precision != 32 && precision != 64) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "precision must be 32 or 64");
	for (int i = 0; i < n; ++i)
		if (texture_id[i] < 0 || texture_id[i] >= ctx->dev.n_tex) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "texture id out of range");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	Tmp dx, duv, dP, drgb;
	TMP_IN(dx, texture_id, n * sizeof(int)); TMP_IN(duv, uv, (size_t)n * 2 * sizeof(double)); TMP_IN(dP, P, (size_t)n * 3 * sizeof(double));
	TMP_OUT(drgb, (size_t)n * 3 * sizeof(double));
	if (precision == 64) launch_texture64(ctx->dev, n, dx.as<int>(), duv.as<double>(), dP.as<double>(), drgb.as<double>(), ctx->stream);
	else launch_texture32(ctx->dev, n, dx.as<int>(), duv.as<double>(), dP.as<double>(), drgb.as<double>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(rgb, drgb, (size_t)n * 3 * sizeof(double));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_camera_rays(are_cuda_ctx *ctx, const are_camera *cam, int width, int height, int n, const int *px, const int *py, const double *rnd,
	int precision, double *ray_Q, double *ray_D) {
	if (!ctx || !cam || n < 0 || (n && (!px || !py || !rnd))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (width <= 0 || height <= 0 || !(cam->focus_dist > 0.0)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad camera / image size");
	if (precision != 32 && precision != 64) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "precision must be 32 or 64");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	CamBasis cb;
	make_cam_basis(cam->pos, cam->target, cam->up, cam->vfov_deg, cam->focus_dist, cam->defocus_angle_deg, cam->jitter, width, height, cb);
	Tmp dpx, dpy, drnd, dQ, dD;
	TMP_IN(dpx, px, n * sizeof(int)); TMP_IN(dpy, py, n * sizeof(int)); TMP_IN(drnd, rnd, (size_t)n * 4 * sizeof(double));
	TMP_OUT(dQ, (size_t)n * 3 * sizeof(double)); TMP_OUT(dD, (size_t)n * 3 * sizeof(double));
	if (precision == 64) launch_camera64(cb, width, height, n, dpx.as<int>(), dpy.as<int>(), drnd.as<double>(), dQ.as<double>(), dD.as<double>(), ctx->stream);
	else launch_camera32(cb, width, height, n, dpx.as<int>(), dpy.as<int>(), drnd.as<double>(), dQ.as<double>(), dD.as<double>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(ray_Q, dQ, (size_t)n * 3 * sizeof(double)); GET_OUT(ray_D, dD, (size_t)n * 3 * sizeof(double));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_plane_batch(are_cuda_ctx *ctx, int n, const double *plane4, const double *ray_Q, const double *ray_D, int *hit, double *P) {
	if (!ctx || n < 0 || (n && (!plane4 || !ray_Q || !ray_D))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	Tmp dpl, dQ, dD, dhit, dP;
	TMP_IN(dpl, plane4, (size_t)n * 4 * sizeof(double));
	TMP_IN(dQ, ray_Q, (size_t)n * 3 * sizeof(double));
	TMP_IN(dD, ray_D, (size_t)n * 3 * sizeof(double));
	TMP_OUT(dhit, (size_t)n * sizeof(int));
	TMP_OUT(dP, (size_t)n * 3 * sizeof(double));
	launch_plane64(n, dpl.as<double>(), dQ.as<double>(), dD.as<double>(), dhit.as<int>(), dP.as<double>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(hit, dhit, (size_t)n * sizeof(int));
	GET_OUT(P, dP, (size_t)n * 3 * sizeof(double));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_point_in_batch(are_cuda_ctx *ctx, const double Q[3], const double u[3], const double v[3], int n, const double *points, int *inside) {
	if (!ctx || n < 0 || !Q || !u || !v || (n && (!points || !inside))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (const char *why = validate_edges(u, v)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, why);  // are::Triangle's ctor checks
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	double tri9[9];
	std::memcpy(tri9, Q, 24); std::memcpy(tri9 + 3, u, 24); std::memcpy(tri9 + 6, v, 24);
	Tmp dtri, dpts, din;
	TMP_IN(dtri, tri9, sizeof tri9);
	TMP_IN(dpts, points, (size_t)n * 3 * sizeof(double));
	TMP_OUT(din, (size_t)n * sizeof(int));
	launch_point_in64(n, dtri.as<double>(), dpts.as<double>(), din.as<int>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(inside, din, (size_t)n * sizeof(int));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_material_reflect_batch(are_cuda_ctx *ctx, int kind, int n, const double *plane4, const double *origin, int *ok, double *new_origin) {
	if (!ctx || n < 0 || (n && (!plane4 || !origin))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (kind != ARE_MAT_DIFFUSE && kind != ARE_MAT_REFLECTIVE) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "the reference has two materials: Diffuse and Reflective");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	Tmp dpl, dorg, dok, dout;
	TMP_IN(dpl, plane4, (size_t)n * 4 * sizeof(double));
	TMP_IN(dorg, origin, (size_t)n * 3 * sizeof(double));
	TMP_OUT(dok, (size_t)n * sizeof(int));
	TMP_OUT(dout, (size_t)n * 3 * sizeof(double));
	launch_material_reflect64(kind, n, dpl.as<double>(), dorg.as<double>(), dok.as<int>(), dout.as<double>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(ok, dok, (size_t)n * sizeof(int));
	GET_OUT(new_origin, dout, (size_t)n * 3 * sizeof(double));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_philox_batch(are_cuda_ctx *ctx, int n, uint64_t seed, const uint32_t *counter, uint32_t *out) {
	if (!ctx || n < 0 || (n && (!counter || !out))) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (n == 0) return ARE_OK;
	Bind b(ctx);
	Tmp dc, dout;
	TMP_IN(dc, counter, (size_t)n * 16);
	TMP_OUT(dout, (size_t)n * 16);
	launch_philox(n, seed, dc.as<uint32_t>(), dout.as<uint32_t>(), ctx->stream);
	CK(cudaGetLastError());
	GET_OUT(out, dout, (size_t)n * 16);
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

// ---- rendering -------------------------------------------------------------------------------------------
}  // extern "C"

// One device: add the sample sums of [sample_begin, sample_begin + sample_count) into accum (on ctx's GPU).
static int render_single(are_cuda_ctx *ctx, const are_camera *cam, const are_render_params *p, float *accum, are_render_stats *stats, int count_tests) {
	int st = need_commit(ctx);
	if (st) return st;
	if (!cam || !p || !accum) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (p->width <= 0 || p->height <= 0 || p->sample_count < 0 || p->sample_begin < 0 || p->max_depth < 1 || !(cam->focus_dist > 0.0))
		return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad render parameters");
	if ((long long)p->width * p->height > 0x7fffffffLL) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "image too large");
	Bind b(ctx);
	RenderArgs a;
	std::memset(&a, 0, sizeof a);
	a.sc = ctx->dev;
	if (a.sc.nodes_q && !(ctx->parent ? ctx->parent : ctx)->opt_quant_nodes) { a.sc.nodes_q = nullptr; a.sc.qgrid = nullptr; }
	make_cam_basis(cam->pos, cam->target, cam->up, cam->vfov_deg, cam->focus_dist, cam->defocus_angle_deg, cam->jitter, p->width, p->height, a.cam);
	for (int k = 0; k < 3; ++k) {
		a.camf.pos[k] = (float)a.cam.pos[k]; a.camf.fwd[k] = (float)a.cam.fwd[k];
		a.camf.right[k] = (float)a.cam.right[k]; a.camf.up[k] = (float)a.cam.up[k];
	}
	a.camf.sx = (float)a.cam.sx; a.camf.sy = (float)a.cam.sy; a.camf.lens_r = (float)a.cam.lens_r; a.camf.focus = (float)a.cam.focus;
	a.camf.jitter = a.cam.jitter;
	a.key = philox_key(p->seed);
	a.W = p->width; a.H = p->height;
	a.s_begin = p->sample_begin; a.s_count = p->sample_count;
	a.max_depth = p->max_depth;
	a.ao_samples = p->ao_samples > 0 ? p->ao_samples : 32;
	a.tmin = (float)p->t_min;
	for (int k = 0; k < 3; ++k) { a.bg_bottom[k] = (float)p->background_bottom[k]; a.bg_top[k] = (float)p->background_top[k]; }
	a.bg_black = 1;
	for (int k = 0; k < 3; ++k) a.bg_black &= (a.bg_bottom[k] == 0.0f && a.bg_top[k] == 0.0f) ? 1 : 0;
	a.lean = ctx->opt_lean ? 1 : 0;
	a.accum = accum;
	a.counters = ctx->d_counters;
	int mode = 0;
	if (p->integrator == ARE_INTEGRATOR_RT_AO) {
		// experiments/rt.cpp knows triangles only, tested by linear scan from shared memory
		if (ctx->csp->n_tri == 0 || ctx->csp->n_quad != 0 || ctx->csp->n_sph != 0 || ctx->csp->n_tri > 1000)
			return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "RT_AO integrator renders triangle-only scenes of at most 1000 triangles");
		make_rt_cam(cam->pos, cam->target, cam->up, cam->vfov_deg, p->width, p->height, a.rtcam);
	} else if (p->integrator != ARE_INTEGRATOR_PATH && p->integrator != ARE_INTEGRATOR_PATH_WAVEFRONT) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown integrator");
	else if (!resolve_traversal(ctx, p->traversal, mode)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown traversal, or scene too large for brute-force traversal");
	CK(cudaMemsetAsync(ctx->d_counters, 0, CNT_N * sizeof(unsigned long long), ctx->stream));
	if (stats) CK(cudaEventRecord(ctx->ev0, ctx->stream));
	int launched = 0;
	bool baked = false;
	if (p->sample_count > 0) {
		if (p->integrator == ARE_INTEGRATOR_RT_AO) launched = launch_render_rtao(a, ctx->stream);
		else if (p->integrator == ARE_INTEGRATOR_PATH_WAVEFRONT) {
			// ray queues: at most ~16 M rays in flight (1.7 GB), i.e. as many samples per pixel per batch as that allows
			const size_t frame = (size_t)p->width * p->height;
			const int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)p->sample_count, ((size_t)16 << 20) / frame));
			const size_t need = wavefront_workspace_bytes(p->width, p->height, batch);
			if (ctx->wf_ws_bytes < need) {
				if (ctx->wf_ws) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(ctx->wf_ws)); ctx->wf_ws = nullptr; ctx->wf_ws_bytes = 0; }
				CK(cudaMalloc(&ctx->wf_ws, need));
				ctx->wf_ws_bytes = need;
			}
			mode = 1;  // the wavefront extends through the BVH2
			launched = launch_render_wavefront(a, count_tests != 0, batch, ctx->wf_ws, ctx->stream);
		} else {
			int blocks, threads;
			size_t smem;
			// the baked kernel stands for the brute-force kernel it was generated around: the lean one, or the generic one
			baked = mode == 0 && ctx->baked && ctx->opt_bake && render_path_is_lean(a) == ctx->baked_lean && render_path_lean_dims(a, blocks, threads, smem);
			if (baked && !ctx->baked_lean) smem = (size_t)a.sc.n_hot * sizeof(HotPrim);
			if (baked) {
				std::string err;
				launched = bake_launch(ctx->baked, a, blocks, threads, smem, ctx->stream, err);
				if (launched < 0) return fail(ctx, ARE_ERR_CUDA, err);
			} else {
				// the hierarchy of a large scene lives in L2; optionally pin (a share of) the node array there
				// for the duration of the launch so that primitive records and accumulator traffic cannot evict it
				const bool persist = mode != 0 && ctx->opt_l2_persist > 0 && ctx->dev.n_nodes > 0 && ctx->l2_persist_max > 0;
				if (persist) {
					cudaStreamAttrValue av;
					std::memset(&av, 0, sizeof av);
					const size_t bytes = std::min((size_t)ctx->dev.n_nodes * sizeof(BvhNode), ctx->l2_window_max);
					av.accessPolicyWindow.base_ptr = const_cast<BvhNode *>(ctx->dev.nodes);
					av.accessPolicyWindow.num_bytes = bytes;
					// value = per cent of the persisting carve-out the window may claim (100: all of it)
					av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ctx->l2_persist_max * (ctx->opt_l2_persist / 100.0) / (double)bytes);
					av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
					av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
					CK(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av));
				}
				launched = launch_render_path(a, mode, count_tests != 0, ctx->stream);
				if (persist) {
					cudaStreamAttrValue av;
					std::memset(&av, 0, sizeof av);
					CK(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av));
				}
			}
		}
		if (launched < 0) return fail(ctx, ARE_ERR_CUDA, "render launch configuration rejected");
		CK(cudaGetLastError());
	}
	if (stats) {
		CK(cudaEventRecord(ctx->ev1, ctx->stream));
		unsigned long long c[CNT_N];
		CK(cudaMemcpyAsync(c, ctx->d_counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		std::memset(stats, 0, sizeof *stats);
		stats->samples = (uint64_t)p->width * p->height * p->sample_count;
		stats->rays = c[CNT_RAYS];
		if (p->integrator == ARE_INTEGRATOR_RT_AO) {  // rt.cpp's linear scan: every ray tests every triangle (rt.cpp:209-218)
			stats->tri_tests = c[CNT_RAYS] * (uint64_t)ctx->csp->n_tri;
		} else if (mode != 0) {
			stats->node_visits = c[CNT_NODES]; stats->quad_tests = c[CNT_QUADS]; stats->tri_tests = c[CNT_TRIS]; stats->sphere_tests = c[CNT_SPHERES];
			stats->box_tests = c[CNT_BOXES];
		} else {  // brute force: every ray tests every hot primitive — exact by construction
			stats->quad_tests = c[CNT_RAYS] * (uint64_t)ctx->csp->brute_range.nq;
			stats->tri_tests = c[CNT_RAYS] * (uint64_t)ctx->csp->brute_range.nt;
			stats->sphere_tests = c[CNT_RAYS] * (uint64_t)ctx->csp->brute_range.ns;
			stats->box_tests = c[CNT_RAYS] * (uint64_t)ctx->csp->brute_range.nb;
		}
		stats->kernel_ms = ms;
		stats->launches = (uint64_t)launched;
		stats->kernel_variant = launched <= 0 ? ARE_KERNEL_NONE
			: p->integrator == ARE_INTEGRATOR_RT_AO ? ARE_KERNEL_RT_AO
			: p->integrator == ARE_INTEGRATOR_PATH_WAVEFRONT ? ARE_KERNEL_WAVEFRONT
			: mode == 0 ? (baked ? ARE_KERNEL_BRUTE_BAKED : render_path_is_lean(a) ? ARE_KERNEL_BRUTE_LEAN : ARE_KERNEL_BRUTE)
			: mode == 2 ? ARE_KERNEL_WIDE : mode == 3 ? ARE_KERNEL_BVH4 : (render_path_is_big(a) ? (a.sc.nodes_q ? ARE_KERNEL_BVH2_QUANT : ARE_KERNEL_BVH2_BIG) : ARE_KERNEL_BVH2);
	}
	return ARE_OK;
}

static int ensure_own_accum(are_cuda_ctx *ctx, size_t elems) {
	if (ctx->own_accum_elems >= elems) return ARE_OK;
	if (ctx->own_accum) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(ctx->own_accum)); ctx->own_accum = nullptr; ctx->own_accum_elems = 0; }
	CK(cudaMalloc((void **)&ctx->own_accum, elems * sizeof(float)));
	ctx->own_accum_elems = elems;
	return ARE_OK;
}

// A multi-device group (are_cuda_create_multi): device i renders the i-th share of the sample range — the Philox counter
// carries the GLOBAL sample index, so the group draws exactly the samples one device would — the first
// device straight into `accum`, the others into their own zeroed scratch accumulators.  The sum is then formed by ONE
// kernel per device over NVLink peer memory (render.cu: k_peer_reduce): device d adds slice d of every scratch buffer
// into slice d of `accum`, so all N NVSwitch ports carry 1/N of the traffic each (a reduce-scatter whose outputs land
// directly in the root's buffer) instead of the root pulling N-1 whole frames.  Everything is ordered by events between
// the devices' streams; the host does not wait unless stats are requested.
static int render_multi(are_cuda_ctx *ctx, const are_camera *cam, const are_render_params *p, float *accum, are_render_stats *stats, int count_tests) {
	int st = need_commit(ctx);
	if (st) return st;
	if (!cam || !p || !accum) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (p->width <= 0 || p->height <= 0 || p->sample_count < 0) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad render parameters");
	std::vector<are_cuda_ctx *> all;
	all.push_back(ctx);
	for (are_cuda_ctx *c : ctx->peers) all.push_back(c);
	const int N = (int)all.size();
	const size_t elems = (size_t)p->width * p->height * 3;
	const int base = p->sample_count / N, rem = p->sample_count % N;
	for (int i = 1; i < N; ++i) {  // the previous round's reduce kernels (on every device) must be done with this scratch buffer
		are_cuda_ctx *c = all[i];
		Bind b(c);
		c->opt_lean = ctx->opt_lean; c->opt_bake = ctx->opt_bake;
		for (are_cuda_ctx *o : all) CK(cudaStreamWaitEvent(c->stream, o->ev_red, 0));
		if ((st = ensure_own_accum(c, elems)) != ARE_OK) return fail(ctx, st, c->err);
		CK(cudaMemsetAsync(c->own_accum, 0, elems * sizeof(float), c->stream));
	}
	for (int i = 0; i < N; ++i) {
		are_cuda_ctx *c = all[i];
		are_render_params pi = *p;
		pi.sample_begin = p->sample_begin + i * base + std::min(i, rem);
		pi.sample_count = base + (i < rem ? 1 : 0);
		Bind b(c);
		CK(cudaEventRecord(c->ev0, c->stream));
		st = render_single(c, cam, &pi, i == 0 ? accum : c->own_accum, nullptr, count_tests);
		if (st != ARE_OK) return i == 0 ? st : fail(ctx, st, "device " + std::to_string(c->device) + ": " + c->err);
		CK(cudaEventRecord(c->ev1, c->stream));
		CK(cudaEventRecord(c->ev_done, c->stream));
	}
	if (N > 1) {
		PeerReduceArgs ra;
		std::memset(&ra, 0, sizeof ra);
		ra.dst = accum;
		ra.n_src = N - 1;
		for (int i = 1; i < N; ++i) ra.src[i - 1] = all[i]->own_accum;
		const size_t n4 = (reinterpret_cast<uintptr_t>(accum) & 15) == 0 ? elems / 4 : 0;
		if (ctx->p2p_all) {
			for (int i = 0; i < N; ++i) {
				are_cuda_ctx *c = all[i];
				Bind b(c);
				for (are_cuda_ctx *o : all)
					if (o != c) CK(cudaStreamWaitEvent(c->stream, o->ev_done, 0));
				ra.begin4 = n4 * i / N; ra.end4 = n4 * (i + 1) / N;
				ra.tail_begin = i == N - 1 ? 4 * n4 : elems; ra.tail_end = elems;
				launch_peer_reduce(ra, c->sm_count, c->stream);
				CK(cudaGetLastError());
				CK(cudaEventRecord(c->ev_red, c->stream));
			}
			Bind b(ctx);
			for (int i = 1; i < N; ++i) CK(cudaStreamWaitEvent(ctx->stream, all[i]->ev_red, 0));
		} else {  // no peer mapping between some pair: stage every scratch frame through the first device
			Bind b(ctx);
			if (!ctx->stage || ctx->stage_elems < elems) {
				if (ctx->stage) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(ctx->stage)); ctx->stage = nullptr; }
				CK(cudaMalloc((void **)&ctx->stage, elems * sizeof(float)));
				ctx->stage_elems = elems;
			}
			for (int i = 1; i < N; ++i) {
				CK(cudaStreamWaitEvent(ctx->stream, all[i]->ev_done, 0));
				CK(cudaMemcpyPeerAsync(ctx->stage, ctx->device, all[i]->own_accum, all[i]->device, elems * sizeof(float), ctx->stream));
				PeerReduceArgs one = ra;
				one.n_src = 1; one.src[0] = ctx->stage;
				one.begin4 = 0; one.end4 = n4; one.tail_begin = 4 * n4; one.tail_end = elems;
				launch_peer_reduce(one, ctx->sm_count, ctx->stream);
				CK(cudaGetLastError());
			}
			CK(cudaEventRecord(ctx->ev_red, ctx->stream));
		}
	}
	if (stats) {
		std::memset(stats, 0, sizeof *stats);
		for (int i = 0; i < N; ++i) {
			are_cuda_ctx *c = all[i];
			Bind b(c);
			unsigned long long cn[CNT_N];
			CK(cudaMemcpyAsync(cn, c->d_counters, sizeof cn, cudaMemcpyDeviceToHost, c->stream));
			CK(cudaStreamSynchronize(c->stream));
			float ms = 0.f;
			CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
			stats->kernel_ms = std::max(stats->kernel_ms, (double)ms);
			stats->rays += cn[CNT_RAYS];
			stats->node_visits += cn[CNT_NODES]; stats->quad_tests += cn[CNT_QUADS]; stats->tri_tests += cn[CNT_TRIS];
			stats->sphere_tests += cn[CNT_SPHERES]; stats->box_tests += cn[CNT_BOXES];
			stats->launches += (base + (i < rem ? 1 : 0)) > 0 ? 1 : 0;
		}
		stats->launches += N > 1 ? (uint64_t)N : 0;  // the reduce kernels
		stats->samples = (uint64_t)p->width * p->height * p->sample_count;
		stats->kernel_variant = ARE_KERNEL_NONE;  // per-device variants may differ; query a single-device context for it
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return ARE_OK;
}

extern "C" {

int are_cuda_render_device(are_cuda_ctx *ctx, const are_camera *cam, const are_render_params *p, float *accum, are_render_stats *stats, int count_tests) {
	Range nvtx_range("are_cuda_render_device");
	if (ctx && !ctx->peers.empty()) return render_multi(ctx, cam, p, accum, stats, count_tests);
	return render_single(ctx, cam, p, accum, stats, count_tests);
}

int are_cuda_alloc_accum(are_cuda_ctx *ctx, int width, int height, float **out) {
	if (!ctx || !out || width <= 0 || height <= 0) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad argument");
	Bind b(ctx);
	void *d = nullptr;
	CK(cudaMalloc(&d, (size_t)width * height * 3 * sizeof(float)));
	CK(cudaMemsetAsync(d, 0, (size_t)width * height * 3 * sizeof(float), ctx->stream));
	*out = static_cast<float *>(d);
	return ARE_OK;
}
int are_cuda_zero_accum(are_cuda_ctx *ctx, float *accum, int width, int height) {
	if (!ctx || !accum) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad argument");
	Bind b(ctx);
	CK(cudaMemsetAsync(accum, 0, (size_t)width * height * 3 * sizeof(float), ctx->stream));
	return ARE_OK;
}
int are_cuda_download_accum(are_cuda_ctx *ctx, const float *accum, int width, int height, float *host) {
	if (!ctx || !accum || !host) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad argument");
	Bind b(ctx);
	CK(cudaMemcpyAsync(host, accum, (size_t)width * height * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}
int are_cuda_free_accum(are_cuda_ctx *ctx, float *accum) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	Bind b(ctx);
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaFree(accum));
	return ARE_OK;
}
int are_cuda_synchronize(are_cuda_ctx *ctx) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	Bind b(ctx);
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

int are_cuda_render(are_cuda_ctx *ctx, const are_camera *cam, const are_render_params *p, float *accum_host, are_render_stats *stats) {
	int st = need_commit(ctx);
	if (st) return st;
	if (!p || !accum_host) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (p->width <= 0 || p->height <= 0) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad render parameters");
	Bind b(ctx);
	const size_t elems = (size_t)p->width * p->height * 3;
	if ((st = ensure_own_accum(ctx, elems)) != ARE_OK) return st;
	CK(cudaMemsetAsync(ctx->own_accum, 0, elems * sizeof(float), ctx->stream));
	are_render_stats local;
	st = are_cuda_render_device(ctx, cam, p, ctx->own_accum, stats ? stats : &local, 0);
	if (st) return st;
	CK(cudaMemcpyAsync(accum_host, ctx->own_accum, elems * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

// thr[v-1] = the smallest float c in [0,1] with (unsigned char)(powf(c, 1/2.2f) * 255) >= v (experiments/rt.cpp:383-386),
// by bisection over the bit patterns of the non-negative floats with the host's powf — what rt.cpp itself calls.
static void gamma22_thresholds(float thr[256]) {
	auto enc = [](float c) { return (int)(unsigned char)(powf(fmaxf(0.0f, fminf(1.0f, c)), 1 / 2.2f) * 255); };
	auto as_float = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
	uint32_t one;
	const float f1 = 1.0f;
	std::memcpy(&one, &f1, 4);
	for (int v = 1; v <= 255; ++v) {
		uint32_t lo = 0, hi = one;  // enc(0) = 0 < v <= 255 = enc(1)
		while (hi - lo > 1) {
			const uint32_t mid = lo + (hi - lo) / 2;
			if (enc(as_float(mid)) >= v) hi = mid;
			else lo = mid;
		}
		thr[v - 1] = as_float(hi);
	}
	thr[255] = INFINITY;
}

int are_cuda_tonemap(are_cuda_ctx *ctx, const float *accum, int width, int height, double inv_spp, int encoder, uint8_t *out_host) {
	Range nvtx_range("are_cuda_tonemap");
	if (!ctx || !accum || !out_host || width <= 0 || height <= 0) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad argument");
	if (encoder < ARE_ENCODE_GAMMA22_TRUNC || encoder > ARE_ENCODE_SQRT_TRUNC) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "unknown encoder");
	Bind b(ctx);
	const size_t n = (size_t)width * height * 3;
	if (encoder == ARE_ENCODE_GAMMA22_TRUNC && !ctx->d_gamma_thr) {
		float thr[256];
		gamma22_thresholds(thr);
		CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_gamma_thr), sizeof thr));
		CK(cudaMemcpy(ctx->d_gamma_thr, thr, sizeof thr, cudaMemcpyHostToDevice));
	}
	if (ctx->d_rgb8_bytes < n) {
		if (ctx->d_rgb8) { CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(ctx->d_rgb8)); ctx->d_rgb8 = nullptr; ctx->d_rgb8_bytes = 0; }
		CK(cudaMalloc(reinterpret_cast<void **>(&ctx->d_rgb8), n));
		ctx->d_rgb8_bytes = n;
	}
	launch_tonemap(accum, width, height, inv_spp, encoder, ctx->d_gamma_thr, ctx->d_rgb8, ctx->stream);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out_host, ctx->d_rgb8, n, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

// Host part of Texture::paste: the homography through the four corner correspondences (8x8 Gauss-Jordan with partial
// pivoting, pivot threshold GEOMETRY_EPSILON) and its 3x3 inverse by cofactors — same elimination order as the
// reference (src/texture.cpp:196-300), so the coefficients handed to the kernel are the reference's bit for bit.
static bool paste_homography(const double sxy[4][2], const double dxy[4][2], double hinv[9]) {
	const double eps = 1e-12;
	double M[8][9];
	for (int k = 0; k < 4; ++k) {
		const double x = sxy[k][0], y = sxy[k][1], u = dxy[k][0], v = dxy[k][1];
		const double r0[9] = { x, y, 1.0, 0.0, 0.0, 0.0, -u * x, -u * y, u }, r1[9] = { 0.0, 0.0, 0.0, x, y, 1.0, -v * x, -v * y, v };
		std::memcpy(M[2 * k], r0, sizeof r0);
		std::memcpy(M[2 * k + 1], r1, sizeof r1);
	}
	for (int col = 0; col < 8; ++col) {
		int pivot = col;
		double best = std::fabs(M[col][col]);
		for (int r = col + 1; r < 8; ++r)
			if (std::fabs(M[r][col]) > best) { best = std::fabs(M[r][col]); pivot = r; }
		if (best < eps) return false;
		if (pivot != col)
			for (int c = col; c < 9; ++c) std::swap(M[col][c], M[pivot][c]);
		const double div = M[col][col];
		for (int c = col; c < 9; ++c) M[col][c] /= div;
		for (int r = 0; r < 8; ++r) {
			if (r == col) continue;
			const double factor = M[r][col];
			if (std::fabs(factor) < eps) continue;
			for (int c = col; c < 9; ++c) M[r][c] -= factor * M[col][c];
		}
	}
	const double a = M[0][8], b = M[1][8], c = M[2][8], d = M[3][8], e = M[4][8], f = M[5][8], g = M[6][8], h = M[7][8], i = 1.0;
	const double A = (e * i - f * h), B = -(d * i - f * g), C = (d * h - e * g), D = -(b * i - c * h), E = (a * i - c * g), F = -(a * h - b * g);
	const double G = (b * f - c * e), Hc = -(a * f - c * d), I = (a * e - b * d);
	const double det = a * A + b * B + c * C;
	if (std::fabs(det) < eps) return false;
	const double inv_det = 1.0 / det;
	const double out[9] = { A * inv_det, D * inv_det, G * inv_det, B * inv_det, E * inv_det, Hc * inv_det, C * inv_det, F * inv_det, I * inv_det };
	std::memcpy(hinv, out, sizeof out);
	return true;
}

int are_cuda_texture_paste(are_cuda_ctx *ctx, double *dst_rgb, int dst_w, int dst_h, const double *src_rgb, int src_w, int src_h, const int corners[8]) {
	Range nvtx_range("are_cuda_texture_paste");
	if (!ctx || !corners) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	if (!dst_rgb || dst_w <= 0 || dst_h <= 0) return fail(ctx, ARE_ERR_RUNTIME, "Texture is not initialized.");  // texture.cpp:90-92
	if (src_w <= 0 || src_h <= 0 || !src_rgb) return ARE_OK;                                                       // texture.cpp:93-95
	// corners arrive as LT, RT, LB, RB; the polygon test walks LT, RT, RB, LB
	const double lt[2] = { (double)corners[0], (double)corners[1] }, rt[2] = { (double)corners[2], (double)corners[3] };
	const double lb[2] = { (double)corners[4], (double)corners[5] }, rb[2] = { (double)corners[6], (double)corners[7] };
	PasteArgs a;
	const double qx[4] = { lt[0], rt[0], rb[0], lb[0] }, qy[4] = { lt[1], rt[1], rb[1], lb[1] };
	std::memcpy(a.qx, qx, sizeof qx);
	std::memcpy(a.qy, qy, sizeof qy);
	const double min_x = std::min(std::min(qx[0], qx[1]), std::min(qx[2], qx[3])), max_x = std::max(std::max(qx[0], qx[1]), std::max(qx[2], qx[3]));
	const double min_y = std::min(std::min(qy[0], qy[1]), std::min(qy[2], qy[3])), max_y = std::max(std::max(qy[0], qy[1]), std::max(qy[2], qy[3]));
	if (max_x < 0.0 || max_y < 0.0 || min_x > (double)(dst_w - 1) || min_y > (double)(dst_h - 1)) return ARE_OK;
	a.x0 = std::max(0, (int)std::floor(min_x)); a.x1 = std::min(dst_w - 1, (int)std::ceil(max_x));
	a.y0 = std::max(0, (int)std::floor(min_y)); a.y1 = std::min(dst_h - 1, (int)std::ceil(max_y));
	const double sw = (double)src_w, sh = (double)src_h;
	const double sxy[4][2] = { { 0.0, 0.0 }, { sw - 1.0, 0.0 }, { 0.0, sh - 1.0 }, { sw - 1.0, sh - 1.0 } };
	const double dxy[4][2] = { { lt[0], lt[1] }, { rt[0], rt[1] }, { lb[0], lb[1] }, { rb[0], rb[1] } };
	if (!paste_homography(sxy, dxy, a.hinv)) return ARE_OK;  // degenerate mapping: nothing pasted
	Bind b(ctx);
	const size_t dbytes = (size_t)dst_w * dst_h * 3 * sizeof(double), sbytes = (size_t)src_w * src_h * 3 * sizeof(double);
	Tmp ddst, dsrc;
	TMP_IN(ddst, dst_rgb, dbytes);
	TMP_IN(dsrc, src_rgb, sbytes);
	a.dst = ddst.as<double>(); a.src = dsrc.as<double>();
	a.dw = dst_w; a.dh = dst_h; a.sw = src_w; a.sh = src_h;
	launch_paste(a, ctx->stream);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(dst_rgb, ddst.p, dbytes, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return ARE_OK;
}

// ---- patch-as-viewport renderer (patch.cu) ------------------------------------------------------------------
static const char *patch_check(const are_patch_scene *sc, const are_patch_config *cfg) {
	if (!sc || !cfg) return "null argument";
	if (sc->n_tri < 0 || sc->n_mat < 0) return "negative count";
	if (sc->n_tri > 0 && (!sc->P || !sc->UV)) return "scene arrays missing";
	if (sc->n_mat > 0 && (!sc->mat_type || !sc->mat_albedo || !sc->mat_metalness)) return "material arrays missing";
	if (cfg->min_tex_res < 1 || cfg->max_tex_res < cfg->min_tex_res || cfg->max_tex_res > 16384) return "texture resolution limits out of range";
	if (cfg->max_depth < 0 || cfg->max_depth > 64) return "max_depth out of range";
	if (!(cfg->gamma > 0.0)) return "gamma must be positive";
	return nullptr;
}
static PatchSceneView patch_view(const are_patch_scene *sc) {
	PatchSceneView v;
	v.n_tri = sc->n_tri; v.n_mat = sc->n_mat; v.P = sc->P; v.UV = sc->UV; v.material = sc->material;
	v.mat_type = sc->mat_type; v.mat_albedo = sc->mat_albedo; v.mat_metalness = sc->mat_metalness;
	return v;
}
static PatchCfg patch_cfg(const are_patch_config *c) {
	PatchCfg o;
	o.max_depth = c->max_depth; o.min_area_px = c->min_area_px; o.max_res = c->max_tex_res; o.min_res = c->min_tex_res;
	std::memcpy(o.env, c->env, sizeof o.env);
	o.gamma = c->gamma;
	return o;
}
static void patch_stats_out(are_patch_stats *out, const PatchStats &st) {
	if (!out) return;
	out->nodes = st.nodes; out->node_texels = st.node_texels; out->ops = st.ops; out->levels = st.levels; out->launches = st.launches;
	out->h2d_bytes = st.h2d_bytes; out->d2h_bytes = st.d2h_bytes; out->plan_ms = st.plan_ms; out->kernel_ms = st.kernel_ms;
}
static double ms_since(const std::chrono::steady_clock::time_point &t0) {
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int are_cuda_patch_render(are_cuda_ctx *ctx, const are_patch_scene *scene, const double origin[3], const double viewport_P[18],
	const double viewport_UV[12], int width, int height, const are_patch_config *cfg, double *out_rgb, uint8_t *out_rgb8, are_patch_stats *stats) {
	Range nvtx_range("are_cuda_patch_render (plan + levels + camera)");
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (const char *why = patch_check(scene, cfg)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, why);
	if (!origin || !viewport_P || !viewport_UV || width < 2 || height < 2) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "camera arguments missing or image smaller than 2x2");
	Bind b(ctx);
	PatchStats st;
	PatchPlan plan;
	const auto t0 = std::chrono::steady_clock::now();
	patch_plan_camera(patch_view(scene), origin, viewport_P, viewport_UV, width, height, patch_cfg(cfg), plan);
	st.plan_ms = ms_since(t0);
	std::string err;
	if (patch_run_camera(ctx->patch_ws, plan, viewport_UV, width, height, patch_cfg(cfg), out_rgb, out_rgb8, st, ctx->stream, err)) return fail(ctx, ARE_ERR_CUDA, err);
	patch_stats_out(stats, st);
	return ARE_OK;
}

int are_cuda_patch_trace_texture(are_cuda_ctx *ctx, const are_patch_scene *scene, const double origin[3], int current, int tex_w, int tex_h,
	double est_area_px, const are_patch_config *cfg, double *out_tex, int out_wh[2], are_patch_stats *stats) {
	if (!ctx) return ARE_ERR_INVALID_ARGUMENT;
	if (const char *why = patch_check(scene, cfg)) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, why);
	if (!origin || !out_tex || !out_wh || current < 0 || current >= scene->n_tri) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "bad triangle index or null output");
	Bind b(ctx);
	PatchStats st;
	PatchPlan plan;
	const auto t0 = std::chrono::steady_clock::now();
	patch_plan_texture(patch_view(scene), origin, current, tex_w, tex_h, est_area_px, patch_cfg(cfg), plan);
	st.plan_ms = ms_since(t0);
	out_wh[0] = plan.root_w;
	out_wh[1] = plan.root_h;
	std::string err;
	if (patch_run_texture(ctx->patch_ws, plan, patch_cfg(cfg), out_tex, st, ctx->stream, err)) return fail(ctx, ARE_ERR_CUDA, err);
	patch_stats_out(stats, st);
	return ARE_OK;
}

int are_cuda_patch_plan_probe(const are_patch_scene *scene, const double origin[3], const double viewport_P[18], const double viewport_UV[12],
	int width, int height, const are_patch_config *cfg, uint64_t out[6]) {
	if (patch_check(scene, cfg) || !origin || !viewport_P || !viewport_UV || !out || width < 2 || height < 2) return ARE_ERR_INVALID_ARGUMENT;
	PatchPlan plan;
	patch_plan_camera(patch_view(scene), origin, viewport_P, viewport_UV, width, height, patch_cfg(cfg), plan);
	out[0] = plan.nodes.size();
	out[1] = (uint64_t)plan.arena_texels;
	out[2] = plan.ops.size();
	uint64_t lv = 0;
	for (const auto &l : plan.level_nodes) lv += !l.empty();
	out[3] = lv;
	out[4] = (uint64_t)(plan.vp_op_end[0] - plan.vp_op_begin[0]);
	out[5] = (uint64_t)(plan.vp_op_end[1] - plan.vp_op_begin[1]);
	return ARE_OK;
}

int are_cuda_write_ppm(const char *path, int width, int height, const uint8_t *rgb8) {
	if (!path || !rgb8 || width <= 0 || height <= 0) return ARE_ERR_INVALID_ARGUMENT;
	FILE *f = fopen(path, "wb");
	if (!f) return ARE_ERR_IO;
	fprintf(f, "P6\n%d %d\n255\n", width, height);  // src/texture.cpp:378, rt.cpp:379
	size_t n = (size_t)width * height * 3;
	bool ok = fwrite(rgb8, 1, n, f) == n;
	fclose(f);
	return ok ? ARE_OK : ARE_ERR_IO;
}

int are_cuda_measure_l2_peak(are_cuda_ctx *ctx, double *gb_per_s, uint64_t *buffer_bytes) {
	if (!ctx || !gb_per_s) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	Bind b(ctx);
	int l2 = 0;
	cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx->device);
	// a quarter of the L2 (B200: 126 MB in two partitions): resident whichever partition a line lands in
	const size_t bytes = std::max<size_t>((size_t)8 << 20, ((size_t)l2 / 4) & ~(size_t)0xfffff);
	Tmp buf, sink;
	TMP_OUT(buf, bytes);
	TMP_OUT(sink, 16);
	CK(cudaMemsetAsync(buf.p, 0, bytes, ctx->stream));
	double best = 0.0;
	for (int rep = 0; rep < 5; ++rep) {
		CK(cudaEventRecord(ctx->ev0, ctx->stream));
		const double moved = launch_l2_peak(buf.as<float>(), bytes, ctx->sm_count, 24, sink.as<float>(), ctx->stream);
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev1, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		if (rep >= 1 && ms > 0.f) best = std::max(best, moved / (ms * 1e-3) / 1e9);
	}
	*gb_per_s = best;
	if (buffer_bytes) *buffer_bytes = bytes;
	return ARE_OK;
}

int are_cuda_measure_fp32_peak(are_cuda_ctx *ctx, double *tflops, int *sm_count, int *sm_clock_khz) {
	if (!ctx || !tflops) return fail(ctx, ARE_ERR_INVALID_ARGUMENT, "null argument");
	Bind b(ctx);
	Tmp sink;
	TMP_OUT(sink, 16);
	double best = 0.0;
	for (int rep = 0; rep < 6; ++rep) {
		CK(cudaEventRecord(ctx->ev0, ctx->stream));
		double flops = launch_fp32_peak(sink.as<float>(), ctx->sm_count, 4096, ctx->stream);
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev1, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		if (rep >= 2 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
	}
	*tflops = best;
	if (sm_count) *sm_count = ctx->sm_count;
	if (sm_clock_khz) {
		int khz = 0;
		cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
		*sm_clock_khz = khz;
	}
	return ARE_OK;
}

}  // extern "C"
