/* are_cuda.h — C ABI of the B200-native path-tracing core behind the Aurora Rendering Engine API.
 *
 * The reference (NanoEra/aurora-rendering-engine) has no FFI or plugin interface: its boundary is the C++
 * class API under include/basic, include/material, include/object and texture.h.  This header
 * is the thin extern "C" layer a host program calls after it has built its are::ObjectSet; every entry point
 * names the reference interface it stands in for.  Plain pointers and sizes only — no C++ or torch types.
 *
 *   scene description      replaces walking  are::ObjectSet::triangles         include/object/object_set.h:10-12
 *     are_cuda_add_triangle          <-      are::Triangle(Q,u,v,Material*,Texture*) include/object/triangle.h:19
 *     are_cuda_add_material          <-      are::Diffuse / are::Reflective    include/material/diffuse.h:11, reflective.h:11
 *     are_cuda_add_texture           <-      are::Texture (fill / PPM image)   include/texture.h:13-15
 *   per-ray harness
 *     are_cuda_hit_batch             <-      are::Object::intersect_ray        include/object/object.h:35 (src/object/triangle.cpp:82-121)
 *     are_cuda_scatter_batch         <-      are::reflect / are::refract       include/basic/vec3.h:79-80 (src/basic/vec3.cpp:182-199)
 *     are_cuda_camera_rays           <-      pinhole ray set-up                experiments/rt.cpp:339-343,364-366
 *   rendering
 *     are_cuda_render[_device]       <-      render()+trace()                  experiments/rt.cpp:251-374 (the reference's only sampling loop)
 *     are_cuda_tonemap               <-      writePPM / Texture::save_texture  experiments/rt.cpp:377-392, src/texture.cpp:362-395
 *
 * Conventions: all functions return ARE_OK (0) or a negative are_status; no exception crosses this boundary
 * (the C++ shim in include/are_cuda.hpp re-throws std::runtime_error / std::invalid_argument to match the
 * reference's error behaviour).  Vectors are 3 consecutive doubles, exactly the layout of are::Vec3
 * (double e_[3], include/basic/vec3.h:11).  A context drives one GPU (are_cuda_create) or several GPUs of one box
 * (are_cuda_create_multi); contexts are not thread-safe (the reference is single-threaded).  There is no CPU fallback: without a CUDA device are_cuda_create fails.
 */
#ifndef ARE_CUDA_H
#define ARE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARE_CUDA_ABI_VERSION 2

typedef struct are_cuda_ctx are_cuda_ctx;

typedef enum are_status {
	ARE_OK = 0,
	ARE_ERR_INVALID_ARGUMENT = -1, /* what the reference reports as std::invalid_argument */
	ARE_ERR_RUNTIME = -2, /* what the reference reports as std::runtime_error */
	ARE_ERR_CUDA = -3, /* CUDA runtime / launch failure; see are_cuda_last_error */
	ARE_ERR_NO_DEVICE = -4, /* no usable sm_100 device: there is deliberately no CPU path */
	ARE_ERR_NOT_COMMITTED = -5, /* scene changed since the last are_cuda_commit */
	ARE_ERR_IO = -6
} are_status;

typedef enum are_material_kind {
	ARE_MAT_DIFFUSE = 0, /* are::Diffuse — rendered as a Lambertian surface coloured by the primitive's texture */
	ARE_MAT_REFLECTIVE = 1, /* are::Reflective(reflectivity): mirror with probability r, Lambertian otherwise. p: [0]=reflectivity [1..3]=mirror tint */
	ARE_MAT_LAMBERTIAN = 2, /* p: [0]=texture override (-1: use the primitive's texture) */
	ARE_MAT_METAL = 3, /* p: [0]=fuzz [1]=texture override (-1: primitive's) */
	ARE_MAT_DIELECTRIC = 4, /* p: [0]=index of refraction */
	ARE_MAT_DIFFUSE_LIGHT = 5 /* p: [0]=texture override (-1: primitive's) [1]=radiance scale */
} are_material_kind;

typedef enum are_texture_kind {
	ARE_TEX_SOLID = 0, /* p: [0..2]=rgb                      (are::Texture(w,h,fill), rt.cpp SolidTexture :85-92) */
	ARE_TEX_CHECKER_UV = 1, /* p: [0]=scale [1..3]=even rgb [4..6]=odd rgb, on surface (u,v) (rt.cpp CheckerTexture :93-102) */
	ARE_TEX_CHECKER_3D = 2, /* p: [0]=scale [1..3]=even rgb [4..6]=odd rgb, on floor(p/scale) */
	ARE_TEX_NOISE = 3, /* p: [0]=scale [1]=table seed; 0.5*(1+sin(scale*p.z+10*turb(p,7))) */
	ARE_TEX_IMAGE = 4 /* rgb = w*h*3 doubles in [0,1], row-major image_[y][x] as are::Texture holds them (src/texture.cpp:31-47) */
} are_texture_kind;

typedef enum are_integrator {
	ARE_INTEGRATOR_PATH = 0, /* unbiased path tracer: camera jitter, spp loop, max_depth bounces */
	ARE_INTEGRATOR_RT_AO = 1, /* the reference's experiments/rt.cpp shading: primary + 32 AO rays + one mirror bounce (+cosine gather) */
	ARE_INTEGRATOR_PATH_WAVEFRONT = 2 /* the PATH estimator (same samples, same image up to summation order) scheduled as wavefront
	                                    stages over ray queues in HBM — generate / extend (BVH2) / shade + compact — instead of
	                                    one megakernel; kept for measurement: slower on every BASELINE scene */
} are_integrator;

typedef enum are_traversal {
	ARE_TRAVERSAL_AUTO = 0, /* brute force (scene-specialised kernel) up to 16 hot slots, else the BVH2 — through quantised 32-byte nodes beyond
	                           8192 nodes (ARE_OPT_QUANTIZED_NODES); the compressed wide BVH only above ARE_OPT_WIDE_MIN_NODES (default: never) */
	ARE_TRAVERSAL_BRUTE = 1,
	ARE_TRAVERSAL_BVH = 2, /* binary BVH, 64-byte nodes with both children's boxes */
	ARE_TRAVERSAL_WIDE = 3, /* compressed 8-wide BVH, 80-byte nodes with 8-bit child boxes (large scenes) */
	ARE_TRAVERSAL_BVH4 = 4 /* uncompressed 4-wide BVH, 128-byte nodes with four fp32 child boxes (needs ARE_OPT_BUILD_BVH4 and the
	                          host builder; BVH2 otherwise) */
} are_traversal;

typedef enum are_encoder {
	ARE_ENCODE_GAMMA22_TRUNC = 0, /* clamp, pow(c,1/2.2)*255, truncate      experiments/rt.cpp:383-386 */
	ARE_ENCODE_LINEAR_TRUNC = 1, /* clamp(c*255,0,255), truncate, no gamma  src/texture.cpp:384-386    */
	ARE_ENCODE_SQRT_TRUNC = 2 /* gamma 2: int(256*clamp(sqrt(c),0,0.999)) */
} are_encoder;

/* Pinhole / thin-lens camera.  Pixel (x,y), sub-pixel offset (sx,sy) in [0,1):
 *   fwd=normalize(target-pos); right=normalize(fwd x up); up'=right x fwd; s=tan(vfov/2); aspect=W/H
 *   fx=(2(x+sx)/W-1)*aspect*s; fy=(1-2(y+sy)/H)*s; dir=normalize(fwd+right*fx+up'*fy)
 * which with sx=sy=0.5 is the reference's ray set-up (experiments/rt.cpp:339-343,364-366).  With
 * defocus_angle>0 the origin is moved on a lens disk of radius focus_dist*tan(defocus_angle/2) and the ray
 * aims at pos+focus_dist*(fwd+right*fx+up'*fy). */
typedef struct are_camera {
	double pos[3];
	double target[3];
	double up[3];
	double vfov_deg;
	double focus_dist; /* > 0; 1.0 for a pinhole */
	double defocus_angle_deg; /* 0 = pinhole */
	int32_t jitter; /* 0: pixel centres (sx=sy=0.5); 1: sx,sy uniform in [0,1) */
	int32_t pad_;
} are_camera;

typedef struct are_render_params {
	int32_t width, height;
	int32_t sample_begin; /* first GLOBAL sample index rendered by this call */
	int32_t sample_count; /* number of samples per pixel rendered by this call */
	int32_t max_depth; /* max ray segments per path (path integrator) */
	int32_t integrator; /* are_integrator */
	int32_t traversal; /* are_traversal */
	int32_t ao_samples; /* RT_AO integrator: AO / gather rays per shaded point (reference: 32) */
	uint64_t seed; /* Philox4x32-10 key; counter = (pixel, global sample, bounce, stream) */
	double t_min; /* self-intersection guard: hits with t <= t_min are ignored */
	double background_bottom[3]; /* radiance of a missed ray: lerp(bottom, top, 0.5*(dir.y+1)) */
	double background_top[3];
} are_render_params;

/* Counters filled by the render kernels (exact, counted on the device). */
typedef struct are_render_stats {
	uint64_t samples; /* W*H*sample_count */
	uint64_t rays; /* every ray segment cast, incl. bounces / AO rays */
	uint64_t tri_tests, quad_tests, sphere_tests; /* ray-primitive tests executed */
	uint64_t node_visits; /* BVH child-box PAIRS tested: 1 per BVH2 node visit, 4 per 8-wide node visit */
	uint64_t box_tests; /* parallelepiped (three slab pairs) tests: boxes detected among the scene's parallelograms */
	double kernel_ms; /* device time of the render kernel(s), CUDA events on the launch stream */
	uint64_t launches; /* kernels launched by this call */
	uint64_t kernel_variant; /* which render kernel ran: ARE_KERNEL_* */
} are_render_stats;
enum {
	ARE_KERNEL_NONE = 0,
	ARE_KERNEL_BRUTE = 1, /* generic brute-force list from shared memory */
	ARE_KERNEL_BRUTE_LEAN = 2, /* small flat-shaded scene: unrolled tests, per-face shading records in shared memory */
	ARE_KERNEL_BVH2 = 3,
	ARE_KERNEL_BVH2_BIG = 4, /* high-occupancy build for hierarchies that live in L2 */
	ARE_KERNEL_WIDE = 5,
	ARE_KERNEL_RT_AO = 6,
	ARE_KERNEL_BVH4 = 9, /* k_render_path over the 4-wide hierarchy */
	ARE_KERNEL_WAVEFRONT = 8, /* k_wf_generate / k_wf_extend / k_wf_shade */
	ARE_KERNEL_BVH2_QUANT = 10, /* big hierarchy through its quantised 32-byte nodes: one 256-bit load per node visit */
	ARE_KERNEL_BRUTE_BAKED = 7 /* the lean kernel with the scene's closest-hit tests compiled in (NVRTC at commit) */
};

/* ---- context ------------------------------------------------------------------------------------------- */
int are_cuda_abi_version(void);
int are_cuda_device_count(void); /* number of visible CUDA devices, 0 if none / no driver */
int are_cuda_create(are_cuda_ctx **out, int device);
/* One context driving SEVERAL GPUs of one box in this process: the scene is compiled once and
 * committed to every device; are_cuda_render / are_cuda_render_device split the sample range over the devices (device i
 * renders the i-th share of [sample_begin, sample_begin + sample_count); the Philox counter carries the global sample
 * index, so the group draws exactly the samples one device would) and return the SUM — in a buffer on devices[0], which
 * is also the device every other entry point (per-ray batches, tonemap, patch renderer) runs on.  The sum is formed by
 * one kernel per device over NVLink peer memory (each device adds one slice of all frames into the root's buffer); where
 * two devices cannot map each other the frames are staged through the first.  No MPI, NCCL or host threads involved. */
#define ARE_MAX_GROUP_DEVICES 16
int are_cuda_create_multi(are_cuda_ctx **out, const int *devices, int n_devices);
/* n_devices of the group (1 for are_cuda_create), peer_mapped = 1 when the peer-memory reduce is in use. */
int are_cuda_group_info(are_cuda_ctx *ctx, int *n_devices, int *peer_mapped);
void are_cuda_destroy(are_cuda_ctx *ctx);
const char *are_cuda_last_error(are_cuda_ctx *ctx); /* ctx may be NULL: last error of a failed create */
/* Use an existing CUDA stream (cudaStream_t passed as void*; NULL = the legacy default stream) for all work. */
int are_cuda_set_stream(are_cuda_ctx *ctx, void *cuda_stream);

/* Who builds the BVH at the next commit.  The reference has no hierarchy (its ObjectSet is scanned linearly,
 * include/object/object_set.h:10-12); both builders are NEW.
 *   HOST_SAH     binned surface-area heuristic on all host threads (default): best trees, ~0.75 s per million primitives
 *   DEVICE_LBVH  Morton codes + radix sort + Karras' radix tree + bottom-up boxes on the GPU: milliseconds per million
 *                primitives, somewhat slower traversal — for scenes that change every frame.  Falls back to HOST_SAH when
 *                the tree would be deeper than the traversal stack (are_commit_info.builder tells). */
enum { ARE_BVH_BUILDER_HOST_SAH = 0, ARE_BVH_BUILDER_DEVICE_LBVH = 1 };
int are_cuda_set_bvh_builder(are_cuda_ctx *ctx, int builder);
typedef struct are_commit_info {
	int builder; /* ARE_BVH_BUILDER_* actually used by the last commit */
	int bvh_nodes, bvh_height, hot_slots;
	double host_compile_ms; /* flattening, fusion, box detection (+ the host BVH build under HOST_SAH) */
	double host_bvh_ms; /* of which: the host BVH builders (HOST_SAH) / preparing the device builder's input (DEVICE_LBVH) */
	double device_bvh_ms; /* DEVICE_LBVH: CUDA-event time of the build kernels */
	uint64_t device_bvh_launches;
	int32_t baked; /* a scene-specialised render kernel is loaded for this scene (ARE_OPT_BAKED_KERNEL): 1 = around the lean kernel,
	                  2 = around the generic brute-force kernel (scenes of at most 16 hot slots: spheres, textures, any material), 0 = none */
	int32_t quant_area_permille; /* quantised 32-byte nodes (ARE_OPT_QUANTIZED_NODES): 0 = not built; else 1000 x the mean over all child
	                                boxes of (surface area as quantised : surface area of the fp32 box) (1 M-primitive test scene: ~1040).  Above 1250 the grid
	                                is too coarse for the scene's small boxes and renders use the fp32 nodes */
	double bake_compile_ms; /* NVRTC + module load time of this commit; 0 when the kernel came from the process-wide cache */
} are_commit_info;
int are_cuda_get_commit_info(are_cuda_ctx *ctx, are_commit_info *out);

/* Context options (what used to be ARE_CUDA_* environment switches).  Scene-compiler options (3-6, 8) take effect at the
 * next are_cuda_commit; kernel-choice options (1, 2, 7) at the next render. */
typedef enum are_option {
	ARE_OPT_LEAN_KERNEL = 1, /* 1 (default): small flat-shaded scenes use the lean brute-force kernel; 0: the generic one */
	ARE_OPT_BAKED_KERNEL = 2, /* 1 (default): ... and the scene-specialised form of either (lean form, or any brute-force list of at most
	                             16 hot slots), generated and compiled with NVRTC at commit:
	                             every plane / slab coefficient an immediate, zero components left out, no loads, no guards.
	                             Bit-identical output; silently absent where NVRTC / the driver API cannot be loaded */
	ARE_OPT_BAKED_PACKED = 3, /* 0 (default) / 1: baked slab products as fma.rn.f32x2 pairs (FFMA2) */
	ARE_OPT_FUSE_PARALLELOGRAMS = 4, /* 1 (default): coplanar triangle pairs forming a parallelogram become one test */
	ARE_OPT_FUSE_BOXES = 5, /* 1 (default): parallelograms forming a parallelepiped become one slab test */
	ARE_OPT_BUILD_WIDE = 6, /* 0 (default) / 1: also build the compressed 8-wide BVH for scenes beyond 65536 nodes */
	ARE_OPT_WIDE_MIN_NODES = 7, /* ARE_TRAVERSAL_AUTO picks the 8-wide BVH above this node count (default: never) */
	ARE_OPT_LBVH_MAX_HEIGHT = 8, /* test hook: device-built trees taller than this fall back to the host builder */
	ARE_OPT_L2_PERSIST_NODES = 9, /* 0 (default) / 1..100: BVH renders mark the node array as an L2-persisting access window
	                                (cudaAccessPolicyWindow) claiming this per cent of the device's persisting carve-out */
	ARE_OPT_BUILD_BVH4 = 10, /* 0 (default) / 1: the host builder also collapses its BVH2 into 4-wide nodes (ARE_TRAVERSAL_BVH4) */
	ARE_OPT_BAKED_MIN_BLOCKS = 11, /* tuning: CTAs per SM the baked kernel is compiled for (0 = default: 6 lean, 5 generic) */
	ARE_OPT_QUANTIZED_NODES = 12 /* 1 (default): BVH2 hierarchies of more than 8192 nodes are also stored as 32-byte nodes — child boxes as
	                                16-bit planes on a grid over the root box, padded outwards, so every hit the fp32 nodes report is still
	                                found — and ARE_TRAVERSAL_BVH2 renders through them (ARE_KERNEL_BVH2_QUANT); 0: fp32 nodes only
	                                (renders: at once; the copy itself: from the next commit) */
} are_option;
int are_cuda_set_option(are_cuda_ctx *ctx, int option, int value);
/* The sm_100a CUBIN of the committed scene's baked kernel (what cuobjdump -sass / nvdisasm -g read next to an ncu capture).
 * *size = its length; copied to out when cap suffices.  ARE_ERR_RUNTIME when the scene has none (last_error says why). */
int are_cuda_get_baked_cubin(are_cuda_ctx *ctx, void *out, uint64_t cap, uint64_t *size);
/* Host threads of the scene compiler, process-wide (0 = all hardware threads).  The compiled scene never depends on it. */
void are_cuda_set_build_threads(int n);

/* ---- scene description (host side, cheap; nothing touches the GPU until commit) ---------------------- */
/* Each add_* returns the new non-negative id, or a negative are_status. */
int are_cuda_add_texture(are_cuda_ctx *ctx, int kind, const double params[8], const double *rgb, int w, int h);
int are_cuda_add_material(are_cuda_ctx *ctx, int kind, const double params[8]);
/* Same validation as are::Triangle's ctor (src/object/triangle.cpp:9-46): u, v and u x v must not be near_zero. */
int are_cuda_add_triangle(are_cuda_ctx *ctx, const double Q[3], const double u[3], const double v[3], int material_id, int texture_id);
/* Optional texture coordinates of the three vertices Q, Q+u, Q+v (default (0,0),(1,0),(0,1)). */
int are_cuda_set_triangle_uv(are_cuda_ctx *ctx, int prim_id, const double uv[6]);
int are_cuda_add_quad(are_cuda_ctx *ctx, const double Q[3], const double u[3], const double v[3], int material_id, int texture_id);
int are_cuda_add_sphere(are_cuda_ctx *ctx, const double center[3], double radius, int material_id, int texture_id);
/* Bulk forms for large scenes (arrays of n items, 3 doubles per vector); return the id of the first item added. */
int are_cuda_add_triangles(are_cuda_ctx *ctx, int n, const double *Q, const double *u, const double *v, const int *material_id, const int *texture_id);
int are_cuda_add_spheres(are_cuda_ctx *ctx, int n, const double *center, const double *radius, const int *material_id, const int *texture_id);
int are_cuda_clear(are_cuda_ctx *ctx);
int are_cuda_num_primitives(are_cuda_ctx *ctx);
/* Flatten to SoA, build the BVH (host, binned SAH) and upload. Returns bytes uploaded via *h2d_bytes (may be NULL). */
int are_cuda_commit(are_cuda_ctx *ctx, uint64_t *h2d_bytes);

/* Moving primitives of a committed scene without rebuilding its hierarchy (refit).  update_* replace the
 * geometry of existing primitives (same ids, same types, same validation as the add_* calls); are_cuda_refit then writes
 * the moved primitives' records into their leaves and recomputes every node box bottom-up over the UNCHANGED topology of
 * the device-built tree: three kernels, well under a millisecond per million primitives, against a full commit.  The
 * tree loses quality as primitives move far; commit again from time to time.  Requires ARE_BVH_BUILDER_DEVICE_LBVH and a
 * scene whose hot items are its primitives (no fused parallelograms / boxes) — otherwise ARE_ERR_RUNTIME, and the caller
 * commits instead (the updates are kept).  device_ms (may be NULL): CUDA-event time of the refit kernels. */
int are_cuda_update_triangles(are_cuda_ctx *ctx, int n, const int *prim_ids, const double *Q, const double *u, const double *v);
int are_cuda_update_spheres(are_cuda_ctx *ctx, int n, const int *prim_ids, const double *center, const double *radius);
int are_cuda_refit(are_cuda_ctx *ctx, double *device_ms);

/* Host-only probe of the scene compiler (no GPU needed): flattens n_tri triangles exactly as are_cuda_commit would
 * and reports out[8] = { hot slots, fused triangle pairs, boxes, BVH nodes, BVH depth, brute quads, brute triangles,
 * brute boxes }.  Lets CPU-only test boxes check parallelogram fusion / box detection / BVH construction. */
int are_cuda_compile_probe(int n_tri, const double *Q, const double *u, const double *v, int out[8]);
/* Same, plus an FNV-1a digest of the compiled hierarchy (nodes, leaf-ordered primitives, ids): the BVH is built by
 * several host threads (are_cuda_set_build_threads overrides the count) and must not depend on how many. */
int are_cuda_compile_probe_digest(int n_tri, const double *Q, const double *u, const double *v, int out[8], uint64_t *digest);
/* Host-only probe of the two derived forms: compiles with the device BVH builder selected and reports out[8] =
 * { lean form valid, lean shading records, open boxes leading the brute list, device-builder items, their slots,
 *   items whose fp32 box contains the fp64 vertices behind it (must equal the item count), host BVH nodes (0: the
 *   hierarchy is left to the device), brute boxes }. */
int are_cuda_compile_probe_forms(int n_tri, const double *Q, const double *u, const double *v, int out[8]);

/* Host-only probe of the scene-specialised ("baked") render kernel (no GPU needed): compiles the n_tri triangles as
 * are_cuda_commit would and writes the CUDA source generated for the lean form of that scene to source_out (at most
 * source_cap bytes, NUL-terminated; *source_len = full length).  packed: bit 0 = slab products as fma.rn.f32x2 pairs, bit 1 =
 * generate around the GENERIC brute-force kernel even if the scene has a lean form.
 * cubin_path != NULL: additionally compile it with NVRTC for sm_100a and write the CUBIN there (cuobjdump -sass reads
 * it).  ARE_ERR_RUNTIME when the scene has no lean form or NVRTC fails (are_cuda_last_error(NULL) has the log). */
int are_cuda_bake_probe(int n_tri, const double *Q, const double *u, const double *v, int packed, char *source_out, uint64_t source_cap,
	uint64_t *source_len, const char *cubin_path);

/* ---- per-ray harness ----------------------------------------------------------------------------------- */
/* Closest hit of n rays against the committed scene.  D is normalised first, as are::Ray's ctor does
 * (src/basic/ray.cpp:5).  precision: 64 = fp64 kernels with the reference's GEOMETRY_EPSILON decisions,
 * 32 = the fp32 device routines the renderer itself uses.  traversal: are_traversal.
 * Outputs (host pointers, any may be NULL): prim[n] (-1 = miss), t[n], P[3n] hit point, N[3n] geometric unit
 * normal (not flipped), uv[2n] surface coordinates.  Miss => t, P, N, uv are NaN. */
int are_cuda_hit_batch(are_cuda_ctx *ctx, int n, const double *ray_Q, const double *ray_D, double t_min,
	int precision, int traversal, int *prim, double *t, double *P, double *N, double *uv);

/* Material response for n surface interactions with explicit random numbers rnd[4n] in [0,1).
 * wi = unit incoming direction, N = geometric unit normal (either side), P = hit point, uv = surface coords.
 * Outputs: wo[3n] unit scattered direction (NaN when !alive), attenuation[3n], emitted[3n], alive[n]. */
int are_cuda_scatter_batch(are_cuda_ctx *ctx, int n, const int *material_id, const int *texture_id,
	const double *wi, const double *N, const double *P, const double *uv, const double *rnd, int precision,
	double *wo, double *attenuation, double *emitted, int *alive);

/* Texture evaluation at n points (texture_id[n], uv[2n], P[3n]) -> rgb[3n]. */
int are_cuda_texture_batch(are_cuda_ctx *ctx, int n, const int *texture_id, const double *uv, const double *P,
	int precision, double *rgb);

/* Camera rays for n (pixel x, pixel y, sx, sy, lens r0, lens r1) tuples: px[n], py[n], rnd[4n]. */
int are_cuda_camera_rays(are_cuda_ctx *ctx, const are_camera *cam, int width, int height, int n,
	const int *px, const int *py, const double *rnd, int precision, double *ray_Q, double *ray_D);

/* Library routines that are not on the render loop, as fp64 batches with the reference's decisions bit for bit
 * (SURVEY.md §8a rows a5, a8, a10).  None needs a committed scene.
 *   plane_batch            are::Plane::intersect_ray        src/basic/plane.cpp:13-27   plane4 = (normal, d) per ray;
 *                          D is normalised first; hit[n], P[3n] (NaN on a miss)
 *   point_in_batch         are::Triangle::point_in          src/object/triangle.cpp:49-79   one triangle (Q,u,v), n points
 *   material_reflect_batch are::Material::reflect           src/material/diffuse.cpp:5-7, reflective.cpp:9-26
 *                          kind = ARE_MAT_DIFFUSE (always declines) or ARE_MAT_REFLECTIVE (mirrors the viewport origin
 *                          across the plane); ok[n], new_origin[3n] (NaN when declined) */
int are_cuda_plane_batch(are_cuda_ctx *ctx, int n, const double *plane4, const double *ray_Q, const double *ray_D, int *hit, double *P);
int are_cuda_point_in_batch(are_cuda_ctx *ctx, const double Q[3], const double u[3], const double v[3], int n, const double *points, int *inside);
int are_cuda_material_reflect_batch(are_cuda_ctx *ctx, int kind, int n, const double *plane4, const double *origin, int *ok, double *new_origin);

/* The counter-based generator itself: out[4n] = Philox4x32-10(key = seed, counter[4n]). */
int are_cuda_philox_batch(are_cuda_ctx *ctx, int n, uint64_t seed, const uint32_t *counter, uint32_t *out);

/* ---- rendering ----------------------------------------------------------------------------------------- */
/* Adds the radiance SUM of samples [sample_begin, sample_begin+sample_count) of every pixel into
 * accum_rgb_device (W*H*3 floats, row-major, device pointer on this context's GPU; the caller zeroes it before
 * the first call and divides by the total spp at the end).  Asynchronous on the context's stream unless
 * stats != NULL, in which case the call synchronises and fills *stats (count_tests != 0 additionally runs the
 * counting variant of the kernel so tri/quad/sphere/node counters are exact; rays and samples are always counted). */
int are_cuda_render_device(are_cuda_ctx *ctx, const are_camera *cam, const are_render_params *params,
	float *accum_rgb_device, are_render_stats *stats, int count_tests);

/* Whole call with HOST buffers: zero a device accumulator, render, copy W*H*3 floats (sample sums) back. */
int are_cuda_render(are_cuda_ctx *ctx, const are_camera *cam, const are_render_params *params,
	float *accum_rgb_host, are_render_stats *stats);

/* 8-bit encode of accum/spp on the device; out_rgb8_host = W*H*3 bytes, P6 payload order. */
int are_cuda_tonemap(are_cuda_ctx *ctx, const float *accum_rgb_device, int width, int height, double inv_spp,
	int encoder, uint8_t *out_rgb8_host);
/* Write a binary P6 file exactly as the reference does ("P6\n%d %d\n255\n" + payload). */
int are_cuda_write_ppm(const char *path, int width, int height, const uint8_t *rgb8);

/* are::Texture::paste on the GPU (reference src/texture.cpp:85-360): warps the src image into the quad
 * (left_top, right_top, left_bottom, right_bottom) of the dst image — 4-corner homography, even-odd point-in-quad at
 * pixel centres, bilinear fetch, fp64; only pixels inside the quad whose pre-image lies inside src are overwritten.
 * dst_rgb (in/out) and src_rgb are host arrays of w*h*3 doubles, rows as are::Texture holds them.
 * corners = { lt.x, lt.y, rt.x, rt.y, lb.x, lb.y, rb.x, rb.y }.  Degenerate mappings leave dst untouched (ARE_OK). */
int are_cuda_texture_paste(are_cuda_ctx *ctx, double *dst_rgb, int dst_w, int dst_h, const double *src_rgb, int src_w, int src_h,
	const int corners[8]);

/* ---- patch-as-viewport renderer ------------------------------------------------------------------------ */
/* The reference's OWN rendering algorithm (not ray tracing): what are::Object::trace_texture(object_set,
 * viewport_origin) -> Texture is declared for (reference include/object/object.h:37-38; never defined there) and
 * what its prototype experiments/rt10.cpp implements: a reflective triangle is a viewport seen from the eye point
 * mirrored across its plane (are::Reflective::reflect, src/material/reflective.cpp:9-26); the triangles visible
 * through it are projected, clipped, painted far-to-near into its texture, recursively (rt10.cpp:551-664), and the
 * camera is a two-triangle viewport (rt10.cpp:677-772).  The host plans footprints, the GPU produces every texel in
 * fp64 with the reference's operation order: output is bit-identical to rt10.cpp (its shipped
 * experiments/output_rt10.ppm reproduces byte for byte). */
typedef struct are_patch_scene {
	int32_t n_tri, n_mat;
	const double *P; /* n_tri*9: three vertices per triangle                       rt10.cpp:152-156 */
	const double *UV; /* n_tri*6: texture coordinates of the vertices */
	const int32_t *material; /* n_tri: material index; out of range = none (white, diffuse) */
	const int32_t *mat_type; /* n_mat: 0 = Diffuse, 1 = Reflective                         rt10.cpp:140-149 */
	const double *mat_albedo; /* n_mat*3 */
	const double *mat_metalness; /* n_mat: reflection / base colour mix, clamped to [0,1] */
} are_patch_scene;

typedef struct are_patch_config { /* RenderConfig, rt10.cpp:536-542 */
	int32_t max_depth;
	int32_t max_tex_res, min_tex_res;
	int32_t pad_;
	double min_area_px; /* footprints smaller than this (in pixels) are not recursed into */
	double env[3]; /* colour where nothing is reflected */
	double gamma; /* 8-bit encode: lround(255*pow(clamp(c,0,1), 1/gamma))     rt10.cpp:118-143 */
} are_patch_config;

typedef struct are_patch_stats {
	uint64_t nodes, node_texels, ops, levels, launches;
	uint64_t h2d_bytes, d2h_bytes;
	double plan_ms; /* host: footprint planning */
	double kernel_ms; /* device: all launches of the call, CUDA events on the context's stream */
} are_patch_stats;

/* Camera::render (rt10.cpp:755-772).  viewport_P[18] / viewport_UV[12]: the two viewport triangles.  Outputs (host,
 * either may be NULL): out_rgb = W*H*3 doubles (linear, clamped to [0,1]); out_rgb8 = W*H*3 bytes, the P6 payload. */
int are_cuda_patch_render(are_cuda_ctx *ctx, const are_patch_scene *scene, const double origin[3], const double viewport_P[18],
	const double viewport_UV[12], int width, int height, const are_patch_config *cfg, double *out_rgb, uint8_t *out_rgb8, are_patch_stats *stats);

/* renderTriangleWithTriangle (rt10.cpp:551-664) = Object::trace_texture for scene triangle `current` seen from
 * `origin`.  The texture size is clamped to [min_tex_res, max_tex_res] and returned in out_wh; out_tex must hold
 * max_tex_res^2*3 doubles.  est_area_px: projected size hint (0 = unknown; < min_area_px stops the recursion). */
int are_cuda_patch_trace_texture(are_cuda_ctx *ctx, const are_patch_scene *scene, const double origin[3], int current, int tex_w, int tex_h,
	double est_area_px, const are_patch_config *cfg, double *out_tex, int out_wh[2], are_patch_stats *stats);

/* Host-only probe of the planner (no GPU needed): out[6] = { reflective nodes, node texels, warp triangles, levels,
 * warp triangles of viewport A, of viewport B }. */
int are_cuda_patch_plan_probe(const are_patch_scene *scene, const double origin[3], const double viewport_P[18], const double viewport_UV[12],
	int width, int height, const are_patch_config *cfg, uint64_t out[6]);

/* Device scratch owned by the context (so hosts without a CUDA allocator can still drive render_device). */
int are_cuda_alloc_accum(are_cuda_ctx *ctx, int width, int height, float **accum_rgb_device);
int are_cuda_zero_accum(are_cuda_ctx *ctx, float *accum_rgb_device, int width, int height);
int are_cuda_download_accum(are_cuda_ctx *ctx, const float *accum_rgb_device, int width, int height, float *host);
int are_cuda_free_accum(are_cuda_ctx *ctx, float *accum_rgb_device);
int are_cuda_synchronize(are_cuda_ctx *ctx);

/* Measured FP32 FMA issue peak of this device (TFLOP/s): a register-resident FFMA micro-kernel, used as the
 * roofline denominator next to the nominal SMs x 128 lanes x 2 x clock figure. */
int are_cuda_measure_fp32_peak(are_cuda_ctx *ctx, double *tflops, int *sm_count, int *sm_clock_khz);
/* Measured L2 read bandwidth (GB/s): every CTA streams a buffer of a quarter of the L2 with ld.global.cg.v4 — the
 * roofline denominator of the traversal kernels whose hierarchy lives in L2.  buffer_bytes may be NULL. */
int are_cuda_measure_l2_peak(are_cuda_ctx *ctx, double *gb_per_s, uint64_t *buffer_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ARE_CUDA_H */
