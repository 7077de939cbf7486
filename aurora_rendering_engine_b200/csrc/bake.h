// bake.h — scene-specialised ("baked") render kernel.
//
// A small flat-shaded scene (CompiledScene::lean_ok: at most LEAN_MAX boxes / quad tests / triangle tests — the Cornell
// boxes of the reference's own experiments and of BASELINE configs 0 and 3) is a handful of constants.  The lean kernel
// reads them from shared memory behind twelve uniform guards; the baked kernel has them in its instruction stream:
// bake_source() writes the closest-hit tests of THIS scene as straight-line CUDA — every plane / slab coefficient an
// immediate operand, the zero components of every normal left out (an axis-aligned room costs three additions where
// the general slab test spends eighteen FFMA), no loads, no guards — and NVRTC compiles it around the very same
// render_path.cuh the precompiled kernels are built from.  Results are bit-identical to the lean and generic kernels.
//
// NVRTC (libnvrtc.so.12) and the driver API (libcuda.so.1) are dlopen()ed on first use: the library has no link-time
// dependency on either, and where they are missing the lean kernel keeps rendering (are_commit_info.baked says which).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "render_args.h"
#include "scene.h"

namespace areb {

// The generated translation unit for the lean form of `cs` (empty string when the scene has no lean form).
// packed: emit the (s, e) slab products as fma.rn.f32x2 pairs (FFMA2) where an axis has two or more non-zero components.
// min_blocks > 0: CTAs per SM the kernel is compiled for (__launch_bounds__; 0 = the lean kernel's RENDER_MIN_BLOCKS_LEAN).
// lean: the lean form (cs.lean_ok) around the lean kernel; else the whole brute-force list (boxes, parallelograms, triangles,
// spheres; at most BAKE_MAX_SLOTS hot slots) around the generic brute-force kernel with its general shading path.
enum { BAKE_MAX_SLOTS = 16 };
std::string bake_source(const CompiledScene &cs, bool lean, bool packed, int min_blocks = 0);

struct BakedKernel;  // one loaded module + function (per device), owned by the process-wide cache

// Compile (or fetch from the cache) the baked kernel of `cs` for the CURRENT device.  nullptr + err on failure
// (no NVRTC, compile error, ...).  compile_ms: time spent in NVRTC + module load, 0 on a cache hit.
const BakedKernel *bake_get(const CompiledScene &cs, bool lean, bool packed, int min_blocks, int device, std::string &err, double *compile_ms);

// The CUBIN image a baked kernel was loaded from (for cuobjdump / nvdisasm next to an ncu capture).
const std::string *bake_cubin(const BakedKernel *k);

// Launch on stream s.  Returns 1 (kernels launched) or -1 with err.
int bake_launch(const BakedKernel *k, const RenderArgs &a, int blocks, int threads, size_t smem, cudaStream_t s, std::string &err);

// NVRTC-only compile of a source (no GPU needed): returns the CUBIN bytes; used by the host-side probe / tests.
bool bake_compile_cubin(const std::string &src, std::string &cubin, std::string &log);

}  // namespace areb
