// path_shade.cuh — one bounce of the path estimator for ONE ray whose closest hit is known: classify (miss / light /
// surface), attenuate, scatter.  Exactly the classify + scatter stages of the megakernel's loop (render_path.cuh) over
// the BVH arrays, in the form the queue-based schedules need (wavefront.cu: queues in HBM; pool.cu: queues in shared
// memory).  Same Philox counter (pixel, global sample, bounce), same device routines, so all schedules trace the same rays.
#pragma once
#include "render_path.cuh"

namespace areb {

struct PathRay {
	F3 o, d, thr;
	int bounce;  // segments traced so far + 1 (1 = the camera ray)
	int orig;    // hot slot the ray starts on (-1: none)
};

// In: the ray and its hit (idx < 0: miss).  Out: done (the path ended, `contrib` is its radiance) or alive (r holds the
// continuing ray) or neither (absorbed).
__device__ __forceinline__ void shade_one(const RenderArgs &A, PathRay &r, float t, int idx, uint32_t pixel, uint32_t sample, bool &done, bool &alive, F3 &contrib) {
	contrib = mk<float>(0.f, 0.f, 0.f);
	alive = false;
	if (idx < 0) {
		if (!A.bg_black) contrib = r.thr * background(A, r.d);
		done = true;
		return;
	}
	const F3 sP = r.o + t * r.d;
	r.orig = idx;
	const HotIds id = hit_ids<ldg4>(A.sc, A.sc.bvh_prims, A.sc.bvh_ids, idx, sP);
	const int sdev = id.b >= 0 ? resolve_exact(A.sc, id, sP).dev_prim : id.a;
	const float4 s0 = __ldg(reinterpret_cast<const float4 *>(&A.sc.shade[sdev].r0));
	const float4 s1 = __ldg(reinterpret_cast<const float4 *>(&A.sc.shade[sdev].r1));
	const int sbits = __float_as_int(s0.w);
	F3 sN = mk<float>(s0.x, s0.y, s0.z), scol = mk<float>(s1.x, s1.y, s1.z);
	if (sdev >= A.sc.n_tri + A.sc.n_quad) {  // sphere: outward normal from (centre, radius)
		const float4 cc = __ldg(reinterpret_cast<const float4 *>(&A.sc.prim_plane[sdev].r0));
		sN = (1.0f / cc.w) * (sP - mk<float>(cc.x, cc.y, cc.z));
	}
	const int kind = sbits & 255;
	const bool fast = (sbits >> 8) & 1;
	if (kind == MK_LIGHT) {
		if (!fast) {  // textured light: general path
			const Resolved rs = resolve_exact(A.sc, id, sP);
			const PrimInfo pi = A.sc.info[rs.dev_prim];
			float u, v;
			surface_at(A.sc, rs.dev_prim, sP, rs.a, rs.b, sN, u, v);
			const MaterialRec &m = A.sc.mats[pi.mat];
			scol = m.pf[1] * tex_eval<float>(A.sc, mat_texture(m, pi.tex), u, v, sP);
		}
		contrib = r.thr * scol;
		done = true;
		return;
	}
	done = r.bounce >= A.max_depth;  // a truncated path contributes nothing
	if (fast) r.thr = r.thr * scol;
	if (done) return;
	const Rnd4<float> rn = rnd4<float>(A.key, pixel, sample, (uint32_t)r.bounce, 0u);
	F3 wo;
	if (fast) alive = scatter_dir<float>(kind, s1.w, r.d, sN, rn, wo);
	else {
		const Resolved rs = resolve_exact(A.sc, HotIds{ sdev, -1 }, sP);
		const PrimInfo pi = A.sc.info[rs.dev_prim];
		float u, v;
		F3 att, emit;
		const int tk = (sbits >> SHADE_TEXKIND_SHIFT) & 7;
		surface_at(A.sc, rs.dev_prim, sP, rs.a, rs.b, sN, u, v, tk == TK_CHECKER_UV || tk == TK_IMAGE);
		alive = scatter<float>(A.sc, pi.mat, pi.tex, r.d, sN, sP, u, v, rn, wo, att, emit);
		r.thr = r.thr * att;
	}
	r.o = sP;
	r.d = wo;
	++r.bounce;
}

}  // namespace areb
