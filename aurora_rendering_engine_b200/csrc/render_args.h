// render_args.h — the one argument of the render kernels (device-safe: also compiled by NVRTC for the baked kernel).
#pragma once
#include "dev_types.h"
#include "philox.cuh"

namespace areb {

struct RenderArgs {
	DevScene sc;
	CamBasis cam;
	CamF camf;       // cam rounded to fp32 (path integrator)
	RtCam rtcam;     // RT_AO integrator only
	PhiloxKey key;   // the ten round keys of the render seed (constant-bank operands in the kernel)
	int W, H;
	int s_begin, s_count;
	int max_depth;
	int ao_samples;
	float tmin;
	float bg_bottom[3], bg_top[3];
	int bg_black;                    // both background colours are zero: a miss adds nothing
	int lean;                        // use the lean brute-force kernel when the compiled scene has a lean form
	float *accum;                    // W*H*3 floats, sample SUMS are added
	unsigned long long *counters;    // [0] rays [1] node visits [2] quad tests [3] tri tests [4] sphere tests [5] box tests
};

enum { CNT_RAYS = 0, CNT_NODES = 1, CNT_QUADS = 2, CNT_TRIS = 3, CNT_SPHERES = 4, CNT_BOXES = 5, CNT_N = 8 };

}  // namespace areb
