// <camera.h> — are::Camera (NEW; the reference library has no camera, its prototypes set rays up inline).
// The ray set-up is the one of experiments/rt.cpp:339-343,364-366 (pinhole, pixel centres), optionally with
// sub-pixel jitter and a thin lens; see are_camera in <are_cuda.h> for the exact formula.
#pragma once

#include <basic/vec3.h>

namespace are {

struct Camera {
	Point3 pos = Point3(0, 0, 1);
	Point3 target = Point3(0, 0, 0);
	Vec3 up = Vec3(0, 1, 0);
	double vfov_deg = 40.0;
	double focus_dist = 1.0;         // distance of the plane in perfect focus
	double defocus_angle_deg = 0.0;  // 0 = pinhole
	bool jitter = true;              // false = one ray through each pixel centre, as rt.cpp does

	Camera() = default;
	Camera(const Point3 &pos_, const Point3 &target_, const Vec3 &up_, double vfov_deg_) : pos(pos_), target(target_), up(up_), vfov_deg(vfov_deg_) {}
};

}  // namespace are
