// <basic/math.h> — numeric constants of the Aurora Rendering Engine API (namespace are).
// Mirrors the one constant the reference exposes (reference include/basic/math.h:7) so host code written against the
// reference compiles unchanged against this tree.
#pragma once

namespace are {

/// Slack used by every geometric accept/reject decision of the fp64 library routines
/// (plane hit, triangle hit, barycentric containment, degenerate-plane check).
inline constexpr double GEOMETRY_EPSILON = 1e-12;

}  // namespace are
