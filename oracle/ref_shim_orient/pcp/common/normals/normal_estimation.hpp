// TEST INFRASTRUCTURE.  Shadows the reference's pcp/common/normals/normal_estimation.hpp on the
// include path of oracle/ref_bridge_orient.cpp ONLY.  The real header defines
// pcp::estimate_normal on concrete Eigen 3.3.8 types (Eigen::Matrix3Xf, SelfAdjointEigenSolver),
// which this image does not have, and pcp/algorithm/estimate_normals.hpp includes it at the
// top.  The function the bridge needs from that file — propagate_normal_orientations
// (algorithm/estimate_normals.hpp:187-302) — never calls estimate_normal, so a declaration is
// enough to compile the rest of the reference unmodified.
#ifndef PCP_COMMON_NORMALS_NORMAL_ESTIMATION_HPP
#define PCP_COMMON_NORMALS_NORMAL_ESTIMATION_HPP
#include "pcp/common/normals/normal.hpp"
namespace pcp {
template <class ForwardIter, class PointViewMap, class Normal = pcp::normal_t>
Normal estimate_normal(ForwardIter it, ForwardIter end, PointViewMap const& point_map);
} // namespace pcp
#endif
