// oracle/shim/pcl/common/eigen.h — TEST INFRASTRUCTURE.  pcl::eigen33 feeds only the curvature
// RATIO test pc_min/pc_max < 0.2 (include/impl/model.hpp:98, scene.hpp:50); the synthetic
// configurations supply that mask themselves, so the stand-in reports fixed
// eigenvalues whose ratio passes the test for every point.
#pragma once
#include <Eigen/Dense>
namespace pcl {
inline void eigen33(const Eigen::Matrix3f&, Eigen::Vector3f& evals) { evals = Eigen::Vector3f(0.f, 0.1f, 1.f); }
inline void computeCorrespondingEigenVector(const Eigen::Matrix3f&, float, Eigen::Vector3f& evec) { evec = Eigen::Vector3f(1.f, 0.f, 0.f); }
}  // namespace pcl
