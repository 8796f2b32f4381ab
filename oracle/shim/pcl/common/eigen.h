// oracle/shim/pcl/common/eigen.h — TEST INFRASTRUCTURE.  pcl::eigen33 feeds only the curvature
// RATIO test pc_min/pc_max < 0.2 (include/impl/model.hpp:98, scene.hpp:50).  PCL is absent from
// this image; the stand-in is the closed-form eigenvalue routine of pcl/common/impl/eigen.hpp
// restated in oracle/oracle.hpp (orc::pcl_eigen33), so the reference's principal_curvatures
// (include/impl/pointcloud.hpp:3-44) runs end to end when compiled against these headers.
// The seed-fixed synthetic configurations carry EXACT normals, for which the ratio is 0/0 (the NaN
// trap of SURVEY section 7) and supply their tangent masks themselves: for those runs
// (model::init, find_*) eigen33_mode() == 0 reports fixed eigenvalues whose ratio passes for every
// point, exactly as the mask input of the C-ABI does; ref_curvature switches to the real routine.
#pragma once
#include <Eigen/Dense>
namespace orc { void pcl_eigen33(const float cov[3][3], float evals[3]); }
namespace pcl {
inline int& eigen33_mode() { static int mode = 0; return mode; }
inline void eigen33(const Eigen::Matrix3f& m, Eigen::Vector3f& evals) {
    if (eigen33_mode() == 0) {
        evals = Eigen::Vector3f(0.f, 0.1f, 1.f);
        return;
    }
    float cov[3][3], ev[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[i][j] = m(i, j);
    orc::pcl_eigen33(cov, ev);
    evals = Eigen::Vector3f(ev[0], ev[1], ev[2]);
}
inline void computeCorrespondingEigenVector(const Eigen::Matrix3f&, float, Eigen::Vector3f& evec) { evec = Eigen::Vector3f(1.f, 0.f, 0.f); }
}  // namespace pcl
