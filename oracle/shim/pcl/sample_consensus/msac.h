// oracle/shim — TEST INFRASTRUCTURE: stand-in so that the reference's impl/cylinder_traits.hpp compiles; the MSAC fit
// (init_from_model) is out of scope and never called.
#pragma once
#include <Eigen/Dense>
namespace pcl {
template <typename P> class MEstimatorSampleConsensus {
public:
    template <typename M> MEstimatorSampleConsensus(const M&, double) {}
    void setProbability(double) {}
    bool computeModel() { return false; }
    void getModelCoefficients(Eigen::VectorXf&) {}
};
}
