// oracle/shim — TEST INFRASTRUCTURE: stand-in so that the reference's impl/cylinder_traits.hpp compiles; the MSAC fit
// (init_from_model) is out of scope and never called.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename P, typename N> class SampleConsensusModelCylinder {
public:
    typedef boost::shared_ptr<SampleConsensusModelCylinder> Ptr;
    template <typename C> explicit SampleConsensusModelCylinder(const C&) {}
    void setIndices(const std::vector<int>&) {}
    template <typename C> void setInputNormals(const C&) {}
};
}
