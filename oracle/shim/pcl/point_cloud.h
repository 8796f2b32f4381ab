// oracle/shim/pcl/point_cloud.h — TEST INFRASTRUCTURE
#pragma once
#include <vector>
#include <boost/enable_shared_from_this.hpp>
namespace pcl {
typedef boost::shared_ptr<std::vector<int>> IndicesPtr;
template <typename P>
struct PointCloud {
    typedef boost::shared_ptr<PointCloud<P>> Ptr;
    typedef boost::shared_ptr<const PointCloud<P>> ConstPtr;
    std::vector<P> points;
    virtual ~PointCloud() {}
    size_t size() const { return points.size(); }
    void push_back(const P& p) { points.push_back(p); }
    typename std::vector<P>::const_iterator begin() const { return points.begin(); }
    typename std::vector<P>::const_iterator end() const { return points.end(); }
};
}  // namespace pcl
