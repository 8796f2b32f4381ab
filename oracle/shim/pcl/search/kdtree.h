// oracle/shim/pcl/search/kdtree.h — TEST INFRASTRUCTURE: exact brute-force stand-in for
// pcl::search::KdTree (FLANN).  Squared L2 distance accumulated (dx*dx + dy*dy) + dz*dz
// (FLANN L2_Simple); k-NN sorted by (distance, index); radius search returns the points with
// dist^2 < r^2 in ascending index order (PCL's unsorted order is FLANN's traversal order and
// cannot be pinned — see oracle/oracle.hpp "parity unpinned").
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl { namespace search {
template <typename P>
class KdTree {
public:
    typedef boost::shared_ptr<KdTree<P>> Ptr;
    typedef boost::shared_ptr<const KdTree<P>> ConstPtr;
    explicit KdTree(bool sorted = true) { (void)sorted; }
    template <typename CloudPtr>
    void setInputCloud(const CloudPtr& cloud, const IndicesPtr& indices = IndicesPtr()) {
        pts_ = &cloud->points;
        idx_.clear();
        if (indices) idx_ = *indices;
        else for (size_t i = 0; i < pts_->size(); ++i) idx_.push_back((int)i);
    }
    static float d2(const P& a, const P& b) {
        float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        return (dx * dx + dy * dy) + dz * dz;
    }
    int nearestKSearch(const P& q, int k, std::vector<int>& is, std::vector<float>& ds) const {
        if (k <= 2) {  // same (distance, index) order as below, without materialising all candidates
            float bd[2] = {0.f, 0.f};
            int bi[2] = {-1, -1};
            for (int i : idx_) {
                const P& p = (*pts_)[i];
                if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
                float d = d2(p, q);
                if (bi[0] < 0 || d < bd[0] || (d == bd[0] && i < bi[0])) {
                    bd[1] = bd[0]; bi[1] = bi[0];
                    bd[0] = d; bi[0] = i;
                } else if (bi[1] < 0 || d < bd[1] || (d == bd[1] && i < bi[1])) {
                    bd[1] = d; bi[1] = i;
                }
            }
            int kk = (bi[0] >= 0) + (bi[1] >= 0);
            kk = std::min(kk, k);
            is.resize(kk); ds.resize(kk);
            for (int j = 0; j < kk; ++j) { is[j] = bi[j]; ds[j] = bd[j]; }
            return kk;
        }
        std::vector<std::pair<float, int>> all;
        all.reserve(idx_.size());
        for (int i : idx_) {
            const P& p = (*pts_)[i];
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            all.push_back({d2(p, q), i});
        }
        int kk = std::min<int>(k, (int)all.size());
        std::partial_sort(all.begin(), all.begin() + kk, all.end());
        is.resize(kk); ds.resize(kk);
        for (int j = 0; j < kk; ++j) { is[j] = all[j].second; ds[j] = all[j].first; }
        return kk;
    }
    int radiusSearch(const P& q, double r, std::vector<int>& is, std::vector<float>& ds) const {
        is.clear(); ds.clear();
        float r2 = (float)r * (float)r;
        for (int i : idx_) {
            float d = d2((*pts_)[i], q);
            if (d < r2) { is.push_back(i); ds.push_back(d); }
        }
        return (int)is.size();
    }
private:
    const std::vector<P>* pts_ = nullptr;
    std::vector<int> idx_;
};
}}  // namespace pcl::search
