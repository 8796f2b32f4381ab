// traits_common.hpp — helpers shared by the four *_traits headers (rigid frames, bulk GPU project)
#ifndef TRIPLET_MATCH_TRAITS_COMMON_HPP_
#define TRIPLET_MATCH_TRAITS_COMMON_HPP_

#include <tuple>

#include "common"
#include "feature"
#include "pointcloud"

namespace triplet_match {

// repeat_t<T, N> of include/cartesian_tuple:40-51: std::tuple<T, T, ...> (N times)
namespace detail {
template <typename T, std::size_t N, typename Idx = std::make_index_sequence<N>>
struct repeat;
template <typename T, std::size_t N, std::size_t... I>
struct repeat<T, N, std::index_sequence<I...>> {
    template <std::size_t>
    using dep = T;
    using type = std::tuple<dep<I>...>;
};

inline vec3f_t unit_orthogonal(const vec3f_t& v) {  // Eigen::unitOrthogonal for 3-vectors
    vec3f_t p;
    if (std::fabs(v[0]) > std::fabs(v[2]) * 1e-5f || std::fabs(v[1]) > std::fabs(v[2]) * 1e-5f) {
        const float inv = 1.f / std::sqrt(v[0] * v[0] + v[1] * v[1]);
        p = vec3f_t(-v[1] * inv, v[0] * inv, 0.f);
    } else {
        const float inv = 1.f / std::sqrt(v[1] * v[1] + v[2] * v[2]);
        p = vec3f_t(0.f, -v[2] * inv, v[1] * inv);
    }
    return p;
}
// g2l from three row vectors and an origin: rows r0,r1,r2; translation = R * (-origin)
inline mat4f_t frame_g2l(const vec3f_t& r0, const vec3f_t& r1, const vec3f_t& r2, const vec3f_t& origin) {
    mat4f_t g = mat4f_t::Identity();
    const vec3f_t rows[3] = {r0, r1, r2};
    const vec3f_t no(-origin[0], -origin[1], -origin[2]);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) g(i, j) = rows[i][j];
        g(i, 3) = rows[i].dot(no);
    }
    return g;
}
inline vec3f_t rotate(const mat4f_t& g, const vec3f_t& v) {  // topLeftCorner<3,3>() * v
    vec3f_t r;
    for (int i = 0; i < 3; ++i) r[i] = g(i, 0) * v[0] + (g(i, 1) * v[1] + g(i, 2) * v[2]);
    return r;
}
inline vec4f_t apply(const mat4f_t& g, const vec3f_t& p) { return g * vec4f_t(p[0], p[1], p[2], 1.f); }

// bulk project on the GPU (tm_traits_project): kind 0 cylinder, 1 plane, 2 plane2, 3 identity
inline void project_bulk(int kind, const mat4f_t& g2l, float radius, float threshold,
                         const std::vector<vec3f_t>& xyz, std::vector<vec3f_t>& uvw,
                         std::vector<uint8_t>& ok) {
    uvw.resize(xyz.size());
    ok.resize(xyz.size());
    check(tm_traits_project(default_ctx(), kind, g2l.data(), radius, threshold,
                            reinterpret_cast<const float*>(xyz.data()), xyz.size(),
                            reinterpret_cast<float*>(uvw.data()), ok.data()));
}
}  // namespace detail

template <typename T, std::size_t N>
using repeat_t = typename detail::repeat<T, N>::type;

}  // namespace triplet_match

#endif
