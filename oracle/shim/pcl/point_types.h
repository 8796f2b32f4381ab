// oracle/shim/pcl/point_types.h — TEST INFRASTRUCTURE: 48-byte PointSurfel with PCL's field
// layout and Eigen::Map-like accessors (const maps copy; the mutable map writes through).
#pragma once
#include <cstdint>
#include <cstring>
// PCL's pcl_macros.h (1.8 / 1.9, the releases of the reference's time) does `#define _USE_MATH_DEFINES`
// + `#include <math.h>`; with libstdc++ that header pulls std::atan2 / std::fabs ... into the global
// namespace, which is what makes the reference's unqualified `atan2(loc[1], loc[0])` on floats
// (impl/cylinder_traits.hpp:109) resolve to the binary32 overload.  With <cmath> alone the call
// would bind ::atan2(double, double) and round once more (<= 1 ulp apart).
#include <math.h>
#include <Eigen/Dense>
namespace pcl {
struct Vector3fMapConst : Eigen::Vector3f {
    explicit Vector3fMapConst(const float* p) : Eigen::Vector3f(p[0], p[1], p[2]) {}
};
struct Vector3fMap : Eigen::Vector3f {
    float* ptr;
    explicit Vector3fMap(float* p) : Eigen::Vector3f(p[0], p[1], p[2]), ptr(p) {}
    Vector3fMap& operator=(const Eigen::Vector3f& o) {
        for (int i = 0; i < 3; ++i) { ptr[i] = o[i]; (*this)[i] = o[i]; }
        return *this;
    }
};
struct alignas(16) PointSurfel {
    union { float data[4]; struct { float x, y, z; }; };
    union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
    union { struct { uint32_t rgba; float radius, confidence, curvature; }; float data_c[4]; };
    PointSurfel() { std::memset(static_cast<void*>(this), 0, sizeof(*this)); data[3] = 1.f; }
    Vector3fMapConst getVector3fMap() const { return Vector3fMapConst(data); }
    Vector3fMap getVector3fMap() { return Vector3fMap(data); }
    Vector3fMapConst getNormalVector3fMap() const { return Vector3fMapConst(data_n); }
    Vector3fMap getNormalVector3fMap() { return Vector3fMap(data_n); }
};
static_assert(sizeof(PointSurfel) == 48, "PointSurfel is 48 bytes");
}  // namespace pcl
