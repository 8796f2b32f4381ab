// compat.hpp — layout-compatible stand-ins for the few Eigen / PCL types that appear in the
// reference's public API (include/common:31-55, pcl::PointSurfel, pcl::PointCloud).  Neither
// library exists in the build image, so the drop-in headers use these; the memory layouts are
// the ones the C-ABI views assume (column-major 4x4 floats, 48-byte PointSurfel).
#ifndef TRIPLET_MATCH_COMPAT_HPP_
#define TRIPLET_MATCH_COMPAT_HPP_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace Eigen {

template <typename T, int R, int C = 1>
struct Matrix {  // fixed size, column-major storage like Eigen's default
    T v[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    template <typename... A, typename = std::enable_if_t<sizeof...(A) == R * C && (R * C > 1)>>
    Matrix(A... a) : v{static_cast<T>(a)...} {}  // coefficient list (vectors: x, y, z[, w])
    T& operator()(int r, int c) { return v[c * R + r]; }
    const T& operator()(int r, int c) const { return v[c * R + r]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& coeffRef(int i) { return v[i]; }
    const T& coeff(int i) const { return v[i]; }
    T* data() { return v; }
    const T* data() const { return v; }
    static constexpr int size() { return R * C; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Constant(T c) { Matrix m; for (auto& x : m.v) x = c; return m; }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    bool operator==(const Matrix& o) const { for (int i = 0; i < R * C; ++i) if (!(v[i] == o.v[i])) return false; return true; }
    bool operator!=(const Matrix& o) const { return !(*this == o); }
    Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] - o.v[i]; return m; }
    Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] + o.v[i]; return m; }
    Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] * s; return m; }
    // 3-vector helpers in the Eigen evaluation order a0 + (a1 + a2)
    T dot(const Matrix& o) const { static_assert(R * C == 3 || R * C == 2 || R * C == 4, ""); if (R * C == 3) return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); T s = T(0); for (int i = 0; i < R * C; ++i) s += v[i] * o.v[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const { T z = squaredNorm(); if (z > T(0)) { T s = std::sqrt(z); Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] / s; return m; } return *this; }
    void normalize() { *this = normalized(); }
    Matrix cross(const Matrix& b) const { static_assert(R * C == 3, ""); Matrix m; m.v[0] = v[1] * b.v[2] - v[2] * b.v[1]; m.v[1] = v[2] * b.v[0] - v[0] * b.v[2]; m.v[2] = v[0] * b.v[1] - v[1] * b.v[0]; return m; }
    bool allFinite() const { for (auto x : v) if (!std::isfinite((double)x)) return false; return true; }
};

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<int, 4, 1> Vector4i;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;

// Matrix4f * Vector4f in the reference's packet order ((c0*x + c1*y) + c2*z) + c3*w
inline Vector4f operator*(const Matrix4f& m, const Vector4f& p) {
    Vector4f r;
    for (int i = 0; i < 4; ++i) r[i] = ((m(i, 0) * p[0] + m(i, 1) * p[1]) + m(i, 2) * p[2]) + m(i, 3) * p[3];
    return r;
}
inline Vector3f operator*(const Matrix3f& m, const Vector3f& p) {
    Vector3f r;
    for (int i = 0; i < 3; ++i) r[i] = m(i, 0) * p[0] + (m(i, 1) * p[1] + m(i, 2) * p[2]);
    return r;
}
inline Matrix4f operator*(const Matrix4f& a, const Matrix4f& b) {
    Matrix4f r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += (double)a(i, k) * (double)b(k, j);
            r(i, j) = (float)s;
        }
    return r;
}
// inverse of an affine transform [A t; 0 1] (all the reference ever inverts, scene.hpp:92)
inline Matrix4f affine_inverse(const Matrix4f& m) {
    double a[3][3], inv[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = m(i, j);
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    double id = 1.0 / det;
    inv[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * id; inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id; inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    inv[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id; inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id; inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    inv[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id; inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id; inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    Matrix4f r = Matrix4f::Identity();
    for (int i = 0; i < 3; ++i) {
        double t = 0;
        for (int j = 0; j < 3; ++j) { r(i, j) = (float)inv[i][j]; t -= inv[i][j] * (double)m(j, 3); }
        r(i, 3) = (float)t;
    }
    return r;
}

template <typename T, int N>
struct AlignedBox {
    Matrix<T, N, 1> mn, mx;
    AlignedBox() { setEmpty(); }
    void setEmpty() { for (int i = 0; i < N; ++i) { mn[i] = std::numeric_limits<T>::max(); mx[i] = std::numeric_limits<T>::lowest(); } }
    void extend(const Matrix<T, N, 1>& p) { for (int i = 0; i < N; ++i) { mn[i] = p[i] < mn[i] ? p[i] : mn[i]; mx[i] = p[i] > mx[i] ? p[i] : mx[i]; } }
    Matrix<T, N, 1>& min() { return mn; }
    Matrix<T, N, 1>& max() { return mx; }
    const Matrix<T, N, 1>& min() const { return mn; }
    const Matrix<T, N, 1>& max() const { return mx; }
    Matrix<T, N, 1> diagonal() const { return mx - mn; }
};

}  // namespace Eigen

namespace pcl {

// 48-byte surfel record with PCL's field layout: data[4] | data_n[4] | {rgba, radius, confidence, curvature}
struct alignas(16) PointSurfel {
    union { float data[4]; struct { float x, y, z; }; };
    union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
    union { struct { uint32_t rgba; float radius, confidence, curvature; }; float data_c[4]; };
    PointSurfel() { std::memset(static_cast<void*>(this), 0, sizeof(*this)); data[3] = 1.f; }
    Eigen::Vector3f getVector3f() const { return Eigen::Vector3f(x, y, z); }
    Eigen::Vector3f getNormalVector3f() const { return Eigen::Vector3f(normal_x, normal_y, normal_z); }
};
static_assert(sizeof(PointSurfel) == 48, "PointSurfel must be 48 bytes");

template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    virtual ~PointCloud() {}
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void push_back(const PointT& p) { points.push_back(p); }
    typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
    typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

}  // namespace pcl

#endif  // TRIPLET_MATCH_COMPAT_HPP_
