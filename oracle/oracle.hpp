// oracle/oracle.hpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A dependency-free C++17 restatement of the CPU search path of
// richard-vock/triplet_match (feature/discretise -> hash probe -> RANSAC
// hypothesis generation + inlier scoring -> ICP correspondence).  Nothing in
// the product (triplet_match_b200/, include/) may include, link or call this
// file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / CPU baseline.
//
// PARITY STATUS: the reference ships no tests, golden vectors or fixtures, and
// cannot be built here as-is (PCL, FLANN, Eigen3, range-v3, boost, fmt are all
// absent; no network).  Pinning is therefore two-fold and stated per function:
//   * "ref-compiled": discretise/murmur/feature/valid/base_transform_/project_/
//     voxel_query are cross-checked against the reference's OWN sources compiled
//     where they lie against header shims (oracle/shim/, recipe oracle/Makefile
//     -> oracle/_ref/), see tests/test_oracle_vs_ref.py.  The shims fix the
//     Eigen evaluation order in writing (below); Eigen itself is unpinned.
//   * "parity unpinned": everything that depends on third-party arithmetic the
//     reference does not vendor: FLANN kd-tree visiting/tie-break order
//     (radius-search order, 1-NN ties), Eigen's SSE 4x4 inverse, Eigen::umeyama
//     + JacobiSVD, range-v3 sampling, libm atan2f.  The oracle fixes one
//     behaviour in writing for each (documented at the function).
//
// Float discipline (reference is built -O3, no -march, no fast-math =>
// SSE2 scalar/packet IEEE single, NO FMA; CMakeLists.txt:28-40): compile this
// file with -ffp-contract=off.  Eigen 3.3 fixed-size evaluation orders restated:
//   Matrix4f*Vector4f : ((c0*x + c1*y) + c2*z) + c3*w   (packet, sequential)
//   3-vector redux (dot, squaredNorm, Matrix3f row*vec) : a0 + (a1 + a2)
//   normalize()/normalized(): z=squaredNorm; if (z>0) v /= sqrt(z)  (true division)
//   cross(): (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
//   Matrix3f::inverse(): cofactor_3x3<i,j> * (1/det), det = col0 . cofactors_col0
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

// the one definition shared with the product (device + host): sinf / cosf on [0, pi/3] for pcl::eigen33
#include "../include/triplet_match/tm_sincosf.h"

namespace orc {

// ---------------------------------------------------------------- small math
struct v3 {
    float x, y, z;
};
struct m3 {
    float m[3][3];  // m[row][col]
};
struct m4 {
    float m[4][4];  // m[row][col]
};

inline v3 ld3(const float* p, uint32_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
inline v3 sub(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
inline float dot(v3 a, v3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float sqnorm(v3 a) { return dot(a, a); }
inline float norm(v3 a) { return sqrtf(sqnorm(a)); }
inline v3 cross(v3 a, v3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline v3 normalized(v3 a) {
    float z = sqnorm(a);
    if (z > 0.f) {
        float s = sqrtf(z);
        return {a.x / s, a.y / s, a.z / s};
    }
    return a;
}
inline v3 mul(const m3& A, v3 v) {
    return {sum3(A.m[0][0] * v.x, A.m[0][1] * v.y, A.m[0][2] * v.z),
            sum3(A.m[1][0] * v.x, A.m[1][1] * v.y, A.m[1][2] * v.z),
            sum3(A.m[2][0] * v.x, A.m[2][1] * v.y, A.m[2][2] * v.z)};
}
inline m3 mul(const m3& A, const m3& B) {
    m3 C;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            C.m[r][c] = sum3(A.m[r][0] * B.m[0][c], A.m[r][1] * B.m[1][c], A.m[r][2] * B.m[2][c]);
    return C;
}
// Matrix4f * (x,y,z,w): per row ((m0*x + m1*y) + m2*z) + m3*w
inline void mul4(const m4& A, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r)
        out[r] = ((A.m[r][0] * v[0] + A.m[r][1] * v[1]) + A.m[r][2] * v[2]) + A.m[r][3] * v[3];
}
template <int I, int J>
inline float cofactor3(const m3& A) {
    constexpr int i1 = (I + 1) % 3, i2 = (I + 2) % 3, j1 = (J + 1) % 3, j2 = (J + 2) % 3;
    return A.m[i1][j1] * A.m[i2][j2] - A.m[i1][j2] * A.m[i2][j1];
}
inline m3 inverse3(const m3& A) {
    float c0 = cofactor3<0, 0>(A), c1 = cofactor3<1, 0>(A), c2 = cofactor3<2, 0>(A);
    float det = sum3(c0 * A.m[0][0], c1 * A.m[1][0], c2 * A.m[2][0]);
    float invdet = 1.f / det;
    m3 R;
    R.m[0][0] = c0 * invdet;
    R.m[0][1] = c1 * invdet;
    R.m[0][2] = c2 * invdet;
    R.m[1][0] = cofactor3<0, 1>(A) * invdet;
    R.m[1][1] = cofactor3<1, 1>(A) * invdet;
    R.m[1][2] = cofactor3<2, 1>(A) * invdet;
    R.m[2][0] = cofactor3<0, 2>(A) * invdet;
    R.m[2][1] = cofactor3<1, 2>(A) * invdet;
    R.m[2][2] = cofactor3<2, 2>(A) * invdet;
    return R;
}

// ------------------------------------------------------------------- atan2f
// libm's atan2f (include/impl/feature.hpp:7, include/impl/cylinder_traits.hpp:109) is
// third-party arithmetic.  Restated here: the algorithm glibc 2.39 ships on the reference's
// platform (x86-64; sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c = fdlibm's single-precision
// atan2f/atanf, built without FMA), i.e. plain binary32 +,-,*,/.  PINNED: bit-identical to this
// image's libm on 10^8 inputs in all quadrants incl. special values (tests/test_oracle_math.py
// repeats the check on whatever host the tests run on and reports if that host's libm differs).
inline float f32_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t bits_f32(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float atanf_glibc(float x) {
    static const uint32_t HI[4] = {0x3eed6338u, 0x3f490fdau, 0x3f7b985eu, 0x3fc90fdau};  // atan(.5,1,1.5,inf) hi
    static const uint32_t LO[4] = {0x31ac3769u, 0x33222168u, 0x33140fb4u, 0x33a22168u};  // ... lo
    static const uint32_t AT[11] = {0x3eaaaaabu, 0xbe4ccccdu, 0x3e124925u, 0xbde38e38u, 0x3dba2e6eu, 0xbd9d8795u,
                                    0x3d886b35u, 0xbd6ef16bu, 0x3d4bda59u, 0xbd15a221u, 0x3c8569d7u};
    float aT[11];
    for (int i = 0; i < 11; ++i) aT[i] = f32_bits(AT[i]);
    uint32_t hx = bits_f32(x), ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u) {  // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;
        float r = f32_bits(HI[3]) + f32_bits(LO[3]);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {  // |x| < 0.4375
        if (ix < 0x31000000u) return x;
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3f980000u) {
            if (ix < 0x3f300000u) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    float z = x * x;
    float w = z * z;
    float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = f32_bits(HI[id]) - ((x * (s1 + s2) - f32_bits(LO[id])) - x);
    return (hx >> 31) ? -z : z;
}
inline float atan2f_full(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = f32_bits(0x3f490fdbu), pi_o_2 = f32_bits(0x3fc90fdbu),
                pi = f32_bits(0x40490fdbu), pi_lo = f32_bits(0xb3bbbd2eu);
    uint32_t hx = bits_f32(x), hy = bits_f32(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
    if (hx == 0x3f800000u) return atanf_glibc(y);
    uint32_t m = (hy >> 31) | ((hx >> 30) & 2u);
    if (iy == 0) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    int32_t k = ((int32_t)iy - (int32_t)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = atanf_glibc(std::fabs(y / x));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}
inline float atan2f_q1(float y, float x) { return atan2f_full(y, x); }  // the call site has y, x >= 0

// ------------------------------------------------- discretise + murmur (a3,a4)
// src/discretize.cpp:19-25
inline uint32_t discretize(float value, float min_value, float range_value, uint32_t steps) {
    float nval = (value - min_value) / range_value;
    if (nval < 0.f) return 0;
    if (nval >= 1.f) return steps - 1;
    return static_cast<uint32_t>(nval * steps);
}
// src/discretize.cpp:27-30
inline uint32_t discretize(float value, float step_size) {
    return static_cast<uint32_t>(value / step_size);
}
inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
// include/impl/discretize.hpp:10-45 (MurmurHash3_x86_32 body, seed 42, len 4*Dim)
inline uint32_t murmur4(const uint32_t key[4]) {
    uint32_t h1 = 42u;
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    for (int i = 0; i < 4; ++i) {
        uint32_t k1 = key[i];
        k1 *= c1;
        k1 = rotl32(k1, 15);
        k1 *= c2;
        h1 ^= k1;
        h1 = rotl32(h1, 13);
        h1 = h1 * 5u + 0xe6546b64u;
    }
    h1 ^= 16u;
    h1 ^= h1 >> 16;
    h1 *= 0x85ebca6bu;
    h1 ^= h1 >> 13;
    h1 *= 0xc2b2ae35u;
    h1 ^= h1 >> 16;
    return h1;
}

struct discretization_params {  // include/discretize:8-12
    float distance_step_count;
    float angle_step;
    float curvature_ratio_step_count;
};
struct feature_bounds {  // Eigen::AlignedBox<float,4>
    float mn[4], mx[4];
    void set_empty() {
        for (int i = 0; i < 4; ++i) {
            mn[i] = std::numeric_limits<float>::max();
            mx[i] = std::numeric_limits<float>::lowest();
        }
    }
    void extend(const float f[4]) {
        for (int i = 0; i < 4; ++i) {
            mn[i] = std::min(mn[i], f[i]);
            mx[i] = std::max(mx[i], f[i]);
        }
    }
};

// ------------------------------------------------------------- feature (a1,a2)
// include/impl/feature.hpp:5-8
inline float angle(v3 a, v3 b, bool use_libm = false) {
    float y = norm(cross(a, b));
    float x = fabsf(dot(a, b));
    return use_libm ? ::atan2f(y, x) : atan2f_q1(y, x);
}
// include/impl/feature.hpp:15-34
inline void feature(v3 p0, v3 t0, v3 p1, v3 t1, float f[4], bool use_libm = false) {
    v3 d0 = sub(p1, p0);
    f[0] = norm(d0);
    f[1] = angle(d0, t0, use_libm);
    f[2] = angle(d0, t1, use_libm);
    f[3] = f[0];
}
// include/impl/feature.hpp:48-88
inline bool valid(const float f[4], const feature_bounds& b) {
    if (f[0] < b.mn[0] || f[0] > b.mx[0]) return false;
    float pi = static_cast<float>(M_PI);
    return (f[1] >= 0.f && f[1] <= pi) && (f[2] >= 0.f && f[2] <= pi);
}
// include/impl/feature.hpp:36-46
inline void discretize_feature(const float f[4], const feature_bounds& b,
                               const discretization_params& p, uint32_t key[4]) {
    uint32_t steps = static_cast<uint32_t>(p.distance_step_count);
    float diag0 = b.mx[0] - b.mn[0];
    key[0] = discretize(f[0], b.mn[0], diag0, steps);
    key[1] = discretize(f[1], p.angle_step);
    key[2] = discretize(f[2], p.angle_step);
    key[3] = discretize(f[3], b.mn[0], diag0, steps);
}
// include/impl/feature.hpp:90-114
inline feature_bounds valid_bounds(const feature_bounds& b, float min_rel, float max_rel) {
    feature_bounds nb = b;
    nb.mn[0] = b.mn[0] + min_rel * (b.mx[0] - b.mn[0]);
    nb.mx[0] = b.mn[0] + max_rel * (b.mx[0] - b.mn[0]);
    nb.mn[3] = b.mn[3] + min_rel * (b.mx[3] - b.mn[3]);
    nb.mx[3] = b.mn[3] + max_rel * (b.mx[3] - b.mn[3]);
    return nb;
}

// -------------------------------------------------------------------- clouds
struct cloud {  // packed n x 3 arrays; PointSurfel fields pos / normal / tangent (common:62-70)
    const float* pos;
    const float* nrm;
    const float* tgt;
    uint32_t n;
};

// brute-force k=2 nearest: FLANN L2_Simple accumulates (dx*dx + dy*dy) + dz*dz
inline float sqdist_seq(v3 a, v3 b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}
// include/impl/pointcloud.hpp:66-82 — mean 1-NN distance with the running
// average of include/common:104-115.  [parity unpinned: FLANN]
inline float resolution(const cloud& c) {
    float accum = 0.f;
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < c.n; ++i) {
        v3 p = ld3(c.pos, i);
        float best = std::numeric_limits<float>::max();
        for (uint32_t j = 0; j < c.n; ++j) {
            if (j == i) continue;
            float d = sqdist_seq(ld3(c.pos, j), p);
            if (d < best) best = d;
        }
        float val = sqrtf(best);
        accum = accum + (val - accum) / (++cnt);
    }
    return accum;
}

// --------------------------------------------------------------------- model
struct key4 {
    uint32_t k[4];
    bool operator==(const key4& o) const {
        return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2] && k[3] == o.k[3];
    }
};
struct key4_hash {
    std::size_t operator()(const key4& k) const { return murmur4(k.k); }
};
typedef std::unordered_multimap<key4, std::pair<uint32_t, uint32_t>, key4_hash> hash_map_t;

struct sample_parameters {  // include/common:72-82 (only the read fields)
    float min_diameter_factor;
    float max_diameter_factor;
    bool force_up;
};

// include/impl/model.hpp:63 — to_voxel_.inverse() for to_voxel_ = diag(scale) + trans, as Eigen's SSE
// Matrix4f::inverse() (2x2-block cofactor routine, Intel AP-928) leaves it once the products with the
// structural zeros are dropped; see model::init below.  ia = inverse diagonal, ib = inverse translation.
inline void voxel_centre_map(const float scale[3], const float trans[3], float ia[3], float ib[3]) {
    const float sxy = scale[0] * scale[1], rd = 1.0f / (sxy * scale[2]);
    ia[0] = rd * (scale[1] * scale[2]);
    ia[1] = rd * (scale[0] * scale[2]);
    ia[2] = rd * sxy;
    ib[0] = (-rd) * (0.0f - (0.0f - scale[1] * (scale[2] * trans[0])));
    ib[1] = rd * (0.0f - scale[0] * (scale[2] * trans[1]));
    ib[2] = (-rd) * (trans[2] * sxy);
}

struct model {
    cloud c;
    discretization_params dp;
    float resolution_ = 0.f;
    float diameter = 0.f;
    int extents[3] = {0, 0, 0};
    int margin = 5;
    float scale[3] = {1, 1, 1};   // to_voxel_ diagonal
    float trans[3] = {0, 0, 0};   // to_voxel_ translation
    std::vector<uint32_t> voxel;  // voxel_data_[k*ex*ey + j*ex + i]
    std::vector<uint32_t> subset; // tangent subset (point_count())
    feature_bounds fb;
    hash_map_t map;
    bool init_ = false;

    m4 to_voxel() const {
        m4 t{};
        for (int i = 0; i < 3; ++i) {
            t.m[i][i] = scale[i];
            t.m[i][3] = trans[i];
        }
        t.m[3][3] = 1.f;
        return t;
    }

    // include/impl/model.hpp:16-167.  curv_ok[i] != 0 replaces the PCL
    // curvature criterion pc_min/pc_max < 0.2 (:98) — the synthetic generators
    // supply the mask; nullptr = all true.  resolution < 0 =>
    // computed by brute force.
    // model.hpp:87-88: nearest model point of the centre of cell (i, j, k)
    uint32_t cell_nearest(int i, int j, int k) const {
        float ia[3], ib[3];
        voxel_centre_map(scale, trans, ia, ib);
        v3 q = {ia[0] * (float)i + ib[0], ia[1] * (float)j + ib[1], ia[2] * (float)k + ib[2]};
        float best = std::numeric_limits<float>::max();
        uint32_t bi = 0;
        for (uint32_t p = 0; p < c.n; ++p) {
            float d = sqdist_seq(ld3(c.pos, p), q);
            if (d < best) {
                best = d;
                bi = p;
            }
        }
        return bi;
    }

    // voxel_in != nullptr: the grid is supplied instead of filled (checks at sizes where the brute-force
    // fill below, O(cells x points), does not finish: the supplied grid is then spot-checked cell by cell
    // with cell_nearest()).
    void init(const cloud& cl, const discretization_params& params, const sample_parameters& sp,
              const uint8_t* curv_ok, float given_resolution, const uint32_t* voxel_in = nullptr,
              const uint8_t* in_subset = nullptr) {
        c = cl;
        dp = params;
        std::vector<uint32_t> all;
        for (uint32_t i = 0; i < c.n; ++i) {  // :17-22 subset_ (every point when the caller's is empty), :24-30 allFinite
            if (in_subset && !in_subset[i]) continue;
            bool fin = true;
            for (int k = 0; k < 3; ++k)
                fin = fin && std::isfinite(c.pos[3 * i + k]) && std::isfinite(c.nrm[3 * i + k]) &&
                      std::isfinite(c.tgt[3 * i + k]);
            if (fin) all.push_back(i);
        }
        float lo[3], hi[3];  // :34-39 bbox
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::numeric_limits<float>::max();
            hi[k] = std::numeric_limits<float>::lowest();
        }
        for (uint32_t i : all)
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], c.pos[3 * i + k]);
                hi[k] = std::max(hi[k], c.pos[3 * i + k]);
            }
        v3 range = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        diameter = norm(range);
        resolution_ = given_resolution > 0.f ? given_resolution : resolution(c);
        float half_res = 0.5f * resolution_;  // :45-46
        float ext[3] = {std::max(range.x / half_res, 1.f), std::max(range.y / half_res, 1.f),
                        std::max(range.z / half_res, 1.f)};
        float rg[3] = {range.x, range.y, range.z};
        margin = 5;
        for (int k = 0; k < 3; ++k) {
            extents[k] = static_cast<int>(ext[k] + 2.f * margin);          // :50
            scale[k] = rg[k] < 1e-5f ? 1.f : ext[k] / rg[k];               // :52-55
            // :58-61  diag(scale)*(-min) via Matrix3f*Vector3f redux s*(-m) + (0 + 0)
            trans[k] = (scale[k] * (-lo[k]) + static_cast<float>(margin)) - 0.5f;
        }
        // :63, :81-94 grid fill.  World centre of voxel (i,j,k) = (to_voxel_.inverse() * (i,j,k,1)).head(3).
        // Matrix4f::inverse() is Eigen's SSE 2x2-block cofactor routine (Intel AP-928); for to_voxel_ =
        // diag(scale) + trans all its products with structural zeros vanish and it leaves
        //   rd = 1 / ((sx*sy)*sz),  inv_aa = rd * (product of the other two scales),
        //   inv_03 = -rd*(sy*(sz*tx)), inv_13 = -rd*(sx*(sz*ty)), inv_23 = -rd*(tz*(sx*sy))
        // (checked against the lane-by-lane restatement in oracle/shim/Eigen/inverse_size4_sse.h), and
        // inv * (i,j,k,1) in Eigen's packet order is inv_aa * index + inv_a3 per axis.
        // 1-NN [parity unpinned: FLANN] is exact brute force, squared distance (dx*dx + dy*dy) + dz*dz,
        // lowest index wins ties.
        voxel.assign((size_t)extents[0] * extents[1] * extents[2], 0u);
        if (voxel_in) {
            std::copy(voxel_in, voxel_in + voxel.size(), voxel.begin());
        } else {
            for (int k = 0; k < extents[2]; ++k)
                for (int j = 0; j < extents[1]; ++j)
                    for (int i = 0; i < extents[0]; ++i)
                        voxel[(size_t)k * extents[0] * extents[1] + (size_t)j * extents[0] + i] = cell_nearest(i, j, k);
        }
        // :96-99 tangent subset
        subset.clear();
        for (uint32_t i : all)
            if (norm(ld3(c.tgt, i)) > 0.7f && (!curv_ok || curv_ok[i])) subset.push_back(i);
        float lower_bound = diameter * sp.min_diameter_factor;  // :101-102
        float upper_bound = diameter * sp.max_diameter_factor;
        fb.set_empty();
        auto pair_ok = [&](uint32_t i, uint32_t j) {  // :105-111
            if (i == j) return false;
            v3 d1 = sub(ld3(c.pos, j), ld3(c.pos, i));
            float dist1 = norm(d1);
            d1 = {d1.x / dist1, d1.y / dist1, d1.z / dist1};
            if (dist1 < lower_bound || dist1 > upper_bound) return false;
            if (1.f - fabsf(dot(d1, ld3(c.tgt, i))) < 0.01f) return false;
            return true;
        };
        for (uint32_t i : subset)  // :104-118 (cartesian_product: first range is outer)
            for (uint32_t j : subset) {
                if (!pair_ok(i, j)) continue;
                float f[4];
                feature(ld3(c.pos, i), ld3(c.tgt, i), ld3(c.pos, j), ld3(c.tgt, j), f);
                fb.extend(f);
            }
        fb = valid_bounds(fb, 0.0f, 1.f);  // :122
        map.clear();
        for (uint32_t i : subset)  // :125-149
            for (uint32_t j : subset) {
                if (!pair_ok(i, j)) continue;
                float f[4];
                feature(ld3(c.pos, i), ld3(c.tgt, i), ld3(c.pos, j), ld3(c.tgt, j), f);
                if (valid(f, fb)) {
                    key4 k;
                    discretize_feature(f, fb, dp, k.k);
                    map.insert({k, {i, j}});
                }
            }
        init_ = true;
    }

    // include/impl/model.hpp:180-192
    bool voxel_query(const float pos[4], uint32_t& out) const {
        float v[4];
        mul4(to_voxel(), pos, v);
        int i = static_cast<int>(v[0]), j = static_cast<int>(v[1]), k = static_cast<int>(v[2]);
        if (i < 0 || j < 0 || k < 0 || i >= extents[0] || j >= extents[1] || k >= extents[2])
            return false;
        int lin = k * extents[0] * extents[1] + j * extents[0] + i;
        out = voxel[lin];
        return true;
    }
};

// --------------------------------------------------- base_transform_ (a6)
// include/impl/scene.hpp:538-567 (scale_invariant == false)
inline m4 base_transform(v3 src_i, v3 src_j, v3 src_t, v3 tgt_i, v3 tgt_j, v3 tgt_t) {
    v3 o_a = src_i, o_b = tgt_i;
    v3 u_a = normalized(sub(src_j, o_a));
    v3 u_b = normalized(sub(tgt_j, o_b));
    v3 v_a = src_t, v_b = tgt_t;
    float da = dot(v_a, u_a), db = dot(v_b, u_b);
    v_a = {v_a.x - da * u_a.x, v_a.y - da * u_a.y, v_a.z - da * u_a.z};
    v_b = {v_b.x - db * u_b.x, v_b.y - db * u_b.y, v_b.z - db * u_b.z};
    v_a = normalized(v_a);
    v_b = normalized(v_b);
    v3 w_a = normalized(cross(u_a, v_a));
    v3 w_b = normalized(cross(u_b, v_b));
    m3 A = {{{u_a.x, v_a.x, w_a.x}, {u_a.y, v_a.y, w_a.y}, {u_a.z, v_a.z, w_a.z}}};
    m3 B = {{{u_b.x, v_b.x, w_b.x}, {u_b.y, v_b.y, w_b.y}, {u_b.z, v_b.z, w_b.z}}};
    m3 R = mul(B, inverse3(A));
    v3 Ro = mul(R, o_a);
    m4 t{};
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) t.m[r][cc] = R.m[r][cc];
    t.m[0][3] = o_b.x - Ro.x;
    t.m[1][3] = o_b.y - Ro.y;
    t.m[2][3] = o_b.z - Ro.z;
    t.m[3][3] = 1.f;
    return t;
}

// --------------------------------------------------------------------- scene
struct scene {
    cloud c;
    std::vector<int> mask;          // mask_
    std::vector<int> tangent_mask;  // tangent_mask_
};

struct project_result {
    std::vector<uint32_t> scene_corrs, model_corrs;
    double score = 0.0;
    uint32_t saved = 0;
    bool dropped = false;
};

// The reference's `uint32_t upper = -1.0 - static_cast<uint32_t>((x*n + tmp) / N)`
// (include/impl/scene.hpp:498) casts negative doubles to uint32_t (UB).  On
// x86-64/gcc it evaluates via cvttsd2si(64-bit)+truncation; restated here with
// defined arithmetic: q = trunc(v) as int64, a = (uint32)q, b = -1.0 - (double)a,
// upper = (uint32)(int64)trunc(b).
inline uint32_t early_drop_upper(uint32_t tried, uint32_t subset_size, uint32_t corrs) {
    double N = -2.0 - tried;
    double x = -2.0 - subset_size;
    double n = -1.0 - corrs;
    double tmp = std::sqrt((x * n * (N - x) * (N - n)) / (N - 1.0));
    double v = (x * n + tmp) / N;
    uint32_t a = static_cast<uint32_t>(static_cast<uint64_t>(static_cast<int64_t>(v)));
    double b = -1.0 - static_cast<double>(a);
    return static_cast<uint32_t>(static_cast<uint64_t>(static_cast<int64_t>(b)));
}

// checkpoints of project_ (:422-426): tests[i] = step_size*(i+1)*|subset|
inline std::vector<uint32_t> early_drop_tests(size_t subset_size) {
    constexpr float step_size = 0.05f;
    std::vector<uint32_t> tests(static_cast<uint32_t>(1.f / step_size) - 2);
    for (uint32_t i = 0; i < tests.size(); ++i) tests[i] = step_size * (i + 1) * subset_size;
    return tests;
}

// include/impl/scene.hpp:411-510.  `subset` order is the caller's; the
// reference's order is FLANN's radiusSearch order [parity unpinned] — the
// recorded configurations use ascending scene index.
inline project_result project(const scene& s, const model& m, const int* subset, size_t nsub,
                              const m4& t, float accept_prob, float dist_thres, bool early_out) {
    project_result r;
    float thres = dist_thres * m.resolution_;  // :413
    m3 t_tgt;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) t_tgt.m[a][b] = t.m[a][b];
    uint32_t tried = 0;
    size_t next_test = 0;
    std::vector<uint32_t> tests = early_drop_tests(nsub);
    double signed_score = 0;
    for (size_t q = 0; q < nsub; ++q) {
        int idx = subset[q];
        ++tried;
        bool skip = s.mask[idx] != 0;  // :434 (`continue` also skips the checkpoint below)
        if (skip) continue;
        v3 p = ld3(s.c.pos, idx);
        float pos[4] = {p.x, p.y, p.z, 1.f}, tp[4];
        bool use_tangent = s.tangent_mask[idx] != 0;
        v3 ref = use_tangent ? ld3(s.c.tgt, idx) : ld3(s.c.nrm, idx);
        mul4(t, pos, tp);  // :444
        uint32_t pn;
        if (!m.voxel_query(tp, pn)) continue;  // :455-460
        ref = mul(t_tgt, ref);                 // :461
        v3 mp = ld3(m.c.pos, pn);
        float dist = norm(sub(v3{tp[0], tp[1], tp[2]}, mp));  // :464
        if (!(dist > thres)) {
            v3 ref_n = ld3(m.c.tgt, pn);
            bool is_tangent = norm(ref_n) > 0.7f;  // :470
            if (use_tangent == is_tangent) {
                if (!is_tangent) ref_n = ld3(m.c.nrm, pn);
                signed_score += 1.0 * static_cast<double>(fabsf(dot(ref, ref_n)));  // :483
                r.scene_corrs.push_back(idx);
                r.model_corrs.push_back(pn);
            }
        }
        if (early_out && next_test < tests.size() && tried >= tests[next_test]) {  // :492
            uint32_t upper = early_drop_upper(tried, (uint32_t)nsub, (uint32_t)r.model_corrs.size());
            if (upper < accept_prob * m.c.n) {  // :500 (uint32 -> float compare)
                r.saved = (uint32_t)nsub - tried;
                r.score = signed_score;  // un-normalised on drop (:502)
                r.dropped = true;
                return r;
            }
            ++next_test;
        }
    }
    r.score = signed_score / m.c.n;  // :509 / :406-409
    return r;
}

// include/impl/scene.hpp:273 -> pointcloud.hpp:169-177: indices with
// ||p - p1||^2 < r^2 (FLANN radiusSearch takes r^2 from PCL; strict '<' is not
// pinned by the reference [parity unpinned: FLANN]); ascending index order.
inline std::vector<int> ball_subset(const cloud& c, v3 p1, float radius) {
    std::vector<int> out;
    float r2 = radius * radius;
    for (uint32_t i = 0; i < c.n; ++i)
        if (sqdist_seq(ld3(c.pos, i), p1) < r2) out.push_back((int)i);
    return out;
}

// pair filter of find_in_subset (include/impl/scene.hpp:290-302); true => f valid
inline bool scene_pair_feature(const scene& s, const model& m, uint32_t i, uint32_t j, float lower,
                               float upper, float f[4]) {
    if (!s.tangent_mask[j] || s.mask[j] || i == j) return false;  // :290
    v3 p1 = ld3(s.c.pos, i), p2 = ld3(s.c.pos, j);
    v3 d0 = sub(p2, p1);
    float sqn0 = sqnorm(d0);
    d0 = normalized(d0);
    if (sqn0 < lower || sqn0 > upper) return false;                       // :296
    if (1.f - fabsf(dot(d0, ld3(s.c.tgt, i))) < 0.01f) return false;      // :297
    feature(p1, ld3(s.c.tgt, i), p2, ld3(s.c.tgt, j), f);                 // :299
    return valid(f, m.fb);                                                // :300
}

// ---------------------------------------------------------------- umeyama/icp
// 3x3 SVD by one-sided Jacobi in double.  Eigen::umeyama + JacobiSVD<float>
// (include/impl/scene.hpp:393) are third-party and unpinned [parity unpinned];
// north_star's pose tolerance (1e-4) is the contract here.
inline void svd3(const double A[3][3], double U[3][3], double S[3], double V[3][3]) {
    double B[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            B[i][j] = A[i][j];
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < 3; ++i) {
                    a += B[i][p] * B[i][p];
                    b += B[i][q] * B[i][q];
                    g += B[i][p] * B[i][q];
                }
                off = std::max(off, std::fabs(g) / (std::sqrt(a * b) + 1e-300));
                if (std::fabs(g) < 1e-300) continue;
                double zeta = (b - a) / (2.0 * g);
                double tt = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / std::sqrt(1.0 + tt * tt), sn = cs * tt;
                for (int i = 0; i < 3; ++i) {
                    double bp = B[i][p], bq = B[i][q];
                    B[i][p] = cs * bp - sn * bq;
                    B[i][q] = sn * bp + cs * bq;
                    double vp = V[i][p], vq = V[i][q];
                    V[i][p] = cs * vp - sn * vq;
                    V[i][q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < 3; ++j) {
        S[j] = std::sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
    }
    // sort descending
    int order[3] = {0, 1, 2};
    std::sort(order, order + 3, [&](int a, int b) { return S[a] > S[b]; });
    double Bs[3][3], Vs[3][3], Ss[3];
    for (int j = 0; j < 3; ++j) {
        Ss[j] = S[order[j]];
        for (int i = 0; i < 3; ++i) {
            Bs[i][j] = B[i][order[j]];
            Vs[i][j] = V[i][order[j]];
        }
    }
    for (int j = 0; j < 3; ++j) {
        S[j] = Ss[j];
        for (int i = 0; i < 3; ++i) {
            V[i][j] = Vs[i][j];
            U[i][j] = Ss[j] > 1e-300 ? Bs[i][j] / Ss[j] : 0.0;
        }
    }
    // complete U for rank-deficient input: make column 2 = col0 x col1
    if (S[2] <= 1e-12 * S[0]) {
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
}
inline double det3(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
           M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}
// Eigen::umeyama(src, dst, with_scaling=false): dst ~ R*src + t
inline m4 umeyama(const std::vector<v3>& src, const std::vector<v3>& dst) {
    size_t n = src.size();
    double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i) {
        ms[0] += src[i].x; ms[1] += src[i].y; ms[2] += src[i].z;
        md[0] += dst[i].x; md[1] += dst[i].y; md[2] += dst[i].z;
    }
    for (int k = 0; k < 3; ++k) {
        ms[k] /= (double)n;
        md[k] /= (double)n;
    }
    double sigma[3][3] = {{0}};
    for (size_t i = 0; i < n; ++i) {
        double s[3] = {src[i].x - ms[0], src[i].y - ms[1], src[i].z - ms[2]};
        double d[3] = {dst[i].x - md[0], dst[i].y - md[1], dst[i].z - md[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) sigma[a][b] += d[a] * s[b];
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) sigma[a][b] /= (double)n;
    double U[3][3], S[3], V[3][3];
    svd3(sigma, U, S, V);
    double sgn = det3(U) * det3(V) < 0 ? -1.0 : 1.0;
    double R[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            R[a][b] = U[a][0] * V[b][0] + U[a][1] * V[b][1] + sgn * U[a][2] * V[b][2];
    m4 t{};
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) t.m[a][b] = (float)R[a][b];
        t.m[a][3] = (float)(md[a] - (R[a][0] * ms[0] + R[a][1] * ms[1] + R[a][2] * ms[2]));
    }
    t.m[3][3] = 1.f;
    return t;
}

struct match {
    m4 transform;
    std::vector<uint32_t> scene_corrs, model_corrs;
    double score;
};

// include/impl/scene.hpp:100-106
inline match finish_find(const scene& s, const model& m, const m4& t, float accept_prob,
                         float dist_thres) {
    std::vector<int> all(s.c.n);
    for (uint32_t i = 0; i < s.c.n; ++i) all[i] = (int)i;
    project_result r = project(s, m, all.data(), all.size(), t, accept_prob, dist_thres, false);
    return {t, r.scene_corrs, r.model_corrs, r.score};
}

// include/impl/scene.hpp:369-404
inline match icp(const scene& s, const model& m, const match& start, uint32_t max_iterations,
                 float dist_thres, float accept_prob, uint32_t* iters_out = nullptr) {
    if (iters_out) *iters_out = 0;
    if (max_iterations == 0) return start;
    match best = finish_find(s, m, start.transform, accept_prob, 2 * dist_thres);
    uint32_t iter = 0;
    while (true) {
        if (best.scene_corrs.size() < 3) return best;
        std::vector<v3> mdl(best.model_corrs.size()), scn(best.scene_corrs.size());
        for (size_t i = 0; i < best.scene_corrs.size(); ++i) {
            mdl[i] = ld3(m.c.pos, best.model_corrs[i]);
            scn[i] = ld3(s.c.pos, best.scene_corrs[i]);
        }
        m4 trans = umeyama(scn, mdl);
        match proj = finish_find(s, m, trans, accept_prob, 2 * dist_thres);
        if (proj.scene_corrs.size() < best.scene_corrs.size()) return best;
        best = proj;
        ++iter;
        if (iters_out) *iters_out = iter;
        if (iter == max_iterations) return best;
    }
}

// --------------------------------------------------------------------- traits
// per-point closed forms (a14).  g2l is row-major 4x4.
inline void g2l_apply(const m4& g2l, v3 p, float out[4]) {
    float v[4] = {p.x, p.y, p.z, 1.f};
    mul4(g2l, v, out);
}
// include/impl/cylinder_traits.hpp:102-114
inline bool cylinder_project(const m4& g2l, float radius, float threshold, v3 xyz, float uvw[3]) {
    float loc[4];
    g2l_apply(g2l, xyz, loc);
    float height = sqrtf(loc[0] * loc[0] + loc[1] * loc[1]) - radius;  // head(2).norm(): a0+a1
    if (fabsf(height) > threshold) return false;
    uvw[1] = loc[2];
    uvw[2] = height / radius;
    // atan2(loc[1], loc[0]) on floats: the binary32 overload (PCL's <math.h> puts std::atan2 into the
    // global namespace) = glibc atan2f.  PINNED against the reference's own header, oracle/_ref
    float ang = atan2f_full(loc[1], loc[0]);
    if (ang < 0.f) ang = (float)((double)ang + 2.0 * M_PI);  // :111 `+= 2.0*M_PI` in double
    uvw[0] = ang * radius;
    return true;
}
// include/impl/plane_traits.hpp:66-72
inline bool plane_project(const m4& g2l, float threshold, v3 xyz, float uvw[3]) {
    float loc[4];
    g2l_apply(g2l, xyz, loc);
    if (fabsf(loc[2]) > threshold) return false;
    uvw[0] = loc[0]; uvw[1] = loc[1]; uvw[2] = loc[2];
    return true;
}
// include/impl/plane2_traits.hpp:86-89
inline bool plane2_project(const m4& g2l, v3 xyz, float uvw[3]) {
    float loc[4];
    g2l_apply(g2l, xyz, loc);
    uvw[0] = loc[0]; uvw[1] = loc[1]; uvw[2] = loc[2];
    return true;
}
// include/impl/identity_traits.hpp:33-36
inline bool identity_project(v3 xyz, float uvw[3]) {
    uvw[0] = xyz.x; uvw[1] = xyz.y; uvw[2] = xyz.z;
    return true;
}

// --------------------------------------- k-NN + principal curvatures (8f rank 2)
// pointcloud::knn_inclusive (include/impl/pointcloud.hpp:138-152) by brute force, ascending
// (d^2, index) with d^2 = (dx*dx + dy*dy) + dz*dz.  [parity unpinned: FLANN's order among equal
// distances]
inline void knn_inclusive(const cloud& c, uint32_t q, uint32_t k, std::vector<int32_t>& idx, std::vector<float>& d2) {
    std::vector<std::pair<float, uint32_t>> all;
    all.reserve(c.n);
    for (uint32_t i = 0; i < c.n; ++i) {
        float d = sqdist_seq(ld3(c.pos, i), ld3(c.pos, q));
        if (d == d) all.push_back({d, i});
    }
    uint32_t kk = std::min<uint32_t>(k, (uint32_t)all.size());
    std::partial_sort(all.begin(), all.begin() + kk, all.end());
    idx.assign(k, -1);
    d2.assign(k, 3.4e38f);
    for (uint32_t j = 0; j < kk; ++j) {
        idx[j] = (int32_t)all[j].second;
        d2[j] = all[j].first;
    }
}
// pcl::eigen33 (pcl/common/impl/eigen.hpp; third-party, restated from the published algorithm):
// closed-form eigenvalues of a symmetric 3x3, ascending.  [parity unpinned: PCL version; cos / sin: tm_sincosf.h]
inline void pcl_roots2(float b, float c, float r[3]) {
    r[0] = 0.f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);
    if (d < 0.f) d = 0.f;
    float sd = std::sqrt(d);
    r[2] = 0.5f * (b + sd);
    r[1] = 0.5f * (b - sd);
}
inline void pcl_eigen33(const float cov[3][3], float evals[3]) {
    float scale = 0.f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(cov[i][j]));
    if (scale <= std::numeric_limits<float>::min()) scale = 1.f;
    float m[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i][j] = cov[i][j] / scale;
    float c0 = m[0][0] * m[1][1] * m[2][2] + 2.f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
               m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
               m[1][2] * m[1][2];
    float c2 = m[0][0] + m[1][1] + m[2][2];
    float r[3];
    if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
        pcl_roots2(c2, c1, r);
    } else {
        const float s_inv3 = (float)(1.0 / 3.0);
        const float s_sqrt3 = std::sqrt(3.0f);
        float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
        if (a_over_3 > 0.f) a_over_3 = 0.f;
        float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
        if (q > 0.f) q = 0.f;
        float rho = std::sqrt(-a_over_3);
        float theta = atan2f_full(std::sqrt(-q), half_b) * s_inv3;
        // theta in [0, pi/3].  The reference takes cos / sin from its libm; glibc's binary32 routines are not
        // correctly rounded and no device routine reproduces them bit for bit, so oracle and device share
        // tm_sincosf.h (binary64 Taylor series, rounded once: equal to (float)cos((double)theta) for every
        // binary32 theta in [0, 1.6], within 1 ulp of glibc's cosf / sinf).
        float cos_theta = tm_math::cosf_small(theta);
        float sin_theta = tm_math::sinf_small(theta);
        r[0] = c2_over_3 + 2.f * rho * cos_theta;
        r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
        r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
        if (r[0] >= r[1]) std::swap(r[0], r[1]);
        if (r[1] >= r[2]) {
            std::swap(r[1], r[2]);
            if (r[0] >= r[1]) std::swap(r[0], r[1]);
        }
        if (r[0] <= 0.f) pcl_roots2(c2, c1, r);
    }
    for (int i = 0; i < 3; ++i) evals[i] = r[i] * scale;
}
// include/impl/pointcloud.hpp:3-44 — principal_curvatures over the given neighbours
inline void principal_curvatures(const cloud& c, uint32_t p_idx, const int32_t* indices, uint32_t n_idx,
                                 float cov[3][3], float& pc_min, float& pc_max) {
    v3 n = ld3(c.nrm, p_idx);
    float nv[3] = {n.x, n.y, n.z};
    float M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = (i == j ? 1.f : 0.f) - nv[i] * nv[j];
    std::vector<std::array<float, 3>> proj(n_idx);
    float cen[3] = {0.f, 0.f, 0.f};
    for (uint32_t k = 0; k < n_idx; ++k) {
        v3 nn = ld3(c.nrm, (uint32_t)indices[k]);
        for (int a = 0; a < 3; ++a) proj[k][a] = M[a][0] * nn.x + (M[a][1] * nn.y + M[a][2] * nn.z);
        for (int a = 0; a < 3; ++a) cen[a] = cen[a] + (proj[k][a] - cen[a]) / (float)(k + 1);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[i][j] = 0.f;
    for (uint32_t k = 0; k < n_idx; ++k) {
        float d[3] = {proj[k][0] - cen[0], proj[k][1] - cen[1], proj[k][2] - cen[2]};
        double xy = d[0] * d[1], xz = d[0] * d[2], yz = d[1] * d[2];
        cov[0][0] += d[0] * d[0]; cov[0][1] += (float)xy; cov[0][2] += (float)xz;
        cov[1][0] += (float)xy; cov[1][1] += d[1] * d[1]; cov[1][2] += (float)yz;
        cov[2][0] += (float)xz; cov[2][1] += (float)yz; cov[2][2] += d[2] * d[2];
    }
    float ev[3];
    pcl_eigen33(cov, ev);
    float area_inv = 1.0f / (float)n_idx;
    pc_min = ev[1] * area_inv;
    pc_max = ev[2] * area_inv;
}

// ------------------------------------------------------- opencl/*.cl (row a15)
// The reference ships these OpenCL kernels without host code; restated per work-item.
// [parity unpinned: OpenCL builtins atan2pi / length / convert_int and the device compiler's
// contraction choices are third-party; fixed here as: no FMA, atan2pi(y,x) = atan2f(y,x)/pi_f32,
// length = sqrtf of the left-to-right sum of squares, convert_int = round toward zero,
// saturating, NaN -> 0, integer adds wrap.]
// opencl/util.cl:1-9
inline void cl_mat44_multiply(const float p[4], const float* mat, float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = ((mat[r] * p[0] + mat[4 + r] * p[1]) + mat[8 + r] * p[2]) + mat[12 + r] * p[3];
}
// opencl/cylinder.cl:1-25
inline void cl_uv_project(const float loc[4], const float* cyl2ncoord, float out[4]) {
    float nc[4];
    cl_mat44_multiply(loc, cyl2ncoord, nc);
    float u = atan2f_full(nc[1], nc[0]) / 3.14159274101257324219f;
    if (u < 0.f) u += 2.f;
    u /= 2.f;
    out[0] = u;
    out[1] = nc[2];
    out[2] = sqrtf(nc[0] * nc[0] + nc[1] * nc[1]) - 1.0f;
    out[3] = 1.f;
}
inline int32_t cl_convert_int(float v) {
    if (!(v == v)) return 0;
    if (v >= 2147483648.f) return INT32_MAX;
    if (v <= -2147483648.f) return INT32_MIN;
    return (int32_t)v;
}
// opencl/icp.cl:1-53, one work-item.  projector 0 = cylinder.cl's uv_project, 1 = linear.
inline void cl_icp_projection(int projector, const float* pnts4, int index, const float* image4,
                              const int32_t img_size[2], const int32_t img_margin[2], const float* mat_align,
                              const float* mat_uvw, const float* mat_proj, const float* mat_norm,
                              float max_corr_dist, float* out_positions4, int32_t* model_indices,
                              int32_t* scene_indices) {
    const float* pnt = pnts4 + 4 * (size_t)index;
    float loc[4], uv[4], tmp[4], uv_nrm[4];
    cl_mat44_multiply(pnt, mat_align, loc);
    if (projector == 0) cl_uv_project(loc, mat_proj, uv);
    else cl_mat44_multiply(loc, mat_proj, uv);
    cl_mat44_multiply(uv, mat_norm, tmp);
    cl_mat44_multiply(tmp, mat_uvw, uv_nrm);
    float ext[2] = {(float)(img_size[0] - 2 * img_margin[0] - 1), (float)(img_size[1] - 2 * img_margin[1] - 1)};
    int32_t px[2];
    for (int a = 0; a < 2; ++a)
        px[a] = (int32_t)((uint32_t)cl_convert_int(uv_nrm[a] * ext[a]) + (uint32_t)img_margin[a]);
    if (px[1] == img_size[1]) px[1] = img_size[1] - 1;
    model_indices[index] = -1;
    scene_indices[index] = -1;
    float* op = out_positions4 + 4 * (size_t)index;
    op[0] = op[1] = op[2] = op[3] = 0.f;
    if (px[0] >= 0 && px[0] < img_size[0] && px[1] >= 0 && px[1] < img_size[1]) {
        int idx = px[1] * img_size[0] + px[0];
        float dx = image4[4 * (size_t)idx] - uv_nrm[0], dy = image4[4 * (size_t)idx + 1] - uv_nrm[1];
        float dist = sqrtf(dx * dx + dy * dy);
        op[3] = dist;
        if (dist < max_corr_dist) {
            model_indices[index] = idx;
            scene_indices[index] = index;
            for (int a = 0; a < 4; ++a) op[a] = uv_nrm[a];
        }
    }
}
// opencl/icp.cl:55-86, one work-item -> 16 floats
inline void cl_icp_correlation(const float* scene4, const float* model4, const int32_t* indices_scene,
                               const int32_t* indices_model, int n, int index, const float* centroid_scene,
                               const float* centroid_model, float out16[16]) {
    const float* s = scene4 + 4 * (size_t)indices_scene[index];
    const float* m = model4 + 4 * (size_t)indices_model[index];
    float spt[3], mpt[3];
    for (int a = 0; a < 3; ++a) {
        spt[a] = s[a] - centroid_scene[a];
        mpt[a] = m[a] - centroid_model[a];
    }
    float norm = 1.f / (float)(n - 1);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) out16[3 * j + i] = spt[i] * mpt[j] * norm;
    for (int a = 9; a < 16; ++a) out16[a] = 0.f;
}

// --------------------------------------------------------------------- octree
// include/impl/octree.hpp:11-17 — octant bit i = pos[i] > center[i]
inline uint8_t get_octant(v3 center, v3 pos) {
    return (uint8_t)((pos.x > center.x ? 1 : 0) | (pos.y > center.y ? 2 : 0) | (pos.z > center.z ? 4 : 0));
}

}  // namespace orc
