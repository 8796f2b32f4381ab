// tm_device.cuh — device layouts and the bit-exact FP32 helpers shared by all
// kernels.  Every translation unit that includes this file is compiled with
// -fmad=false (no FFMA/DFMA contraction) and nvcc's default -prec-div=true
// -prec-sqrt=true, so +,-,*,/ and sqrt round exactly like the reference's SSE2
// build (CMakeLists.txt:28-40: -O3, no -march, no fast-math).  The evaluation
// orders are the Eigen 3.3 fixed-size orders (see DESIGN.md "Float contract"):
//   Matrix4f*Vector4f : ((c0*x + c1*y) + c2*z) + c3*w
//   3-redux (dot, squaredNorm, Matrix3f row*vec) : a0 + (a1 + a2)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/triplet_match/tm_atan2f.h"

namespace tmk {

// ---- resident layouts ----------------------------------------------------
// Clouds are SoA float4: pos.xyz + flag bits in pos.w, nrm.xyz, tgt.xyz.
constexpr uint32_t FLAG_TANGENT = 1u;  // scene: tangent_mask_ ; model: ||tangent|| > 0.7
constexpr uint32_t FLAG_MASKED = 2u;   // scene: mask_

struct CloudDev {
    const float4* pos;
    const float4* nrm;
    const float4* tgt;
    uint32_t n;
    const float4* seg_lo = nullptr;  // scenes: bounding box of every BALL_SEG-point segment
    const float4* seg_hi = nullptr;
};

struct HashSlot {  // open addressing, linear probing, home = murmur4(key) & mask
    uint32_t k[4];
    uint32_t begin;  // into hits[]
    uint32_t count;  // 0 = empty slot
    uint32_t pad[2];
};

struct ModelDev {
    CloudDev cloud;
    const uint32_t* voxel;  // lin = k*ex*ey + j*ex + i
    const float4* vcell;    // optional fused grid (or null): nearest model point pos.xyz; .w bits =
                            // (model index << 1) | class flag
    const float4* mref;     // per model point: its ref vector (tangent if ||tangent|| > 0.7 else normal),
                            // read only for inliers; n_model x 16 B, small enough to live in L1
    int ex, ey, ez;
    float exf, eyf, ezf;
    float sx, sy, sz, tx, ty, tz;  // to_voxel_ = diag(s) + t
    const HashSlot* slots;
    uint32_t slot_mask;
    const uint2* hits;
    uint32_t n_hits;
    float fb_min0, fb_max0;  // feat_bounds_ distance range
    uint32_t dist_steps;
    float angle_step;
    float resolution, diameter;
    // optional block-occupancy mask for one distance threshold (occupancy_kernel): bit b of 8x8x8-cell
    // block b is clear when no position inside the block can be within the threshold of its cell's
    // nearest model point, so the cell gathers can be skipped there without changing any result
    const uint32_t* occ = nullptr;
    int obx = 0, oby = 0;
};
#ifndef TM_OCC_SHIFT
#define TM_OCC_SHIFT 3
#endif
constexpr int OCC_SHIFT = TM_OCC_SHIFT;  // 2^OCC_SHIFT cells per block edge
__device__ __forceinline__ bool occ_test(const ModelDev& m, int i, int j, int k) {
    const uint32_t b = (uint32_t)(((k >> OCC_SHIFT) * m.oby + (j >> OCC_SHIFT)) * m.obx + (i >> OCC_SHIFT));
    return (__ldg(&m.occ[b >> 5]) >> (b & 31u)) & 1u;
}

// stride of the evenly sampling walk (tm_walk_stride): floor(n * 0.6180339887) made coprime with n
__host__ __device__ inline uint32_t walk_stride(uint32_t n) {
    if (n <= 2u) return 1u;
    uint32_t s = (uint32_t)(((unsigned long long)n * 2654435769ull) >> 32);  // n * (sqrt(5) - 1) / 2
    if (s < 1u) s = 1u;
    for (;; ++s) {
        if (s >= n) s = 1u;
        uint32_t a = n, b = s;
        while (b) { const uint32_t t = a % b; a = b; b = t; }
        if (a == 1u) return s;
    }
}

// first walk position of level L for a subset of n elements: b[0] = 0, b[L] = t_L - 1 (0 when t_L <= 1) with
// t_L = uint32(0.05f * L * n) the reference's tests[L-1] (scene.hpp:422-426), b[19] = n (19 ranges: tmk::EL_LEVELS).  A checkpoint with threshold t
// fires at the first reaching element whose 1-based position is >= t.
__host__ __device__ inline uint32_t level_begin(uint32_t n, int L) {
    if (L <= 0) return 0u;
    if (L >= 19) return n;
    const uint32_t t = (uint32_t)(0.05f * (float)L * (float)n);
    const uint32_t b = t > 1u ? t - 1u : 0u;
    return b < n ? b : n;
}

// hypothesis transform: rows 0..2 of the 4x4 (row 3 is 0,0,0,1)
struct Rows {
    float4 r0, r1, r2;
};

// ---- no-FMA float helpers -------------------------------------------------
__host__ __device__ __forceinline__ float sum3(float a0, float a1, float a2) {
    return a0 + (a1 + a2);
}
struct f3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ f3 mk3(float4 v) { return {v.x, v.y, v.z}; }
__host__ __device__ __forceinline__ f3 sub3(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ float dot3(f3 a, f3 b) {
    return sum3(a.x * b.x, a.y * b.y, a.z * b.z);
}
__host__ __device__ __forceinline__ float sqnorm3(f3 a) { return dot3(a, a); }
__device__ __forceinline__ float norm3(f3 a) { return sqrtf(sqnorm3(a)); }
__host__ __device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ f3 normalized3(f3 a) {
    float z = sqnorm3(a);
    if (z > 0.f) {
        float s = sqrtf(z);
        return {a.x / s, a.y / s, a.z / s};
    }
    return a;
}
// row of Matrix4f * (x,y,z,1): ((r.x*x + r.y*y) + r.z*z) + r.w
__host__ __device__ __forceinline__ float row_apply(float4 r, float x, float y, float z) {
    return ((r.x * x + r.y * y) + r.z * z) + r.w;
}
// row of Matrix3f * v
__host__ __device__ __forceinline__ float row_rot(float4 r, f3 v) {
    return sum3(r.x * v.x, r.y * v.y, r.z * v.z);
}

// ---- atan2f: the reference platform's libm algorithm in binary32 (include/triplet_match/tm_atan2f.h)
__host__ __device__ inline float atan2f_q1(float y, float x) { return tm_math::atan2f_libm(y, x); }
__host__ __device__ inline float atan2f_full(float y, float x) { return tm_math::atan2f_libm(y, x); }

// ---- discretise + murmur (src/discretize.cpp:19-30, impl/discretize.hpp:10-45)
__host__ __device__ __forceinline__ uint32_t discretize_range(float value, float min_value,
                                                              float range_value, uint32_t steps) {
    float nval = (value - min_value) / range_value;
    if (nval < 0.f) return 0u;
    if (nval >= 1.f) return steps - 1u;
    return (uint32_t)(nval * (float)steps);
}
__host__ __device__ __forceinline__ uint32_t discretize_step(float value, float step) {
    return (uint32_t)(value / step);
}
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
    return (x << r) | (x >> (32 - r));
}
__host__ __device__ __forceinline__ uint32_t murmur4(uint32_t k0, uint32_t k1, uint32_t k2,
                                                     uint32_t k3) {
    uint32_t h1 = 42u;
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t ks[4] = {k0, k1, k2, k3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t k = ks[i];
        k *= c1;
        k = rotl32(k, 15);
        k *= c2;
        h1 ^= k;
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

// fixed-point score accumulation: |ref.ref_n| in [0, ~1] -> 2^-36 quanta in u64.
// Integer sums are order-independent, so counts AND scores are reproducible for
// any tiling, launch order or multi-GPU sharding.
constexpr double SCORE_SCALE = 68719476736.0;  // 2^36
__device__ __forceinline__ unsigned long long score_fixed(float term) {
    // == (unsigned long long)((double)term * 2^36): scaling by a power of two is exact in binary32 too,
    // and the conversion truncates toward zero either way
    return __float2ull_rz(term * 68719476736.0f);
}

// The reference's bound (scene.hpp:493-500) restated with defined integer
// arithmetic (the original casts negative doubles to uint32_t): see DESIGN.md.
__device__ __forceinline__ uint32_t early_drop_upper(uint32_t tried, uint32_t nsub, uint32_t corrs) {
    double N = -2.0 - (double)tried;
    double x = -2.0 - (double)nsub;
    double n = -1.0 - (double)corrs;
    double tmp = sqrt((x * n * (N - x) * (N - n)) / (N - 1.0));
    double v = (x * n + tmp) / N;
    uint32_t a = (uint32_t)(unsigned long long)(long long)v;
    double b = -1.0 - (double)a;
    return (uint32_t)(unsigned long long)(long long)b;
}


}  // namespace tmk
