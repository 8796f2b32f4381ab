// k_model.cu — the O(T^2) half of model::init (include/impl/model.hpp:100-149) on the device:
// every ordered pair (i, j), i != j, of the T tangent-subset points goes through the reference's
// filters (distance window :110, collinearity :111), the feature (impl/feature.hpp:15-34) and, in
// the second pass, valid() + discretize_feature().
//   model_pair_bounds_kernel  pass 1: component-wise min / max of the features (feat_bounds_).
//   model_pair_keys_kernel    pass 2: packed key per pair in insertion order (row-major i, j), or
//                             ~0 for pairs that are filtered / invalid.
// The insertion-order bookkeeping (multimap LIFO order, 200-value cap) stays on the host: it is a
// linear integer pass over the keys.
#include "tm_kernels.cuh"

namespace tmk {

__device__ __forceinline__ bool model_pair_feature(const float* __restrict__ pos3, const float* __restrict__ tgt3,
                                                   uint32_t a, uint32_t b, float lower, float upper, float f[4]) {
    if (a == b) return false;
    const f3 pa = {pos3[3 * a], pos3[3 * a + 1], pos3[3 * a + 2]}, pb = {pos3[3 * b], pos3[3 * b + 1], pos3[3 * b + 2]};
    const f3 ta = {tgt3[3 * a], tgt3[3 * a + 1], tgt3[3 * a + 2]}, tb = {tgt3[3 * b], tgt3[3 * b + 1], tgt3[3 * b + 2]};
    const f3 d1 = {pb.x - pa.x, pb.y - pa.y, pb.z - pa.z};
    const float dist1 = sqrtf(dot3(d1, d1));
    const f3 dn = {d1.x / dist1, d1.y / dist1, d1.z / dist1};
    if (dist1 < lower || dist1 > upper) return false;         // model.hpp:110
    if (1.f - fabsf(dot3(dn, ta)) < 0.01f) return false;       // model.hpp:111
    f[0] = dist1;
    f[1] = atan2f_q1(sqrtf(sqnorm3(cross3(d1, ta))), fabsf(dot3(d1, ta)));  // impl/feature.hpp:5-8
    f[2] = atan2f_q1(sqrtf(sqnorm3(cross3(d1, tb))), fabsf(dot3(d1, tb)));
    f[3] = f[0];
    return true;
}

// bounds[0..2] = min of f0, f1, f2 (uint bit patterns of non-negative floats order like the floats),
// bounds[3..5] = max, count = pairs that passed the filters
__global__ void __launch_bounds__(256)
    model_pair_bounds_kernel(const float* __restrict__ pos3, const float* __restrict__ tgt3, uint32_t T, float lower,
                             float upper, uint32_t* __restrict__ bounds, unsigned long long* __restrict__ count) {
    const unsigned long long total = (unsigned long long)T * T;
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {0.f, 0.f, 0.f};
    uint32_t passed = 0;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < total;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        float f[4];
        if (!model_pair_feature(pos3, tgt3, (uint32_t)(s / T), (uint32_t)(s % T), lower, upper, f)) continue;
        ++passed;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (f[k] < mn[k]) mn[k] = f[k];  // std::min / std::max semantics: NaN never replaces
            if (f[k] > mx[k]) mx[k] = f[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
        }
    }
    const uint32_t wp = __reduce_add_sync(0xffffffffu, passed);
    if ((threadIdx.x & 31) == 0 && wp) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bounds[k], __float_as_uint(mn[k]));
            atomicMax(&bounds[3 + k], __float_as_uint(mx[k]));
        }
        atomicAdd(count, (unsigned long long)wp);
    }
}

// key = k0 | k1 << 24 | k2 << 44 (k0 < 2^24, k1, k2 < 2^20; k3 == k0), ~0 = no entry
__global__ void __launch_bounds__(256)
    model_pair_keys_kernel(const float* __restrict__ pos3, const float* __restrict__ tgt3, uint32_t T, float lower,
                           float upper, float fmn0, float fmx0, uint32_t steps, float angle_step,
                           unsigned long long* __restrict__ keys) {
    const unsigned long long total = (unsigned long long)T * T;
    const float pi = 3.14159274101257324219f;  // static_cast<float>(M_PI)
    const float diag0 = fmx0 - fmn0;
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < total;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        float f[4];
        unsigned long long key = ~0ull;
        if (model_pair_feature(pos3, tgt3, (uint32_t)(s / T), (uint32_t)(s % T), lower, upper, f)) {
            const bool valid = !(f[0] < fmn0 || f[0] > fmx0) && (f[1] >= 0.f && f[1] <= pi) && (f[2] >= 0.f && f[2] <= pi);
            if (valid) {
                const unsigned long long k0 = discretize_range(f[0], fmn0, diag0, steps);
                const unsigned long long k1 = discretize_step(f[1], angle_step), k2 = discretize_step(f[2], angle_step);
                key = k0 | (k1 << 24) | (k2 << 44);
            }
        }
        keys[s] = key;
    }
}

void launch_model_pair_bounds(cudaStream_t st, const float* pos3, const float* tgt3, uint32_t T, float lower, float upper,
                              uint32_t* bounds, unsigned long long* count, int grid) {
    if (!T) return;
    ++g_launch_count;
    model_pair_bounds_kernel<<<grid, 256, 0, st>>>(pos3, tgt3, T, lower, upper, bounds, count);
}
void launch_model_pair_keys(cudaStream_t st, const float* pos3, const float* tgt3, uint32_t T, float lower, float upper,
                            float fmn0, float fmx0, uint32_t steps, float angle_step, unsigned long long* keys, int grid) {
    if (!T) return;
    ++g_launch_count;
    model_pair_keys_kernel<<<grid, 256, 0, st>>>(pos3, tgt3, T, lower, upper, fmn0, fmx0, steps, angle_step, keys);
}

}  // namespace tmk
