// k_sort.cu — Z-curve (Morton) ordering of a cloud on the device.  The radius search, the k-NN and
// the scorer's tile culling all rely on consecutive points being close in space; callers with clouds
// in arbitrary order get that order here (tm_scene_upload_sorted) instead of sorting on the host.
//   morton_code_kernel    30-bit code (10 bits per axis over the cloud's bounding box) per point.
//   radix_hist_kernel /   stable LSD radix sort, 8-bit digits, 4 passes over (code, index) pairs:
//   radix_scatter_kernel  per-block digit histograms -> global exclusive scan (digit-major) ->
//                         order-preserving scatter (ranks inside a block from __match_any_sync in
//                         element order, warps taking turns).
//   gather_cloud_kernel   pos / nrm / tgt permuted into the sorted order.
// Stable: points with equal codes keep their original relative order.
#include "tm_kernels.cuh"

namespace tmk {

constexpr int RS_TILE = 2048;  // elements per block and pass

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void __launch_bounds__(256)
    morton_code_kernel(const float4* __restrict__ pos, uint32_t n, float lox, float loy, float loz, float ix, float iy,
                       float iz, uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    float t[3] = {(p.x - lox) * ix, (p.y - loy) * iy, (p.z - loz) * iz};
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = t[k];
        if (!(v >= 0.f)) v = 0.f;  // also NaN
        if (v > 1.f) v = 1.f;
        q[k] = (uint32_t)(v * 1023.f);
    }
    codes[i] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    idx[i] = i;
}
__global__ void __launch_bounds__(256)
    radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t n_blocks,
                      uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int s = 0; s < RS_TILE / 256; ++s) {
        const uint32_t i = base + s * 256 + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}
__global__ void __launch_bounds__(256)
    radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, int shift,
                         uint32_t n_blocks, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ out_keys,
                         uint32_t* __restrict__ out_vals) {
    __shared__ uint32_t digit_off[256];  // next free slot of (digit, this block)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    digit_off[threadIdx.x] = offsets[(size_t)threadIdx.x * n_blocks + blockIdx.x];
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int s = 0; s < RS_TILE / 256; ++s) {
        const uint32_t i = base + s * 256 + threadIdx.x;
        const bool live = i < n;
        uint32_t key = 0, val = 0, d = 256u + (uint32_t)lane;  // dead lanes: unique pseudo digits, no peers
        if (live) {
            key = keys[i];
            val = vals[i];
            d = (key >> shift) & 255u;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t dst = 0;
        for (int w = 0; w < 8; ++w) {  // warps take turns so that element order is kept across warps
            if (warp == w) {
                uint32_t b = 0;
                if (live && lane == leader) {
                    b = digit_off[d];
                    digit_off[d] = b + __popc(peers);
                }
                b = __shfl_sync(0xffffffffu, b, leader);
                dst = b + rank;
            }
            __syncthreads();
        }
        if (live) {
            out_keys[dst] = key;
            out_vals[dst] = val;
        }
    }
}
__global__ void __launch_bounds__(256)
    gather_cloud_kernel(const float4* __restrict__ pos, const float4* __restrict__ nrm, const float4* __restrict__ tgt,
                        const uint32_t* __restrict__ perm, uint32_t n, float4* __restrict__ opos,
                        float4* __restrict__ onrm, float4* __restrict__ otgt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = perm[i];
    opos[i] = pos[s];
    onrm[i] = nrm[s];
    otgt[i] = tgt[s];
}

void launch_morton_codes(cudaStream_t st, const float4* pos, uint32_t n, const float lo[3], const float inv[3],
                         uint32_t* codes, uint32_t* idx) {
    if (!n) return;
    ++g_launch_count;
    morton_code_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, n, lo[0], lo[1], lo[2], inv[0], inv[1], inv[2], codes, idx);
}
uint32_t radix_blocks(uint32_t n) { return (n + RS_TILE - 1) / RS_TILE; }
// keys/vals in `a`, scratch in `b`; after 4 passes the result is back in `a`.  hist: 256 * blocks + 1 u32 (x2).
void launch_radix_sort_pairs(cudaStream_t st, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                             uint32_t n, uint32_t* hist, uint32_t* offsets, unsigned long long* scan_scratch,
                             int max_ctas) {
    if (!n) return;
    const uint32_t nb = radix_blocks(n);
    uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        g_launch_count += 2;
        radix_hist_kernel<<<nb, 256, 0, st>>>(ki, n, shift, nb, hist);
        launch_exclusive_scan_u32_chained(st, hist, offsets, (uint64_t)256 * nb, scan_scratch, max_ctas);
        radix_scatter_kernel<<<nb, 256, 0, st>>>(ki, vi, n, shift, nb, offsets, ko, vo);
        uint32_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
}
void launch_gather_cloud(cudaStream_t st, const float4* pos, const float4* nrm, const float4* tgt, const uint32_t* perm,
                         uint32_t n, float4* opos, float4* onrm, float4* otgt) {
    if (!n) return;
    ++g_launch_count;
    gather_cloud_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, nrm, tgt, perm, n, opos, onrm, otgt);
}

}  // namespace tmk
