// k_knn.cu — pre-processing the reference does with PCL/FLANN before the search
// (SURVEY §8f rank 2): exact k-nearest neighbours and the principal-curvature test behind the
// tangent masks (scene.hpp:46-58, model.hpp:68-71,96-99 -> pointcloud.hpp:200-204, 3-44).
//
//   knn_kernel        one warp per query point; the k <= 32 best (d^2, index) pairs live one per
//                     lane, sorted.  Phase 1 scans the query's own 1024-point segment and its two
//                     neighbours (close in space when the cloud is in a space-filling-curve order)
//                     to get an upper bound on the k-th distance; phase 2 screens the remaining
//                     segments 32 at a time by box distance against the current bound and scans
//                     the survivors.  Exact for any point order (loose boxes only cost time).
//                     d^2 = (dx*dx + dy*dy) + dz*dz (FLANN L2_Simple); ties -> lower index.
//   curvature_kernel  pointcloud.hpp:3-44 restated in binary32 (projection of the neighbours'
//                     normals into the tangent plane, running-mean centroid, covariance in the
//                     reference's accumulation order) + pcl::eigen33's closed-form eigenvalues.
#include "tm_kernels.cuh"
#include "../../include/triplet_match/tm_sincosf.h"

namespace tmk {

struct KnnEntry {
    float d2;
    uint32_t idx;
};
__device__ __forceinline__ bool knn_less(float d2a, uint32_t ia, float d2b, uint32_t ib) {
    return d2a < d2b || (d2a == d2b && ia < ib);
}
// insert (d2, idx) into the warp-resident sorted list of length k (lane l = l-th best)
__device__ __forceinline__ void knn_insert(KnnEntry& mine, int lane, uint32_t k, float d2, uint32_t idx) {
    const bool before = knn_less(mine.d2, mine.idx, d2, idx);  // my entry stays in front of the new one
    const uint32_t m = __ballot_sync(0xffffffffu, before) & (k >= 32 ? 0xffffffffu : ((1u << k) - 1u));
    const int p = __popc(m);  // entries are sorted, so `before` is a prefix: insertion position
    const float up_d = __shfl_up_sync(0xffffffffu, mine.d2, 1);
    const uint32_t up_i = __shfl_up_sync(0xffffffffu, mine.idx, 1);
    if (lane > p) {
        mine.d2 = up_d;
        mine.idx = up_i;
    } else if (lane == p) {
        mine.d2 = d2;
        mine.idx = idx;
    }
}
__device__ __forceinline__ void knn_scan_segment(const float4* __restrict__ pos, uint32_t n, uint32_t seg,
                                                 float qx, float qy, float qz, KnnEntry& mine, int lane, uint32_t k) {
    const uint32_t base = seg * BALL_SEG;
    for (uint32_t s = 0; s < BALL_SEG / 32; ++s) {
        const uint32_t i = base + s * 32 + lane;
        float d2 = 3.4e38f;
        bool ok = false;
        if (i < n) {
            const float4 p = pos[i];
            const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            d2 = (dx * dx + dy * dy) + dz * dz;
            ok = d2 == d2;  // NaN coordinates never enter
        }
        const float wd = __shfl_sync(0xffffffffu, mine.d2, (int)k - 1);
        const uint32_t wi = __shfl_sync(0xffffffffu, mine.idx, (int)k - 1);
        uint32_t cand = __ballot_sync(0xffffffffu, ok && knn_less(d2, i, wd, wi));
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1u;
            const float nd = __shfl_sync(0xffffffffu, d2, src);
            const uint32_t ni = base + s * 32 + src;
            const float cwd = __shfl_sync(0xffffffffu, mine.d2, (int)k - 1);
            const uint32_t cwi = __shfl_sync(0xffffffffu, mine.idx, (int)k - 1);
            if (knn_less(nd, ni, cwd, cwi)) knn_insert(mine, lane, k, nd, ni);
        }
    }
}

__global__ void __launch_bounds__(256)
    knn_kernel(CloudDev cloud, const uint32_t* __restrict__ query, uint32_t n_query, uint32_t k,
               int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
    const int lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_query) return;
    const uint32_t qi = query[w];
    const float4 q = cloud.pos[qi];
    const uint32_t n_seg = (cloud.n + BALL_SEG - 1) / BALL_SEG;
    KnnEntry mine{3.4e38f, 0xffffffffu};
    // phase 1: own segment and its neighbours
    const uint32_t own = qi / BALL_SEG;
    const uint32_t s_lo = own > 0 ? own - 1 : 0, s_hi = min(own + 1, n_seg - 1);
    for (uint32_t s = s_lo; s <= s_hi; ++s) knn_scan_segment(cloud.pos, cloud.n, s, q.x, q.y, q.z, mine, lane, k);
    // phase 2: every other segment whose box can still hold a better neighbour
    for (uint32_t s0 = 0; s0 < n_seg; s0 += 32) {
        const uint32_t s = s0 + lane;
        const float wd0 = __shfl_sync(0xffffffffu, mine.d2, (int)k - 1);  // current k-th distance (uniform)
        bool cand = s < n_seg && (s < s_lo || s > s_hi);
        if (cand && cloud.seg_lo) {
            const float4 lo = cloud.seg_lo[s], hi = cloud.seg_hi[s];
            const float dx = fmaxf(fmaxf(lo.x - q.x, q.x - hi.x), 0.f);
            const float dy = fmaxf(fmaxf(lo.y - q.y, q.y - hi.y), 0.f);
            const float dz = fmaxf(fmaxf(lo.z - q.z, q.z - hi.z), 0.f);
            const float bd = dx * dx + dy * dy + dz * dz;
            cand = !(bd > wd0 * 1.00002f + 1e-30f);  // ties (equal d^2, lower index) stay reachable
        }
        uint32_t todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t ss = s0 + src;
            if (cloud.seg_lo) {  // the bound may have tightened since the ballot
                const float4 lo = cloud.seg_lo[ss], hi = cloud.seg_hi[ss];
                const float dx = fmaxf(fmaxf(lo.x - q.x, q.x - hi.x), 0.f);
                const float dy = fmaxf(fmaxf(lo.y - q.y, q.y - hi.y), 0.f);
                const float dz = fmaxf(fmaxf(lo.z - q.z, q.z - hi.z), 0.f);
                const float bd = dx * dx + dy * dy + dz * dz;
                const float wd = __shfl_sync(0xffffffffu, mine.d2, (int)k - 1);
                if (bd > wd * 1.00002f + 1e-30f) continue;
            }
            knn_scan_segment(cloud.pos, cloud.n, ss, q.x, q.y, q.z, mine, lane, k);
        }
    }
    if ((uint32_t)lane < k) {
        out_idx[(size_t)w * k + lane] = mine.idx == 0xffffffffu ? -1 : (int32_t)mine.idx;
        if (out_d2) out_d2[(size_t)w * k + lane] = mine.d2;
    }
}
void launch_knn(cudaStream_t st, const CloudDev& cloud, const uint32_t* query, uint32_t n_query, uint32_t k,
                int32_t* out_idx, float* out_d2) {
    if (!n_query) return;
    ++g_launch_count;
    knn_kernel<<<(n_query + 7) / 8, 256, 0, st>>>(cloud, query, n_query, k, out_idx, out_d2);
}

// pcl::eigen33 eigenvalues (ascending) of a symmetric 3x3, closed form (pcl/common/impl/eigen.hpp,
// third-party: restated from the published algorithm)
__device__ __forceinline__ void roots2(float b, float c, float r[3]) {
    r[0] = 0.f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);
    if (d < 0.f) d = 0.f;
    const float sd = sqrtf(d);
    r[2] = 0.5f * (b + sd);
    r[1] = 0.5f * (b - sd);
}
__device__ inline void eigen33_values(const float cov[3][3], float evals[3]) {
    float scale = 0.f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(cov[i][j]));
    if (scale <= 1.17549435e-38f) scale = 1.f;
    float m[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i][j] = cov[i][j] / scale;
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
                     m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
                     m[1][2] * m[1][2];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    float r[3];
    if (fabsf(c0) < 1.1920929e-07f) {
        roots2(c2, c1, r);
    } else {
        const float s_inv3 = (float)(1.0 / 3.0);
        const float s_sqrt3 = sqrtf(3.0f);
        const float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
        if (a_over_3 > 0.f) a_over_3 = 0.f;
        const float half_b = 0.5f * (c0 + c2_over_3 * (2.f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
        if (q > 0.f) q = 0.f;
        const float rho = sqrtf(-a_over_3);
        const float theta = atan2f_full(sqrtf(-q), half_b) * s_inv3;
        // theta in [0, pi/3]: the shared binary64-Taylor routine, same bits as the oracle (tm_sincosf.h)
        const float cos_theta = tm_math::cosf_small(theta);
        const float sin_theta = tm_math::sinf_small(theta);
        r[0] = c2_over_3 + 2.f * rho * cos_theta;
        r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
        r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
        float t;
        if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
        if (r[1] >= r[2]) {
            t = r[1]; r[1] = r[2]; r[2] = t;
            if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
        }
        if (r[0] <= 0.f) roots2(c2, c1, r);
    }
    for (int i = 0; i < 3; ++i) evals[i] = r[i] * scale;
}

// one thread per query: pointcloud.hpp:3-44 over its k neighbours (in (d^2, index) order)
__global__ void __launch_bounds__(128)
    curvature_kernel(CloudDev cloud, const uint32_t* __restrict__ query, uint32_t n_query, uint32_t k,
                     const int32_t* __restrict__ nbr, float* __restrict__ pc_min, float* __restrict__ pc_max,
                     float* __restrict__ cov_out) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_query) return;
    const float4 nq = cloud.nrm[query[w]];
    // M = I - n n^T
    const float nv[3] = {nq.x, nq.y, nq.z};
    float M[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = (i == j ? 1.f : 0.f) - nv[i] * nv[j];
    const int32_t* my = nbr + (size_t)w * k;
    uint32_t cnt = 0;
    for (uint32_t j = 0; j < k; ++j) cnt += my[j] >= 0;
    float cen[3] = {0.f, 0.f, 0.f};
    for (uint32_t j = 0; j < cnt; ++j) {
        const float4 nn = cloud.nrm[my[j]];
        for (int a = 0; a < 3; ++a) {
            const float pr = M[a][0] * nn.x + (M[a][1] * nn.y + M[a][2] * nn.z);  // 3-redux a0 + (a1 + a2)
            cen[a] = cen[a] + (pr - cen[a]) / (float)(j + 1);
        }
    }
    float cov[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (uint32_t j = 0; j < cnt; ++j) {
        const float4 nn = cloud.nrm[my[j]];
        float d[3];
        for (int a = 0; a < 3; ++a) d[a] = (M[a][0] * nn.x + (M[a][1] * nn.y + M[a][2] * nn.z)) - cen[a];
        const float xy = d[0] * d[1], xz = d[0] * d[2], yz = d[1] * d[2];
        cov[0][0] += d[0] * d[0]; cov[0][1] += xy; cov[0][2] += xz;
        cov[1][0] += xy; cov[1][1] += d[1] * d[1]; cov[1][2] += yz;
        cov[2][0] += xz; cov[2][1] += yz; cov[2][2] += d[2] * d[2];
    }
    float ev[3];
    eigen33_values(cov, ev);
    const float area_inv = 1.0f / (float)cnt;
    pc_min[w] = ev[1] * area_inv;
    pc_max[w] = ev[2] * area_inv;
    if (cov_out)
        for (int i = 0; i < 9; ++i) cov_out[(size_t)w * 9 + i] = cov[i / 3][i % 3];
}
void launch_curvature(cudaStream_t st, const CloudDev& cloud, const uint32_t* query, uint32_t n_query, uint32_t k,
                      const int32_t* nbr, float* pc_min, float* pc_max, float* cov_out) {
    if (!n_query) return;
    ++g_launch_count;
    curvature_kernel<<<(n_query + 127) / 128, 128, 0, st>>>(cloud, query, n_query, k, nbr, pc_min, pc_max, cov_out);
}

// tangent-mask candidates: ||tangent|| > 0.7 (scene.hpp:50, model.hpp:98); compacted ascending
__global__ void __launch_bounds__(256)
    tangent_candidates_kernel(const float4* __restrict__ tgt, uint32_t n, uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 t = tgt[i];
    const float nrm = sqrtf(t.x * t.x + (t.y * t.y + t.z * t.z));
    flags[i] = nrm > 0.7f ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
    compact_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offsets, uint32_t n,
                   uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[offsets[i]] = i;
}
// mask[i] = candidate and pc_min / pc_max < ratio (NaN -> 0); optionally written into pos.w bit 0
__global__ void __launch_bounds__(256)
    tangent_mask_kernel(float4* __restrict__ pos, uint32_t n, const uint32_t* __restrict__ cand, uint32_t n_cand,
                        const float* __restrict__ pc_min, const float* __restrict__ pc_max, float ratio,
                        uint8_t* __restrict__ mask, int apply) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_cand) return;
    const uint32_t i = cand[w];
    const bool t = (pc_min[w] / pc_max[w]) < ratio;
    mask[i] = t ? 1 : 0;
    if (apply) {
        uint32_t fl = __float_as_uint(pos[i].w);
        fl = t ? (fl | FLAG_TANGENT) : (fl & ~FLAG_TANGENT);
        pos[i].w = __uint_as_float(fl);
    }
}
__global__ void __launch_bounds__(256) clear_tangent_flags_kernel(float4* __restrict__ pos, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i].w = __uint_as_float(__float_as_uint(pos[i].w) & ~FLAG_TANGENT);
}
void launch_tangent_candidates(cudaStream_t st, const float4* tgt, uint32_t n, uint32_t* flags) {
    if (!n) return;
    ++g_launch_count;
    tangent_candidates_kernel<<<(n + 255) / 256, 256, 0, st>>>(tgt, n, flags);
}
void launch_compact(cudaStream_t st, const uint32_t* flags, const uint32_t* offsets, uint32_t n, uint32_t* out) {
    if (!n) return;
    ++g_launch_count;
    compact_kernel<<<(n + 255) / 256, 256, 0, st>>>(flags, offsets, n, out);
}
void launch_tangent_mask(cudaStream_t st, float4* pos, uint32_t n, const uint32_t* cand, uint32_t n_cand,
                         const float* pc_min, const float* pc_max, float ratio, uint8_t* mask, int apply) {
    if (apply && n) {
        ++g_launch_count;
        clear_tangent_flags_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, n);
    }
    if (!n_cand) return;
    ++g_launch_count;
    tangent_mask_kernel<<<(n_cand + 255) / 256, 256, 0, st>>>(pos, n, cand, n_cand, pc_min, pc_max, ratio, mask, apply);
}

}  // namespace tmk
