// k_icp.cu — ICP refinement (scene::impl::icp_, include/impl/scene.hpp:369-404).
// Natively re-creates what the orphaned opencl/icp.cl sketched (icp_projection:
// per-point nearest-neighbour correspondence; icp_correlation: per-correspondence
// outer products) plus the reduction and the rigid solve its absent host did:
//
//   icp_accumulate_kernel  correspondences at 2*dist_thres for K transforms over
//                          the whole scene, fused with the reduction of
//                          n, sum s, sum m, sum s m^T (and the score) into
//                          64-bit fixed-point atomics — integer sums are
//                          order-independent, so the result is reproducible and
//                          can be all-reduced exactly across GPUs.
//   icp_step_kernel        per transform: the reference's accept/stop rule and
//                          Eigen::umeyama(scene, model, false) restated on the
//                          sums (3x3 one-sided Jacobi SVD in double).
//   corr_count/fill        scene_corrs / model_corrs lists of one transform in
//                          ascending scene order (finish_find, scene.hpp:100-106).
#include <algorithm>

#include "tm_kernels.cuh"

namespace tmk {

template <bool FUSED>
__device__ __forceinline__ bool icp_point_test(const ModelDev& m, float4 r0, float4 r1, float4 r2,
                                               float4 v, float sq_thres, float& x, float& y,
                                               float& z, float4& mp, uint32_t& lin) {
    uint32_t fl = __float_as_uint(v.w);
    if (fl & FLAG_MASKED) return false;
    x = row_apply(r0, v.x, v.y, v.z);
    y = row_apply(r1, v.x, v.y, v.z);
    z = row_apply(r2, v.x, v.y, v.z);
    float vx = m.sx * x + m.tx, vy = m.sy * y + m.ty, vz = m.sz * z + m.tz;
    bool inb = (vx > -1.f) & (vx < m.exf) & (vy > -1.f) & (vy < m.eyf) & (vz > -1.f) & (vz < m.ezf);
    if (!inb) return false;
    int i = (int)vx, j = (int)vy, k = (int)vz;
    lin = (uint32_t)((k * m.ey + j) * m.ex + i);
    if (m.occ && !occ_test(m, i, j, k)) return false;
    if (FUSED) {
        mp = __ldg(&m.vcell[lin]);
    } else {
        mp = __ldg(&m.cloud.pos[__ldg(&m.voxel[lin])]);
    }
    float dx = x - mp.x, dy = y - mp.y, dz = z - mp.z;
    float sq = sum3(dx * dx, dy * dy, dz * dz);
    if (sq > sq_thres) return false;
    return ((fl ^ __float_as_uint(mp.w)) & FLAG_TANGENT) == 0u;
}

__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// sums[h][ICP_NSUM]: 0 count, 1..3 sum s', 4..6 sum m', 7..15 sum s'_a m'_b, 16 score.
// s' = (T s) - c, m' = m - c with c the model bbox centre; quantum 2^-fix_bits.
//
// Two launches per pass.
//   icp_cull_kernel        one warp per BALL_SEG-point segment of the scene: its bounding box is tested against every
//                          transform (lane l screens transform h0 + l with the scorer's interval test; NaN never
//                          culls) and the surviving (segment, transform) pairs go to a compact list.  A refinement
//                          touches the few dozen segments around each instance, so a pass over a 10 M-point scene
//                          reads 20 k boxes instead of 10 M points.
//   icp_accumulate_kernel  one warp per (pair, 128-point chunk): exact test of 4 points per lane, the 17 fixed-point
//                          sums reduced in the warp and added with at most 17 atomics.  Integer sums: any order,
//                          same bits (the list order is not deterministic, the result is).
__global__ void __launch_bounds__(256)
    icp_cull_kernel(CloudDev scene, ModelDev model, const float4* __restrict__ T, const uint32_t* __restrict__ active,
                    uint32_t h_begin, uint32_t h_end, uint32_t pt_begin, uint32_t pt_end, uint2* __restrict__ pairs,
                    uint32_t* __restrict__ n_pairs) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool have_box = scene.seg_lo != nullptr;
    const uint32_t seg_first = pt_begin / BALL_SEG, seg_last = (pt_end - 1u) / BALL_SEG;  // pt_end > pt_begin
    for (uint32_t seg = seg_first + warp_id; seg <= seg_last; seg += warps_total) {
        float cx_ = 0.f, cy_ = 0.f, cz_ = 0.f, hx_ = 0.f, hy_ = 0.f, hz_ = 0.f;
        if (have_box) {
            const float4 lo = scene.seg_lo[seg], hi = scene.seg_hi[seg];
            if (!(lo.x <= hi.x)) continue;  // no finite point in the segment
            cx_ = 0.5f * (lo.x + hi.x); hx_ = 0.5f * (hi.x - lo.x);
            cy_ = 0.5f * (lo.y + hi.y); hy_ = 0.5f * (hi.y - lo.y);
            cz_ = 0.5f * (lo.z + hi.z); hz_ = 0.5f * (hi.z - lo.z);
        }
        for (uint32_t h0 = h_begin; h0 < h_end; h0 += 32) {
            bool live = false;
            const uint32_t hl = h0 + lane;
            if (hl < h_end && (!active || active[hl])) {
                live = true;
                if (have_box) {
                    const float4 q0 = __ldg(&T[3 * hl]), q1 = __ldg(&T[3 * hl + 1]), q2 = __ldg(&T[3 * hl + 2]);
                    const float acx = fabsf(cx_) + hx_, acy = fabsf(cy_) + hy_, acz = fabsf(cz_) + hz_;
                    bool out = false;
#define TM_AXIS(r, S, TV, EXF)                                                                  \
    {                                                                                           \
        float cc = r.x * cx_ + r.y * cy_ + r.z * cz_ + r.w;                                     \
        float ee = fabsf(r.x) * hx_ + fabsf(r.y) * hy_ + fabsf(r.z) * hz_;                      \
        float mag = fabsf(r.x) * acx + fabsf(r.y) * acy + fabsf(r.z) * acz + fabsf(r.w);        \
        ee += 1e-5f * mag + 1e-30f;                                                             \
        float sl = 1e-5f * (S * mag + fabsf(TV)) + 1e-30f;                                      \
        float lo_ = S * (cc - ee) + TV - sl, hi_ = S * (cc + ee) + TV + sl;                     \
        out = out || (lo_ >= EXF) || (hi_ <= -1.0f);                                            \
    }
                    TM_AXIS(q0, model.sx, model.tx, model.exf)
                    TM_AXIS(q1, model.sy, model.ty, model.eyf)
                    TM_AXIS(q2, model.sz, model.tz, model.ezf)
#undef TM_AXIS
                    live = !out;
                }
            }
            const uint32_t todo = __ballot_sync(0xffffffffu, live);
            if (!todo) continue;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(n_pairs, (uint32_t)__popc(todo));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) pairs[base + __popc(todo & ((1u << lane) - 1u))] = make_uint2(seg, hl);
        }
    }
}

constexpr uint32_t ICP_CHUNK = 32u * ICP_P;                  // points per (pair, chunk) item
constexpr uint32_t ICP_CHUNKS_PER_SEG = BALL_SEG / ICP_CHUNK;
template <bool FUSED>
__global__ void __launch_bounds__(256)
    icp_accumulate_kernel(CloudDev scene, ModelDev model, const float4* __restrict__ T, const uint2* __restrict__ pairs,
                          const uint32_t* __restrict__ n_pairs, uint32_t pt_begin, uint32_t pt_end, float sq_thres,
                          float cx, float cy, float cz, double fix_scale, long long* __restrict__ sums) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned long long n_items = (unsigned long long)(*n_pairs) * ICP_CHUNKS_PER_SEG;
    for (unsigned long long item = warp_id; item < n_items; item += warps_total) {
        const uint2 pr = pairs[item / ICP_CHUNKS_PER_SEG];
        const uint32_t chunk = (uint32_t)(item % ICP_CHUNKS_PER_SEG);
        const uint32_t base = pr.x * BALL_SEG + chunk * ICP_CHUNK;
        const uint32_t p1 = (uint32_t)min((unsigned long long)pt_end, (unsigned long long)(pr.x + 1u) * BALL_SEG);
        if (base >= p1 || base + ICP_CHUNK <= pt_begin) continue;
        const uint32_t h = pr.y;
        const float4 r0 = __ldg(&T[3 * h]), r1 = __ldg(&T[3 * h + 1]), r2 = __ldg(&T[3 * h + 2]);
        float4 v[ICP_P];
        uint32_t idx[ICP_P];
#pragma unroll
        for (int k = 0; k < ICP_P; ++k) {
            const uint32_t i = base + k * 32 + lane;
            idx[k] = i;
            const float nanv = __int_as_float(0x7fc00000);
            v[k] = (i >= pt_begin && i < p1) ? scene.pos[i] : make_float4(nanv, nanv, nanv, 0.f);
        }
        long long acc[ICP_NSUM];
#pragma unroll
        for (int s = 0; s < ICP_NSUM; ++s) acc[s] = 0;
        bool any = false;
#pragma unroll
        for (int k = 0; k < ICP_P; ++k) {
            float x, y, z;
            float4 mp;
            uint32_t lin;
            if (icp_point_test<FUSED>(model, r0, r1, r2, v[k], sq_thres, x, y, z, mp, lin)) {
                any = true;
                double s[3] = {(double)x - (double)cx, (double)y - (double)cy, (double)z - (double)cz};
                double m[3] = {(double)mp.x - (double)cx, (double)mp.y - (double)cy, (double)mp.z - (double)cz};
                acc[0] += 1;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    acc[1 + a] += __double2ll_rn(s[a] * fix_scale);
                    acc[4 + a] += __double2ll_rn(m[a] * fix_scale);
#pragma unroll
                    for (int b = 0; b < 3; ++b)
                        acc[7 + 3 * a + b] += __double2ll_rn(s[a] * m[b] * fix_scale);
                }
                // score term |ref . ref_n|
                uint32_t fl = __float_as_uint(v[k].w);
                bool use_t = (fl & FLAG_TANGENT) != 0u;
                f3 ref = mk3(use_t ? scene.tgt[idx[k]] : scene.nrm[idx[k]]);
                uint32_t mi = model.voxel[lin];
                f3 rn = mk3(use_t ? model.cloud.tgt[mi] : model.cloud.nrm[mi]);
                f3 rr = {row_rot(r0, ref), row_rot(r1, ref), row_rot(r2, ref)};
                acc[16] += (long long)score_fixed(fabsf(dot3(rr, rn)));
            }
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int s = 0; s < ICP_NSUM; ++s) {
                long long t = warp_sum_i64(acc[s]);
                if (lane == 0 && t)
                    atomicAdd((unsigned long long*)&sums[(size_t)h * ICP_NSUM + s], (unsigned long long)t);
            }
        }
    }
}
// pairs: room for (segments of [pt_begin, pt_end)) x n_hyp entries; n_pairs: one device counter (zeroed here)
void launch_icp_accumulate(cudaStream_t st, const CloudDev& scene, const ModelDev& model,
                           const float4* T, const uint32_t* active, uint32_t n_hyp,
                           uint32_t pt_begin, uint32_t pt_end, float sq_thres, float cx, float cy,
                           float cz, double fix_scale, long long* sums, uint2* pairs, uint32_t* n_pairs, int grid,
                           bool fused) {
    if (!n_hyp || pt_end <= pt_begin) return;
    g_launch_count += 2;
    cudaMemsetAsync(n_pairs, 0, 4, st);
    icp_cull_kernel<<<grid, 256, 0, st>>>(scene, model, T, active, 0u, n_hyp, pt_begin, pt_end, pairs, n_pairs);
    if (fused)
        icp_accumulate_kernel<true><<<grid, 256, 0, st>>>(scene, model, T, pairs, n_pairs, pt_begin, pt_end, sq_thres,
                                                          cx, cy, cz, fix_scale, sums);
    else
        icp_accumulate_kernel<false><<<grid, 256, 0, st>>>(scene, model, T, pairs, n_pairs, pt_begin, pt_end, sq_thres,
                                                           cx, cy, cz, fix_scale, sums);
}
size_t icp_pairs_bytes(uint32_t pt_begin, uint32_t pt_end, uint32_t n_hyp) {
    if (pt_end <= pt_begin) return 8;
    const size_t segs = (size_t)((pt_end - 1u) / BALL_SEG) - (size_t)(pt_begin / BALL_SEG) + 1;
    return segs * (size_t)std::max(n_hyp, 1u) * sizeof(uint2) + 8;
}

// ------------------------------------------------------------- rigid solve
__device__ void svd3_jacobi(const double A[3][3], double U[3][3], double S[3], double V[3][3]) {
    double B[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            B[i][j] = A[i][j];
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    // long-latency FP64 operations are what this single-thread routine spends its time on (it sits on the
    // dependent chain of every ICP iteration): the convergence test compares squares (no sqrt, no division) and
    // the rotation uses one rsqrt.
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool converged = true;  // every pair: |g| <= 1e-15 * sqrt(a * b)
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < 3; ++i) {
                    a += B[i][p] * B[i][p];
                    b += B[i][q] * B[i][q];
                    g += B[i][p] * B[i][q];
                }
                if (g * g > 1e-30 * (a * b)) converged = false;
                if (fabs(g) < 1e-300) continue;
                double zeta = (b - a) / (2.0 * g);
                double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = rsqrt(1.0 + tt * tt), sn = cs * tt;
                for (int i = 0; i < 3; ++i) {
                    double bp = B[i][p], bq = B[i][q];
                    B[i][p] = cs * bp - sn * bq;
                    B[i][q] = sn * bp + cs * bq;
                    double vp = V[i][p], vq = V[i][q];
                    V[i][p] = cs * vp - sn * vq;
                    V[i][q] = sn * vp + cs * vq;
                }
            }
        if (converged) break;
    }
    double Sn[3];
    for (int j = 0; j < 3; ++j)
        Sn[j] = sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
    int o0 = 0, o1 = 1, o2 = 2, t;  // sort descending
    if (Sn[o0] < Sn[o1]) { t = o0; o0 = o1; o1 = t; }
    if (Sn[o1] < Sn[o2]) { t = o1; o1 = o2; o2 = t; }
    if (Sn[o0] < Sn[o1]) { t = o0; o0 = o1; o1 = t; }
    int ord[3] = {o0, o1, o2};
    double Bs[3][3], Vs[3][3];
    for (int j = 0; j < 3; ++j) {
        S[j] = Sn[ord[j]];
        for (int i = 0; i < 3; ++i) {
            Bs[i][j] = B[i][ord[j]];
            Vs[i][j] = V[i][ord[j]];
        }
    }
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            V[i][j] = Vs[i][j];
            U[i][j] = S[j] > 1e-300 ? Bs[i][j] / S[j] : 0.0;
        }
    if (S[2] <= 1e-12 * S[0]) {  // rank-deficient: complete U with col0 x col1
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
}
__device__ double det3d(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
           M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// Eigen::umeyama(src = scene, dst = model, with_scaling = false) on the sums,
// composed with the transform the sums were taken under: T_new = dT * T_best.
__device__ void umeyama_from_sums(const long long* s, double inv_scale, double cx, double cy,
                                  double cz, const float4* Tb, float4* Tn) {
    double n = (double)s[0];
    double ms[3], mm[3], c[3] = {cx, cy, cz};
    for (int a = 0; a < 3; ++a) {
        ms[a] = (double)s[1 + a] * inv_scale / n;
        mm[a] = (double)s[4 + a] * inv_scale / n;
    }
    double sigma[3][3];  // sigma[a][b] = E[(m-mm)_a (s-ms)_b]
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            sigma[a][b] = (double)s[7 + 3 * b + a] * inv_scale / n - mm[a] * ms[b];
    double U[3][3], S[3], V[3][3];
    svd3_jacobi(sigma, U, S, V);
    double sg = det3d(U) * det3d(V) < 0 ? -1.0 : 1.0;
    double R[3][3], dt[3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            R[a][b] = U[a][0] * V[b][0] + U[a][1] * V[b][1] + sg * U[a][2] * V[b][2];
    for (int a = 0; a < 3; ++a) {
        // x -> R (x - c - ms) + mm + c
        double rc = 0;
        for (int b = 0; b < 3; ++b) rc += R[a][b] * (ms[b] + c[b]);
        dt[a] = mm[a] + c[a] - rc;
    }
    double Rb[3][4] = {{Tb[0].x, Tb[0].y, Tb[0].z, Tb[0].w},
                       {Tb[1].x, Tb[1].y, Tb[1].z, Tb[1].w},
                       {Tb[2].x, Tb[2].y, Tb[2].z, Tb[2].w}};
    float o[3][4];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 4; ++b) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += R[a][k] * Rb[k][b];
            if (b == 3) acc += dt[a];
            o[a][b] = (float)acc;
        }
        Tn[a] = make_float4(o[a][0], o[a][1], o[a][2], o[a][3]);
    }
}

// state per transform: Tcur (candidate being measured), Tbest, sums_cur, sums_best,
// n_best, iters, active.  `first` != 0 for the step after the initial accumulate.
__global__ void icp_step_kernel(IcpState st, uint32_t n_hyp, int first, uint32_t max_iterations,
                                double inv_scale, float cx, float cy, float cz) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hyp) return;
    if (!st.active[h]) return;
    long long* cur = st.sums_cur + (size_t)h * ICP_NSUM;
    long long* best = st.sums_best + (size_t)h * ICP_NSUM;
    uint32_t n_cur = (uint32_t)cur[0];
    bool done = false;
    if (first) {
        for (int s = 0; s < ICP_NSUM; ++s) best[s] = cur[s];
        for (int r = 0; r < 3; ++r) st.Tbest[3 * h + r] = st.Tcur[3 * h + r];
        st.iters[h] = 0;
    } else {
        if (n_cur < (uint32_t)best[0]) {  // scene.hpp:396-398
            done = true;
        } else {
            for (int s = 0; s < ICP_NSUM; ++s) best[s] = cur[s];
            for (int r = 0; r < 3; ++r) st.Tbest[3 * h + r] = st.Tcur[3 * h + r];
            st.iters[h] += 1;
            if (st.iters[h] == max_iterations) done = true;  // :400-402
        }
    }
    if (!done && (uint32_t)best[0] < 3u) done = true;  // :381-383
    if (done) {
        st.active[h] = 0;
    } else {
        float4 tn[3];
        umeyama_from_sums(best, inv_scale, cx, cy, cz, st.Tbest + 3 * h, tn);
        for (int r = 0; r < 3; ++r) st.Tcur[3 * h + r] = tn[r];
    }
    for (int s = 0; s < ICP_NSUM; ++s) cur[s] = 0;
}
void launch_icp_step(cudaStream_t stream, const IcpState& st, uint32_t n_hyp, int first,
                     uint32_t max_iterations, double inv_scale, float cx, float cy, float cz) {
    if (!n_hyp) return;
    ++g_launch_count;
    icp_step_kernel<<<(n_hyp + 63) / 64, 64, 0, stream>>>(st, n_hyp, first, max_iterations,
                                                          inv_scale, cx, cy, cz);
}

// --------------------------------------------------------- correspondences
// Two passes over the scene with one warp per CORR_SEG-point segment (ascending
// order, like the ball subsets): count -> scan -> fill.
template <bool FILL, bool FUSED>
__global__ void __launch_bounds__(256)
    corr_kernel(CloudDev scene, ModelDev model, const float4* __restrict__ Trows, float sq_thres, uint32_t n_seg,
                uint32_t* __restrict__ counts, const uint32_t* __restrict__ seg_off,
                uint32_t* __restrict__ scene_corrs, uint32_t* __restrict__ model_corrs,
                unsigned long long* __restrict__ score) {
    const int lane = threadIdx.x & 31;
    const uint32_t seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= n_seg) return;
    const uint32_t t = blockIdx.y;  // transform of the batch; counts / seg_off are [transform][segment]
    Rows T;
    T.r0 = Trows[3 * t]; T.r1 = Trows[3 * t + 1]; T.r2 = Trows[3 * t + 2];
    const size_t slot = (size_t)t * n_seg + seg;
    uint32_t base = seg * CORR_SEG, cnt = 0;
    uint32_t off = FILL ? seg_off[slot] : 0u;
    unsigned long long sc = 0;
    for (uint32_t k = 0; k < CORR_SEG / 32; ++k) {
        uint32_t i = base + k * 32 + lane;
        bool inl = false;
        uint32_t lin = 0;
        if (i < scene.n) {
            float x, y, z;
            float4 mp;
            float4 v = scene.pos[i];
            inl = icp_point_test<FUSED>(model, T.r0, T.r1, T.r2, v, sq_thres, x, y, z, mp, lin);
            if (inl && !FILL) {
                uint32_t fl = __float_as_uint(v.w);
                bool use_t = (fl & FLAG_TANGENT) != 0u;
                f3 ref = mk3(use_t ? scene.tgt[i] : scene.nrm[i]);
                uint32_t mi = model.voxel[lin];
                f3 rn = mk3(use_t ? model.cloud.tgt[mi] : model.cloud.nrm[mi]);
                f3 rr = {row_rot(T.r0, ref), row_rot(T.r1, ref), row_rot(T.r2, ref)};
                sc += score_fixed(fabsf(dot3(rr, rn)));
            }
        }
        uint32_t b = __ballot_sync(0xffffffffu, inl);
        if (FILL && inl) {
            uint32_t o = off + cnt + __popc(b & ((1u << lane) - 1u));
            scene_corrs[o] = i;
            model_corrs[o] = model.voxel[lin];
        }
        cnt += __popc(b);
    }
    if (!FILL) {
        if (lane == 0) counts[slot] = cnt;
#pragma unroll
        for (int d = 16; d; d >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, d);
        if (lane == 0 && sc) atomicAdd(&score[t], sc);
    }
}
void launch_corr_count(cudaStream_t st, const CloudDev& scene, const ModelDev& model, const float4* Trows, uint32_t n_T,
                       float sq_thres, uint32_t n_seg, uint32_t* counts, unsigned long long* score, bool fused) {
    if (!n_seg || !n_T) return;
    ++g_launch_count;
    const dim3 grid((n_seg + 7) / 8, n_T);
    if (fused)
        corr_kernel<false, true><<<grid, 256, 0, st>>>(scene, model, Trows, sq_thres, n_seg, counts, nullptr, nullptr,
                                                       nullptr, score);
    else
        corr_kernel<false, false><<<grid, 256, 0, st>>>(scene, model, Trows, sq_thres, n_seg, counts, nullptr, nullptr,
                                                        nullptr, score);
}
void launch_corr_fill(cudaStream_t st, const CloudDev& scene, const ModelDev& model, const float4* Trows, uint32_t n_T,
                      float sq_thres, uint32_t n_seg, const uint32_t* seg_off, uint32_t* scene_corrs,
                      uint32_t* model_corrs, bool fused) {
    if (!n_seg || !n_T) return;
    ++g_launch_count;
    const dim3 grid((n_seg + 7) / 8, n_T);
    if (fused)
        corr_kernel<true, true><<<grid, 256, 0, st>>>(scene, model, Trows, sq_thres, n_seg, nullptr, seg_off, scene_corrs,
                                                      model_corrs, nullptr);
    else
        corr_kernel<true, false><<<grid, 256, 0, st>>>(scene, model, Trows, sq_thres, n_seg, nullptr, seg_off, scene_corrs,
                                                       model_corrs, nullptr);
}

}  // namespace tmk
