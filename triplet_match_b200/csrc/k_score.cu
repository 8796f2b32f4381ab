// k_score.cu — stage 3b, the hot loop: per-hypothesis inlier scoring
// (scene::impl::project_, include/impl/scene.hpp:411-510) against the model's
// nearest-neighbour voxel grid (model::voxel_query, include/impl/model.hpp:180-192).
//
//   score_full_kernel        finish_find semantics (early_out = false): every
//                            hypothesis tests its whole subset.  Register-tiled:
//                            a thread keeps P scene points in registers and
//                            walks the hypotheses of its outer sample from
//                            shared memory, so scene bytes are read once per
//                            (tile, hypothesis chunk) instead of once per test.
//   score_early_drop_kernel  project_ with early_out = true, sequentially
//                            faithful: one warp walks one hypothesis' subset in
//                            order and resolves the 18 early-drop checkpoints
//                            (scene.hpp:492-506) with ballots.
//   build_work_kernel(s)     device-side work list for the persistent scorer.
//   argmax_kernel            packed (inliers << 32 | ~id) max — the local half
//                            of the best-pose all-reduce.
#include "tm_kernels.cuh"

namespace tmk {

// one hypothesis-point test; returns true for an inlier.  mi_out = model index.
template <bool FUSED>
__device__ __forceinline__ bool point_test(const ModelDev& m, float4 r0, float4 r1, float4 r2,
                                           float px, float py, float pz, uint32_t pflags,
                                           float sq_thres, float& x, float& y, float& z,
                                           uint32_t& lin_out) {
    x = row_apply(r0, px, py, pz);  // pos = t * pos  (scene.hpp:444)
    y = row_apply(r1, px, py, pz);
    z = row_apply(r2, px, py, pz);
    // voxel_query (model.hpp:182): ijk = trunc(to_voxel * pos); diag + translation
    float vx = m.sx * x + m.tx, vy = m.sy * y + m.ty, vz = m.sz * z + m.tz;
    // trunc-toward-zero bounds in the float domain: int(v) in [0, e) <=> -1 < v < e
    bool inb = (vx > -1.f) & (vx < m.exf) & (vy > -1.f) & (vy < m.eyf) & (vz > -1.f) & (vz < m.ezf);
    if (!inb) return false;
    int i = (int)vx, j = (int)vy, k = (int)vz;
    uint32_t lin = (uint32_t)((k * m.ey + j) * m.ex + i);
    lin_out = lin;
    float4 mp;
    if (FUSED) {
        mp = __ldg(&m.vcell[lin]);
    } else {
        uint32_t mi = __ldg(&m.voxel[lin]);
        mp = __ldg(&m.cloud.pos[mi]);
    }
    float dx = x - mp.x, dy = y - mp.y, dz = z - mp.z;
    float sq = sum3(dx * dx, dy * dy, dz * dz);
    // dist = sqrt(sq); `dist > thres` (scene.hpp:465) <=> sq > sq_thres (host-computed
    // exact boundary of the correctly rounded sqrt)
    if (sq > sq_thres) return false;
    // tangent-class agreement (scene.hpp:469-478)
    return ((pflags ^ __float_as_uint(mp.w)) & FLAG_TANGENT) == 0u;
}

// |ref . ref_n| of an inlier (scene.hpp:461,479-483) as 2^-36 fixed point
__device__ __forceinline__ unsigned long long inlier_score(const CloudDev& scene, const ModelDev& m,
                                                           float4 r0, float4 r1, float4 r2,
                                                           uint32_t sidx, uint32_t pflags,
                                                           uint32_t lin) {
    bool use_t = (pflags & FLAG_TANGENT) != 0u;
    f3 ref = mk3(use_t ? scene.tgt[sidx] : scene.nrm[sidx]);
    uint32_t mi = m.voxel[lin];
    f3 rn = mk3(use_t ? m.cloud.tgt[mi] : m.cloud.nrm[mi]);  // classes agree for an inlier
    f3 rr = {row_rot(r0, ref), row_rot(r1, ref), row_rot(r2, ref)};
    return score_fixed(fabsf(dot3(rr, rn)));
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// --------------------------------------------------------------- full scoring
template <int P, bool FUSED, bool WITH_SCORE>
__global__ void __launch_bounds__(SCORE_THREADS)
    score_full_kernel(ScoreArgs a) {
    __shared__ float4 sT[SCORE_HSTAGE * 3];
    __shared__ uint32_t s_item;
    const int lane = threadIdx.x & 31;
    const uint32_t n_items = *a.n_items;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const WorkItem w = a.items[item];
        float px[P], py[P], pz[P];
        uint32_t pfl[P], pidx[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            uint32_t q = k * SCORE_THREADS + threadIdx.x;
            const float nanv = __int_as_float(0x7fc00000);
            px[k] = py[k] = pz[k] = nanv;  // a NaN point fails every test
            pfl[k] = 0u;
            pidx[k] = 0u;
            if (q < w.npts) {
                uint32_t idx = a.sub_idx ? (uint32_t)a.sub_idx[w.sub_begin + q]
                                         : (uint32_t)(w.sub_begin + q);
                float4 v = a.scene.pos[idx];
                uint32_t fl = __float_as_uint(v.w);
                pidx[k] = idx;
                pfl[k] = fl;
                if (!(fl & FLAG_MASKED)) {  // mask_ (scene.hpp:434)
                    px[k] = v.x;
                    py[k] = v.y;
                    pz[k] = v.z;
                }
            }
        }
        for (uint32_t h0 = w.hyp_begin; h0 < w.hyp_end; h0 += SCORE_HSTAGE) {
            const uint32_t nh = min((uint32_t)SCORE_HSTAGE, w.hyp_end - h0);
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < nh * 3; t += SCORE_THREADS)
                sT[t] = a.T[(size_t)h0 * 3 + t];
            __syncthreads();
            uint32_t mycnt = 0;
            unsigned long long mysc = 0;
            for (uint32_t hh = 0; hh < nh; ++hh) {
                const float4 r0 = sT[3 * hh], r1 = sT[3 * hh + 1], r2 = sT[3 * hh + 2];
                uint32_t c = 0;
                unsigned long long sc = 0;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    float x, y, z;
                    uint32_t lin;
                    if (point_test<FUSED>(a.model, r0, r1, r2, px[k], py[k], pz[k], pfl[k],
                                          a.sq_thres, x, y, z, lin)) {
                        ++c;
                        if (WITH_SCORE)
                            sc += inlier_score(a.scene, a.model, r0, r1, r2, pidx[k], pfl[k], lin);
                    }
                }
                const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
                if (tot) {  // warp-uniform
                    unsigned long long s = WITH_SCORE ? warp_sum_u64(sc) : 0ull;
                    if (lane == (int)(hh & 31u)) {
                        mycnt = tot;
                        mysc = s;
                    }
                }
                if ((hh & 31u) == 31u || hh == nh - 1) {
                    if (mycnt) {
                        uint32_t h = h0 + (hh & ~31u) + lane;
                        atomicAdd(&a.counts[h], mycnt);
                        if (WITH_SCORE) atomicAdd(&a.scores[h], mysc);
                    }
                    mycnt = 0;
                    mysc = 0;
                }
            }
        }
    }
}

template <int P>
static void launch_score_full_p(cudaStream_t st, const ScoreArgs& a, int grid, bool fused,
                                bool with_score) {
    if (fused) {
        if (with_score) score_full_kernel<P, true, true><<<grid, SCORE_THREADS, 0, st>>>(a);
        else score_full_kernel<P, true, false><<<grid, SCORE_THREADS, 0, st>>>(a);
    } else {
        if (with_score) score_full_kernel<P, false, true><<<grid, SCORE_THREADS, 0, st>>>(a);
        else score_full_kernel<P, false, false><<<grid, SCORE_THREADS, 0, st>>>(a);
    }
}
void launch_score_full(cudaStream_t st, const ScoreArgs& a, int grid, bool fused, bool with_score) {
    ++g_launch_count;
    launch_score_full_p<SCORE_P>(st, a, grid, fused, with_score);
}
int score_full_max_blocks_per_sm(bool fused, bool with_score) {
    int nb = 0;
    if (fused) {
        if (with_score)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &nb, score_full_kernel<SCORE_P, true, true>, SCORE_THREADS, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &nb, score_full_kernel<SCORE_P, true, false>, SCORE_THREADS, 0);
    } else {
        if (with_score)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &nb, score_full_kernel<SCORE_P, false, true>, SCORE_THREADS, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &nb, score_full_kernel<SCORE_P, false, false>, SCORE_THREADS, 0);
    }
    return nb > 0 ? nb : 1;
}

// ------------------------------------------------------------------ work list
// groups: per outer sample g: subset CSR row [sub_off[g], sub_off[g+1]) and the
// local hypothesis range [g_hyp[g], g_hyp[g+1]) (already clipped to the shard).
__global__ void work_count_kernel(const unsigned long long* __restrict__ sub_off,
                                  const uint32_t* __restrict__ g_hyp, uint32_t n_groups,
                                  uint32_t* __restrict__ n_items_g,
                                  unsigned long long* __restrict__ n_tests) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    unsigned long long np = sub_off[g + 1] - sub_off[g];
    uint32_t nh = g_hyp[g + 1] - g_hyp[g];
    unsigned long long tiles = (np + SCORE_TILE - 1) / SCORE_TILE;
    uint32_t chunks = (nh + SCORE_HCHUNK - 1) / SCORE_HCHUNK;
    n_items_g[g] = (uint32_t)(tiles * chunks);
    if (n_tests && np && nh) atomicAdd(n_tests, np * nh);
}
__global__ void work_fill_kernel(const unsigned long long* __restrict__ sub_off,
                                 const uint32_t* __restrict__ g_hyp, uint32_t n_groups,
                                 const uint32_t* __restrict__ item_off,
                                 WorkItem* __restrict__ items) {
    uint32_t g = blockIdx.x;
    if (g >= n_groups) return;
    unsigned long long sb = sub_off[g], np = sub_off[g + 1] - sb;
    uint32_t hb = g_hyp[g], nh = g_hyp[g + 1] - hb;
    uint32_t tiles = (uint32_t)((np + SCORE_TILE - 1) / SCORE_TILE);
    uint32_t chunks = (nh + SCORE_HCHUNK - 1) / SCORE_HCHUNK;
    uint32_t base = item_off[g];
    for (uint32_t t = threadIdx.x; t < tiles * chunks; t += blockDim.x) {
        uint32_t tile = t % tiles, chunk = t / tiles;
        WorkItem w;
        w.sub_begin = sb + (unsigned long long)tile * SCORE_TILE;
        w.npts = (uint32_t)min((unsigned long long)SCORE_TILE, np - (unsigned long long)tile * SCORE_TILE);
        w.hyp_begin = hb + chunk * SCORE_HCHUNK;
        w.hyp_end = min(hb + nh, w.hyp_begin + SCORE_HCHUNK);
        items[base + t] = w;
    }
}
void launch_work_count(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* g_hyp,
                       uint32_t n_groups, uint32_t* n_items_g, unsigned long long* n_tests) {
    if (!n_groups) return;
    ++g_launch_count;
    work_count_kernel<<<(n_groups + 127) / 128, 128, 0, st>>>(sub_off, g_hyp, n_groups, n_items_g,
                                                              n_tests);
}
void launch_work_fill(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* g_hyp,
                      uint32_t n_groups, const uint32_t* item_off, WorkItem* items) {
    if (!n_groups) return;
    ++g_launch_count;
    work_fill_kernel<<<n_groups, 128, 0, st>>>(sub_off, g_hyp, n_groups, item_off, items);
}

// ------------------------------------------------------------------ early drop
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

// One warp per hypothesis, subset walked in order 32 elements at a time.
// g_of_hyp: subset row of each hypothesis (null => row 0); sub_idx null => identity.
template <bool FUSED>
__global__ void __launch_bounds__(256)
    score_early_drop_kernel(EarlyArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= (a.n_hyp_dev ? *a.n_hyp_dev : a.n_hyp)) return;
    const uint32_t g = a.g_of_hyp ? a.g_of_hyp[h] : 0u;
    const unsigned long long sb = a.sub_off[g];
    const uint32_t nsub = (uint32_t)(a.sub_off[g + 1] - sb);
    const float4 r0 = a.T[3 * (size_t)h], r1 = a.T[3 * (size_t)h + 1], r2 = a.T[3 * (size_t)h + 2];
    // tests[i] = step_size * (i+1) * |subset| (scene.hpp:422-426), float arithmetic
    uint32_t next_test = 0;
    const float accept_bound = a.accept_prob * (float)a.model.cloud.n;  // scene.hpp:500
    uint32_t corrs = 0;
    unsigned long long lane_score = 0;  // per-lane running sum
    bool dropped = false;
    uint32_t drop_corrs = 0, drop_tried = 0;
    unsigned long long drop_score = 0;
    auto test_at = [&](uint32_t i) -> uint32_t {
        return (uint32_t)(0.05f * (float)(i + 1u) * (float)nsub);
    };
    uint32_t cur_test = test_at(0);
    for (uint32_t base = 0; base < nsub && !dropped; base += 32) {
        uint32_t q = base + lane;
        bool reach = false, inl = false;
        unsigned long long term = 0;
        if (q < nsub) {
            uint32_t idx = a.sub_idx ? (uint32_t)a.sub_idx[sb + q] : (uint32_t)(sb + q);
            float4 v = a.scene.pos[idx];
            uint32_t fl = __float_as_uint(v.w);
            if (!(fl & FLAG_MASKED)) {
                float x, y, z;
                uint32_t lin = 0xffffffffu;
                inl = point_test<FUSED>(a.model, r0, r1, r2, v.x, v.y, v.z, fl, a.sq_thres, x, y, z,
                                        lin);
                reach = lin != 0xffffffffu;  // voxel_query succeeded (scene.hpp:458-460)
                if (inl) term = inlier_score(a.scene, a.model, r0, r1, r2, idx, fl, lin);
            }
        }
        const uint32_t reach_mask = __ballot_sync(0xffffffffu, reach);
        const uint32_t inl_mask = __ballot_sync(0xffffffffu, inl);
        if (a.early_out && next_test < 18u && base + 32u >= cur_test) {
            int last_l = -1;
            while (next_test < 18u) {
                // first reaching element at or after position cur_test, one checkpoint per element
                int lmin = (cur_test > base + 1u) ? (int)(cur_test - base - 1u) : 0;
                int lstart = max(lmin, last_l + 1);
                if (lstart >= 32) break;
                uint32_t cand = reach_mask & (0xffffffffu << lstart);
                if (!cand) break;
                int l = __ffs(cand) - 1;
                uint32_t tried = base + (uint32_t)l + 1u;
                uint32_t c_here = corrs + __popc(inl_mask & (0xffffffffu >> (31 - l)));
                uint32_t upper = early_drop_upper(tried, nsub, c_here);
                if ((float)upper < accept_bound) {
                    dropped = true;
                    drop_corrs = c_here;
                    drop_tried = tried;
                    unsigned long long part = lane_score + ((lane <= l) ? term : 0ull);
                    drop_score = warp_sum_u64(part);
                    break;
                }
                ++next_test;
                last_l = l;
                cur_test = next_test < 18u ? test_at(next_test) : 0u;
            }
        }
        corrs += __popc(inl_mask);
        lane_score += term;
    }
    unsigned long long total = dropped ? drop_score : warp_sum_u64(lane_score);
    if (lane == 0) {
        a.counts[h] = dropped ? drop_corrs : corrs;
        if (a.scores) a.scores[h] = total;
        if (a.dropped) a.dropped[h] = dropped ? 1 : 0;
        if (a.tested) a.tested[h] = dropped ? drop_tried : nsub;
        if (a.n_tests) atomicAdd(a.n_tests, (unsigned long long)(dropped ? drop_tried : nsub));
    }
}
void launch_score_early_drop(cudaStream_t st, const EarlyArgs& a, bool fused) {
    if (!a.n_hyp) return;
    ++g_launch_count;
    unsigned grid = (a.n_hyp + 7) / 8;
    if (fused) score_early_drop_kernel<true><<<grid, 256, 0, st>>>(a);
    else score_early_drop_kernel<false><<<grid, 256, 0, st>>>(a);
}

// ---------------------------------------------------------------------- argmax
// key = (inliers << 32) | (0xFFFFFFFF - global hypothesis id); ties -> lowest id.
__global__ void __launch_bounds__(256)
    argmax_kernel(const uint32_t* __restrict__ counts, const uint8_t* __restrict__ valid,
                  const uint32_t* __restrict__ n_local, const unsigned long long* __restrict__ h_begin,
                  unsigned long long* __restrict__ best) {
    __shared__ unsigned long long wbest[8];
    const uint32_t n = *n_local;
    const unsigned long long hb = *h_begin;
    unsigned long long k = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        uint32_t c = counts[i];
        if (!c) continue;
        unsigned long long key =
            ((unsigned long long)c << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)(hb + i));
        k = key > k ? key : k;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, k, d);
        k = o > k ? o : k;
    }
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) k = wbest[w] > k ? wbest[w] : k;
        if (k) atomicMax(best, k);
    }
}
void launch_argmax(cudaStream_t st, const uint32_t* counts, const uint8_t* valid,
                   const uint32_t* n_local, const unsigned long long* h_begin,
                   unsigned long long* best, int grid) {
    ++g_launch_count;
    argmax_kernel<<<grid, 256, 0, st>>>(counts, valid, n_local, h_begin, best);
}

}  // namespace tmk
