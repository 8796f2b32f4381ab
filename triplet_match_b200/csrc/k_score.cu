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
#include <cstdlib>

#include <algorithm>

#include "tm_kernels.cuh"

namespace tmk {

// one hypothesis-point test; returns true for an inlier.  mi_out = model index.
template <bool FUSED, bool OCC = false>
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
    lin_out = lin;  // voxel_query succeeded ("reaching" element), whatever the mask says
    if (OCC && !occ_test(m, i, j, k)) return false;  // provably farther than the threshold
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
    f3 rn = mk3(m.mref[mi]);  // classes agree for an inlier: tangent or normal of the model point
    f3 rr = {row_rot(r0, ref), row_rot(r1, ref), row_rot(r2, ref)};
    return score_fixed(fabsf(dot3(rr, rn)));
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// --------------------------------------------------------------- full scoring
// v4 (round 1).  Every WARP is an independent worker: it pulls (tile, hypothesis
// chunk) work items from a global counter — no CTA barrier anywhere — where a
// tile is SCORE_P*32 CONTIGUOUS subset points (P per lane, kept in registers).
// The hypotheses of an item are walked 32 at a time in two phases:
//   cull   lane l loads the rows of hypothesis h0+l (and parks them in a per-warp
//          shared-memory slab) and bounds it: the tile's bounding box is pushed
//          through the transform with interval arithmetic (|R| * half-extent, plus a
//          1e-5 relative guard that dominates every float rounding of the exact path);
//          if the resulting voxel-coordinate box misses (-1, extent) on any axis no
//          point of the tile can pass voxel_query and the pair is skipped.  NaN
//          bounds never cull.  A ballot gives the surviving hypotheses.
//   test   for each survivor the rows come back from shared memory (broadcast) and
//          the exact reference arithmetic runs on the P points of every lane: all P
//          voxel coordinates first, then the P cell gathers back to back (P loads
//          in flight per lane), then distance / class tests.  Inliers are counted
//          with REDUX and flushed with one coalesced RED per 32 hypotheses.
// The per-inlier score term needs the scene point's ref vector (tangent or normal),
// staged once per item in a per-warp shared-memory slab, and the model point's ref
// vector, which comes from the compact per-model-point array mref (its index rides in the fused cell's .w).
template <int P, bool FUSED, bool WITH_SCORE, bool CULL, bool OCC>
__global__ void __launch_bounds__(SCORE_THREADS, SCORE_MIN_BLOCKS)
    score_full_kernel(ScoreArgs a) {
    __shared__ float4 s_ref[WITH_SCORE ? SCORE_THREADS * P : 1];
    __shared__ float4 s_rows[CULL ? (SCORE_THREADS / 32) * 32 * 3 : 1];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // thread-private slots (each lane re-reads only what it staged): slot k of thread t at k * SCORE_THREADS + t,
    // so the address is tid * 16 + a constant and no lane ever touches another lane's slot (no barrier needed)
    float4* my_ref = s_ref + (WITH_SCORE ? threadIdx.x : 0);
    float4* my_rows = s_rows + (CULL ? warp * 32 * 3 : 0);
    const uint32_t n_items = *a.n_items;
    const ModelDev& m = a.model;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const WorkItem w = a.items[item];
        float px[P], py[P], pz[P];
        uint32_t tflags = 0;  // bit k: tangent_mask_ of point k
        const float nanv = __int_as_float(0x7fc00000);
        float mnx = 3.0e38f, mny = 3.0e38f, mnz = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f, mxz = -3.0e38f;
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const uint32_t q = k * 32 + lane;
            px[k] = py[k] = pz[k] = nanv;  // a NaN point fails every test
            if (q < w.npts) {
                uint32_t idx = a.sub_idx ? (uint32_t)a.sub_idx[w.sub_begin + q]
                                         : (uint32_t)(w.sub_begin + q);
                float4 v = a.scene.pos[idx];
                uint32_t fl = __float_as_uint(v.w);
                if (!(fl & FLAG_MASKED)) {  // mask_ (scene.hpp:434)
                    px[k] = v.x; py[k] = v.y; pz[k] = v.z;
                    if (fl & FLAG_TANGENT) tflags |= 1u << k;
                    mnx = fminf(mnx, v.x); mxx = fmaxf(mxx, v.x);
                    mny = fminf(mny, v.y); mxy = fmaxf(mxy, v.y);
                    mnz = fminf(mnz, v.z); mxz = fmaxf(mxz, v.z);
                    if (WITH_SCORE)  // ref = use_tangent ? tangent : normal (scene.hpp:441-442)
                        my_ref[k * SCORE_THREADS] = (fl & FLAG_TANGENT) ? a.scene.tgt[idx] : a.scene.nrm[idx];
                }
            }
        }
        float cx = 0.f, cy = 0.f, cz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
        if (CULL) {
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
                mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
                mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, d));
                mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
                mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
                mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, d));
            }
            if (!(mnx <= mxx)) continue;  // no live (finite, unmasked) point in this tile
            cx = 0.5f * (mnx + mxx); hx = 0.5f * (mxx - mnx);
            cy = 0.5f * (mny + mxy); hy = 0.5f * (mxy - mny);
            cz = 0.5f * (mnz + mxz); hz = 0.5f * (mxz - mnz);
        }
        for (uint32_t h0 = w.hyp_begin; h0 < w.hyp_end; h0 += 32) {
            const uint32_t h = h0 + lane;
            uint32_t mask;
            if (CULL) {
                bool survive = false;
                __syncwarp();  // readers of the previous batch's rows are done
                if (h < w.hyp_end) {
                    const float4 r0 = __ldg(&a.T[3 * (size_t)h]), r1 = __ldg(&a.T[3 * (size_t)h + 1]),
                                 r2 = __ldg(&a.T[3 * (size_t)h + 2]);
                    my_rows[lane] = r0;  // row-major slabs: conflict-free stores, broadcast loads
                    my_rows[32 + lane] = r1;
                    my_rows[64 + lane] = r2;
                    const float acx = fabsf(cx) + hx, acy = fabsf(cy) + hy, acz = fabsf(cz) + hz;
                    bool out = false;
#define TM_AXIS(r, S, TV, EXF)                                                                  \
    {                                                                                           \
        float cc = r.x * cx + r.y * cy + r.z * cz + r.w;                                        \
        float ee = fabsf(r.x) * hx + fabsf(r.y) * hy + fabsf(r.z) * hz;                         \
        float mag = fabsf(r.x) * acx + fabsf(r.y) * acy + fabsf(r.z) * acz + fabsf(r.w);        \
        ee += 1e-5f * mag + 1e-30f;                                                             \
        float sl = 1e-5f * (S * mag + fabsf(TV)) + 1e-30f;                                      \
        float lo = S * (cc - ee) + TV - sl, hi = S * (cc + ee) + TV + sl;                       \
        out = out || (lo >= EXF) || (hi <= -1.0f);                                              \
    }
                    TM_AXIS(r0, m.sx, m.tx, m.exf)
                    TM_AXIS(r1, m.sy, m.ty, m.eyf)
                    TM_AXIS(r2, m.sz, m.tz, m.ezf)
#undef TM_AXIS
                    survive = !out;
                }
                mask = __ballot_sync(0xffffffffu, survive);  // also orders the smem stores
            } else {
                const uint32_t left = w.hyp_end - h0;
                mask = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
            }
            if (a.stats && lane == 0) {
                atomicAdd(&a.stats[0], (unsigned long long)min(32u, w.hyp_end - h0));
                atomicAdd(&a.stats[1], (unsigned long long)__popc(mask));
            }
            uint32_t mycnt = 0;
            unsigned long long mysc = 0;
            while (mask) {
                const int hh = __ffs(mask) - 1;
                mask &= mask - 1u;
                float4 r0, r1, r2;
                if (CULL) {
                    r0 = my_rows[hh]; r1 = my_rows[32 + hh]; r2 = my_rows[64 + hh];
                } else {
                    const size_t hb = 3 * (size_t)(h0 + hh);
                    r0 = __ldg(&a.T[hb]); r1 = __ldg(&a.T[hb + 1]); r2 = __ldg(&a.T[hb + 2]);
                }
                // ---- exact path, stage 1: pos = t*pos, ijk = trunc(to_voxel*pos) (scene.hpp:444,
                // model.hpp:182-189) for all P points
                float x[P], y[P], z[P];
                uint32_t lin[P];
                uint32_t inb = 0;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    x[k] = row_apply(r0, px[k], py[k], pz[k]);
                    y[k] = row_apply(r1, px[k], py[k], pz[k]);
                    z[k] = row_apply(r2, px[k], py[k], pz[k]);
                    const float vx = m.sx * x[k] + m.tx, vy = m.sy * y[k] + m.ty, vz = m.sz * z[k] + m.tz;
                    const int i = (int)vx, j = (int)vy, kk = (int)vz;
                    // the reference's own test, on the truncated integers (model.hpp:186-189): cvt.rzi saturates, so
                    // huge values stay out; a NaN converts to 0 and is rejected by the distance test (sq is NaN)
                    bool in = ((uint32_t)i < (uint32_t)m.ex) & ((uint32_t)j < (uint32_t)m.ey) &
                              ((uint32_t)kk < (uint32_t)m.ez);
                    if (OCC) {  // branch-free: out-of-grid lanes look at block 0
                        const uint32_t b = in ? (uint32_t)(((kk >> OCC_SHIFT) * m.oby + (j >> OCC_SHIFT)) * m.obx +
                                                           (i >> OCC_SHIFT))
                                              : 0u;
                        in = in & (((__ldg(&m.occ[b >> 5]) >> (b & 31u)) & 1u) != 0u);
                    }
                    lin[k] = in ? (uint32_t)((kk * m.ey + j) * m.ex + i) : 0u;
                    inb |= in ? (1u << k) : 0u;
                }
                if (!__any_sync(0xffffffffu, inb != 0u)) continue;  // nothing reaches the grid
                // ---- stage 2: P cell gathers in flight (out-of-grid lanes read cell 0, unused)
                float4 mp[P];
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    if (FUSED) mp[k] = __ldg(&m.vcell[lin[k]]);
                    else mp[k] = __ldg(&m.cloud.pos[__ldg(&m.voxel[lin[k]])]);
                }
                // ---- stage 3: dist > thres (scene.hpp:464-467), class agreement (:469-478)
                uint32_t c = 0;
                unsigned long long sc = 0;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float dx = x[k] - mp[k].x, dy = y[k] - mp[k].y, dz = z[k] - mp[k].z;
                    const float sq = sum3(dx * dx, dy * dy, dz * dz);
                    const uint32_t pfl = (tflags >> k) & 1u;
                    const bool inl = ((inb >> k) & 1u) && (sq <= a.sq_thres) &&  // NaN: not an inlier
                                     (((pfl ^ __float_as_uint(mp[k].w)) & FLAG_TANGENT) == 0u);
                    if (inl) {
                        ++c;
                        if (WITH_SCORE) {
                            const f3 ref = mk3(my_ref[k * SCORE_THREADS]);
                            f3 rn;
                            // the model point's ref vector (its class agrees with the scene point's)
                            const uint32_t mi = FUSED ? (__float_as_uint(mp[k].w) >> 1) : __ldg(&m.voxel[lin[k]]);
                            rn = mk3(__ldg(&m.mref[mi]));
                            const f3 rr = {row_rot(r0, ref), row_rot(r1, ref), row_rot(r2, ref)};
                            sc += score_fixed(fabsf(dot3(rr, rn)));  // scene.hpp:461,483
                        }
                    }
                }
                const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
                if (tot) {  // warp-uniform
                    if (a.stats && lane == 0) atomicAdd(&a.stats[2], 1ull);
                    // per-lane partial < 2^39 for rigid transforms (P terms <= ~2^36): two 32-bit REDUX
                    // (low 24 bits, the rest) instead of a 5-step 64-bit shuffle tree
                    unsigned long long sv = 0ull;
                    if (WITH_SCORE) {
                        const uint32_t lo = __reduce_add_sync(0xffffffffu, (uint32_t)(sc & 0xffffffull));
                        const uint32_t hi = __reduce_add_sync(0xffffffffu, (uint32_t)(sc >> 24));
                        sv = ((unsigned long long)hi << 24) + lo;
                    }
                    if (lane == hh) {
                        mycnt = tot;
                        mysc = sv;
                    }
                }
            }
            if (mycnt) {
                atomicAdd(&a.counts[h], mycnt);
                if (WITH_SCORE) atomicAdd(&a.scores[h], mysc);
            }
        }
    }
}

template <bool FUSED, bool WITH_SCORE, bool CULL>
static void launch_score_full_v(cudaStream_t st, const ScoreArgs& a, int grid) {
    if (a.model.occ) score_full_kernel<SCORE_P, FUSED, WITH_SCORE, CULL, true><<<grid, SCORE_THREADS, 0, st>>>(a);
    else score_full_kernel<SCORE_P, FUSED, WITH_SCORE, CULL, false><<<grid, SCORE_THREADS, 0, st>>>(a);
}
template <bool FUSED, bool WITH_SCORE, bool CULL>
static int occ_score_full_v() {
    int nb = 0, nb2 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &nb, score_full_kernel<SCORE_P, FUSED, WITH_SCORE, CULL, false>, SCORE_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &nb2, score_full_kernel<SCORE_P, FUSED, WITH_SCORE, CULL, true>, SCORE_THREADS, 0);
    nb = nb < nb2 ? nb : nb2;
    return nb > 0 ? nb : 1;
}
#define TM_DISPATCH3(FN, f, s, c, ...)                                                       \
    ((f) ? ((s) ? ((c) ? FN<true, true, true>(__VA_ARGS__) : FN<true, true, false>(__VA_ARGS__))     \
                : ((c) ? FN<true, false, true>(__VA_ARGS__) : FN<true, false, false>(__VA_ARGS__)))  \
         : ((s) ? ((c) ? FN<false, true, true>(__VA_ARGS__) : FN<false, true, false>(__VA_ARGS__))   \
                : ((c) ? FN<false, false, true>(__VA_ARGS__) : FN<false, false, false>(__VA_ARGS__))))
static bool score_cull_enabled() {  // TM_SCORE_CULL=0: development knob, read once
    static const bool v = [] {
        const char* e = getenv("TM_SCORE_CULL");
        return e ? atoi(e) != 0 : true;
    }();
    return v;
}
void launch_score_full(cudaStream_t st, const ScoreArgs& a, int grid, bool fused, bool with_score) {
    TM_DISPATCH3(launch_score_full_v, fused, with_score, score_cull_enabled(), st, a, grid);
}
int score_full_max_blocks_per_sm(bool fused, bool with_score) {
    return TM_DISPATCH3(occ_score_full_v, fused, with_score, score_cull_enabled());
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
// One warp per hypothesis, subset walked in order 32 elements at a time.
// g_of_hyp: subset row of each hypothesis (null => row 0); sub_idx null => identity.
// bounding box of every 32 consecutive positions of every subset row (mask_ ignored: a
// superset box is still conservative).  One warp per tile; grid.y = group.
__global__ void __launch_bounds__(256)
    subset_tile_boxes_kernel(CloudDev scene, const int32_t* __restrict__ sub_idx,
                             const unsigned long long* __restrict__ sub_off, float4* __restrict__ tile_lo,
                             float4* __restrict__ tile_hi) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.y;
    const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned long long sb = sub_off[g];
    const uint32_t nsub = (uint32_t)(sub_off[g + 1] - sb);
    if ((unsigned long long)t * 32ull >= nsub) return;
    const uint32_t q = t * 32u + lane;
    float mnx = 3.0e38f, mny = 3.0e38f, mnz = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f, mxz = -3.0e38f;
    if (q < nsub) {
        const uint32_t idx = sub_idx ? (uint32_t)sub_idx[sb + q] : (uint32_t)(sb + q);
        const float4 v = scene.pos[idx];
        if (v.x == v.x && v.y == v.y && v.z == v.z) {
            mnx = mxx = v.x; mny = mxy = v.y; mnz = mxz = v.z;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, d));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, d));
    }
    if (lane == 0) {
        const size_t e = (size_t)(sb / 32ull) + g + t;
        tile_lo[e] = make_float4(mnx, mny, mnz, 0.f);
        tile_hi[e] = make_float4(mxx, mxy, mxz, 0.f);
    }
}
void launch_subset_tile_boxes(cudaStream_t st, const CloudDev& scene, const int32_t* sub_idx,
                              const unsigned long long* sub_off, uint32_t n_groups, uint32_t max_sub,
                              float4* tile_lo, float4* tile_hi) {
    if (!n_groups || !max_sub) return;
    const uint32_t tiles = (max_sub + 31) / 32;
    ++g_launch_count;
    dim3 grid((tiles + 7) / 8, n_groups);
    subset_tile_boxes_kernel<<<grid, 256, 0, st>>>(scene, sub_idx, sub_off, tile_lo, tile_hi);
}

// rows of a subset CSR rewritten in the evenly sampling walk order: out[row + p] = in[row + (p * s) mod n]
// (in == null: identity rows).  grid.y = group.
__global__ void __launch_bounds__(256)
    walk_order_rows_kernel(const int32_t* __restrict__ in, const unsigned long long* __restrict__ sub_off,
                           int32_t* __restrict__ out) {
    const uint32_t g = blockIdx.y;
    const unsigned long long sb = sub_off[g];
    const uint32_t n = (uint32_t)(sub_off[g + 1] - sb);
    const uint32_t s = walk_stride(n);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t e = (uint32_t)(((unsigned long long)p * s) % n);
        out[sb + p] = in ? in[sb + e] : (int32_t)(sb + e);
    }
}
void launch_walk_order_rows(cudaStream_t st, const int32_t* in, const unsigned long long* sub_off, uint32_t n_groups,
                            uint32_t max_sub, int32_t* out) {
    if (!n_groups || !max_sub) return;
    ++g_launch_count;
    const uint32_t bx = std::min<uint32_t>((max_sub + 255) / 256, 64u);
    walk_order_rows_kernel<<<dim3(bx, n_groups), 256, 0, st>>>(in, sub_off, out);
}

template <bool FUSED>
__global__ void __launch_bounds__(256)
    score_early_drop_kernel(EarlyArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (a.n_hyp_dev ? *a.n_hyp_dev : a.n_hyp)) return;
    const uint32_t h = a.hyp_list ? a.hyp_list[wid] : wid;
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
    // Positions are walked in steps of 32.  With tile boxes, 32 steps are screened at once: lane l
    // pushes the box of step s0+l through the transform (interval arithmetic, same guard as the
    // full scorer); a step whose box misses the grid has no reaching element, so it changes neither
    // the counts nor the checkpoint state and is skipped.
    const ModelDev& m = a.model;
    const uint32_t n_steps = (nsub + 31u) / 32u;
    const size_t tile_base = (size_t)(sb / 32ull) + g;
    for (uint32_t s0 = 0; s0 < n_steps && !dropped; s0 += 32) {
        uint32_t live = 0xffffffffu;
        if (a.tile_lo) {
            bool survive = false;
            const uint32_t st = s0 + lane;
            if (st < n_steps) {
                const float4 lo = a.tile_lo[tile_base + st], hi = a.tile_hi[tile_base + st];
                if (lo.x <= hi.x) {  // at least one finite point
                    const float cx = 0.5f * (lo.x + hi.x), hx = 0.5f * (hi.x - lo.x);
                    const float cy = 0.5f * (lo.y + hi.y), hy = 0.5f * (hi.y - lo.y);
                    const float cz = 0.5f * (lo.z + hi.z), hz = 0.5f * (hi.z - lo.z);
                    const float acx = fabsf(cx) + hx, acy = fabsf(cy) + hy, acz = fabsf(cz) + hz;
                    bool out = false;
#define TM_AXIS(r, S, TV, EXF)                                                                  \
    {                                                                                           \
        float cc = r.x * cx + r.y * cy + r.z * cz + r.w;                                        \
        float ee = fabsf(r.x) * hx + fabsf(r.y) * hy + fabsf(r.z) * hz;                         \
        float mag = fabsf(r.x) * acx + fabsf(r.y) * acy + fabsf(r.z) * acz + fabsf(r.w);        \
        ee += 1e-5f * mag + 1e-30f;                                                             \
        float sl = 1e-5f * (S * mag + fabsf(TV)) + 1e-30f;                                      \
        float lo_ = S * (cc - ee) + TV - sl, hi_ = S * (cc + ee) + TV + sl;                     \
        out = out || (lo_ >= EXF) || (hi_ <= -1.0f);                                            \
    }
                    TM_AXIS(r0, m.sx, m.tx, m.exf)
                    TM_AXIS(r1, m.sy, m.ty, m.eyf)
                    TM_AXIS(r2, m.sz, m.tz, m.ezf)
#undef TM_AXIS
                    survive = !out;
                }
            }
            live = __ballot_sync(0xffffffffu, survive);
        } else {
            const uint32_t left = n_steps - s0;
            live = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
        }
        while (live && !dropped) {
            const uint32_t base = (s0 + (uint32_t)(__ffs(live) - 1)) * 32u;
            live &= live - 1u;
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
                    inl = a.model.occ ? point_test<FUSED, true>(a.model, r0, r1, r2, v.x, v.y, v.z, fl, a.sq_thres, x, y, z, lin)
                                      : point_test<FUSED, false>(a.model, r0, r1, r2, v.x, v.y, v.z, fl, a.sq_thres, x, y, z, lin);
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

// First reaching element of one checkpoint range ("level", k_early2.cu) per hypothesis, found by walking the range in
// walk order: key = (position << 1) | !inlier into minkey[h], untouched (0xFFFFFFFF) when the range reaches nothing.
// One warp per listed hypothesis; on real data the first 32 positions hold a reaching element.  This is what a
// checkpoint needs of its own range, so a range can be scored after its checkpoint, for the survivors only.
template <bool FUSED>
__global__ void __launch_bounds__(256)
    level_probe_kernel(ProbeArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= *a.n_alive) return;
    const uint32_t h = a.hl[wid];
    const uint32_t g = a.g_of_hyp[h];
    const unsigned long long sb = a.sub_off[g];
    const uint32_t nsub = (uint32_t)(a.sub_off[g + 1] - sb);
    const uint32_t b0 = level_begin(nsub, a.level), b1 = level_begin(nsub, a.level + 1);
    const float4 r0 = a.T[3 * (size_t)h], r1 = a.T[3 * (size_t)h + 1], r2 = a.T[3 * (size_t)h + 2];
    for (uint32_t p0 = b0; p0 < b1; p0 += 32) {
        const uint32_t p = p0 + lane;
        bool reach = false, inl = false;
        if (p < b1) {
            const uint32_t idx = (uint32_t)a.sub_idx_walk[sb + p];
            const float4 v = a.scene.pos[idx];
            const uint32_t fl = __float_as_uint(v.w);
            if (!(fl & FLAG_MASKED)) {
                float x, y, z;
                uint32_t lin = 0xffffffffu;
                inl = a.model.occ ? point_test<FUSED, true>(a.model, r0, r1, r2, v.x, v.y, v.z, fl, a.sq_thres, x, y, z, lin)
                                  : point_test<FUSED, false>(a.model, r0, r1, r2, v.x, v.y, v.z, fl, a.sq_thres, x, y, z, lin);
                reach = lin != 0xffffffffu;
            }
        }
        const uint32_t rm = __ballot_sync(0xffffffffu, reach);
        if (rm) {
            const int l = __ffs(rm) - 1;
            const uint32_t im = __ballot_sync(0xffffffffu, inl);
            if (lane == 0) a.minkey[h] = ((p0 + (uint32_t)l) << 1) | (((im >> l) & 1u) ? 0u : 1u);
            return;
        }
    }
}
void launch_level_probe(cudaStream_t st, const ProbeArgs& a, uint32_t n_bound, bool fused) {
    if (!n_bound) return;
    ++g_launch_count;
    const unsigned grid = (n_bound + 7) / 8;
    if (fused) level_probe_kernel<true><<<grid, 256, 0, st>>>(a);
    else level_probe_kernel<false><<<grid, 256, 0, st>>>(a);
}

// ---------------------------------------------------------------------- argmax
// key = (inliers << 32) | (0xFFFFFFFF - global hypothesis id); ties -> lowest id.
__global__ void __launch_bounds__(256)
    argmax_kernel(const uint32_t* __restrict__ counts, const uint8_t* __restrict__ valid,
                  const uint8_t* __restrict__ excluded, const uint32_t* __restrict__ n_local,
                  const unsigned long long* __restrict__ h_begin, unsigned long long* __restrict__ best) {
    __shared__ unsigned long long wbest[8];
    const uint32_t n = *n_local;
    const unsigned long long hb = *h_begin;
    unsigned long long k = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        if (excluded && excluded[i]) continue;  // dropped by the early-drop test
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
void launch_argmax(cudaStream_t st, const uint32_t* counts, const uint8_t* valid, const uint8_t* excluded,
                   const uint32_t* n_local, const unsigned long long* h_begin,
                   unsigned long long* best, int grid) {
    ++g_launch_count;
    argmax_kernel<<<grid, 256, 0, st>>>(counts, valid, excluded, n_local, h_begin, best);
}

}  // namespace tmk
