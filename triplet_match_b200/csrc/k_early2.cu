// k_early2.cu — project_(early_out = true) (include/impl/scene.hpp:411-510, checkpoints :422-426, 492-506) over the
// evenly sampling walk of each subset (early_out = 2), evaluated LEVEL BY LEVEL with the tiled, box-culled scorer
// instead of one warp walking one hypothesis (score_early_drop_kernel).
//
// The reference walks a subset in order and, at the first *reaching* element (unmasked, voxel_query succeeded) at or
// after 5 %, 10 %, ..., 90 % of it, extrapolates the final count from the inliers so far and drops the hypothesis when
// the bound falls below the acceptance threshold — 18 checkpoints, at most one per element.  The 18 thresholds cut the
// walk positions into 19 ranges ("levels"); in walk order p -> (p * s) mod n a level is every ~20th point of the ball,
// so the points of one level, kept in the subset's own (spatial) order, tile like the whole subset does.  Per level:
//
//   score_level_kernel   the count-only scorer of k_score2.cu over the level's tiles x the group's hypotheses; dead
//                        hypotheses (dropped at an earlier checkpoint) are skipped in the cull phase.  Besides the
//                        level's inlier count it reduces, per hypothesis, the smallest walk position of a reaching
//                        element together with that element's inlier bit (atomicMin of pos << 1 | !inlier).
//   el_eval_kernel       checkpoint L fires at the first reaching position of level L (the previous checkpoint fired in
//                        level L - 1, so "one checkpoint per element" is automatic): corrs = inliers of the levels
//                        before + that element's inlier bit, tried = position + 1, same bound, same comparison.  A
//                        hypothesis that fails is dropped with exactly the walker's partial count.
//
// One launch ("stage") scores up to four consecutive levels, each tile holding 32-point chunks of all of them over the
// same stretch of the subset (narrower tiles, fewer of them pass the box cull) with per-level accumulators; el_eval_kernel
// then applies the stage's checkpoints in order, and the hypotheses still alive are compacted into a list the next
// stage's work items run over.  Checkpoint 1 — where most of the dropping happens — is applied from a probe of level 1's
// first reaching element (level_probe_kernel, k_score.cu), so that level 1 is only scored for its survivors.
//
// A level without a reaching element (or an empty level: subsets of a few points) breaks the one-level-one-checkpoint
// correspondence; such hypotheses are flagged and re-walked exactly by score_early_drop_kernel (they touch the grid
// almost nowhere, so they are few and cheap).  Counts, drop flags and drop points equal the walker's bit for bit
// (tests/test_gpu_parity.py); the un-normalised partial scores are produced by the walker on request.
#include <algorithm>

#include "tm_kernels.cuh"
#include "tm_x2.cuh"

namespace tmk {

// modular inverse of s modulo n (gcd(s, n) == 1, n >= 1)
__device__ inline uint32_t mod_inverse(uint32_t s, uint32_t n) {
    if (n <= 1u) return 0u;
    long long t = 0, nt = 1, r = (long long)n, nr = (long long)(s % n);
    while (nr) {
        const long long qq = r / nr;
        long long tmp = t - qq * nt; t = nt; nt = tmp;
        tmp = r - qq * nr; r = nr; nr = tmp;
    }
    if (t < 0) t += (long long)n;
    return (uint32_t)t;
}

// Rows of the subset CSR regrouped by level: within a group the elements of level L occupy
// [sub_off[g] + level_begin(n, L), sub_off[g] + level_begin(n, L + 1)), in ascending element (= spatial) order.
// lvl_idx = scene index, lvl_pos = walk position.  A group is cut into chunks of WL_CHUNK elements, one CTA each:
// walk_levels_hist_kernel counts the chunk's elements per level, walk_levels_kernel places them — a stable partition
// by level through __match_any_sync ranks, starting at the level's base plus what the chunks before it hold.
constexpr uint32_t WL_CHUNK = 2048;
__device__ __forceinline__ int level_of(uint32_t p, const uint32_t* bnd) {
    int L = 0;
#pragma unroll
    for (int k = 1; k < EL_LEVELS; ++k) L += (p >= bnd[k]) ? 1 : 0;
    return L;
}
__global__ void __launch_bounds__(256)
    walk_levels_hist_kernel(const unsigned long long* __restrict__ sub_off, uint32_t max_chunks, uint32_t* __restrict__ hist) {
    __shared__ uint32_t bnd[EL_LEVELS + 1];
    __shared__ uint32_t cnt[EL_LEVELS];
    __shared__ uint32_t sinv_s;
    const uint32_t g = blockIdx.y, chunk = blockIdx.x;
    const uint32_t n = (uint32_t)(sub_off[g + 1] - sub_off[g]);
    const uint32_t e0 = chunk * WL_CHUNK;
    if (e0 >= n) return;
    if (threadIdx.x <= EL_LEVELS) bnd[threadIdx.x] = level_begin(n, (int)threadIdx.x);
    if (threadIdx.x < EL_LEVELS) cnt[threadIdx.x] = 0u;
    if (threadIdx.x == 0) sinv_s = mod_inverse(walk_stride(n), n);
    __syncthreads();
    const uint32_t sinv = sinv_s;
    const int lane = threadIdx.x & 31;
    const uint32_t e1 = min(n, e0 + WL_CHUNK);
    for (uint32_t base = e0; base < e1; base += blockDim.x) {  // uniform trip count: every lane reaches the match
        const uint32_t e = base + threadIdx.x;
        int L = EL_LEVELS;  // dead lanes form their own match group
        if (e < e1) L = level_of((uint32_t)(((unsigned long long)e * sinv) % n), bnd);
        const uint32_t peers = __match_any_sync(0xffffffffu, L);
        if (L < EL_LEVELS && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(&cnt[L], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < EL_LEVELS) hist[((size_t)g * max_chunks + chunk) * EL_LEVELS + threadIdx.x] = cnt[threadIdx.x];
}
__global__ void __launch_bounds__(256)
    walk_levels_kernel(const int32_t* __restrict__ sub_idx, const unsigned long long* __restrict__ sub_off,
                       uint32_t max_chunks, const uint32_t* __restrict__ hist, int32_t* __restrict__ lvl_idx,
                       uint32_t* __restrict__ lvl_pos) {
    __shared__ uint32_t bnd[EL_LEVELS + 1];
    __shared__ uint32_t run[EL_LEVELS];
    __shared__ uint32_t wcnt[8][EL_LEVELS];
    __shared__ uint32_t sinv_s;
    const uint32_t g = blockIdx.y, chunk = blockIdx.x;
    const unsigned long long sb = sub_off[g];
    const uint32_t n = (uint32_t)(sub_off[g + 1] - sb);
    const uint32_t e0 = chunk * WL_CHUNK;
    if (e0 >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x <= EL_LEVELS) bnd[threadIdx.x] = level_begin(n, (int)threadIdx.x);
    if (threadIdx.x < EL_LEVELS) {  // what the chunks before this one put into the level
        uint32_t t = 0;
        for (uint32_t c = 0; c < chunk; ++c) t += hist[((size_t)g * max_chunks + c) * EL_LEVELS + threadIdx.x];
        run[threadIdx.x] = t;
    }
    if (threadIdx.x == 32) sinv_s = mod_inverse(walk_stride(n), n);
    __syncthreads();
    const uint32_t sinv = sinv_s;
    const uint32_t e1 = min(n, e0 + WL_CHUNK);
    for (uint32_t base = e0; base < e1; base += blockDim.x) {
        for (int t = threadIdx.x; t < 8 * EL_LEVELS; t += blockDim.x) (&wcnt[0][0])[t] = 0u;
        __syncthreads();
        const uint32_t e = base + threadIdx.x;
        const bool live = e < e1;
        uint32_t p = 0u;
        int L = EL_LEVELS;  // dead lanes form their own match group
        if (live) {
            p = (uint32_t)(((unsigned long long)e * sinv) % n);  // (p * s) mod n == e
            L = level_of(p, bnd);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, L);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (live && rank == 0u) wcnt[warp][L] = __popc(peers);
        __syncthreads();
        if (live) {
            uint32_t off = run[L];
            for (int w = 0; w < warp; ++w) off += wcnt[w][L];
            const unsigned long long dst = sb + bnd[L] + off + rank;
            lvl_idx[dst] = sub_idx ? sub_idx[sb + e] : (int32_t)(sb + e);
            lvl_pos[dst] = p;
        }
        __syncthreads();
        if (threadIdx.x < EL_LEVELS) {
            uint32_t t = 0u;
            for (int w = 0; w < 8; ++w) t += wcnt[w][threadIdx.x];
            run[threadIdx.x] += t;
        }
        __syncthreads();
    }
}
size_t walk_levels_hist_bytes(uint32_t n_groups, uint32_t max_sub) {
    const size_t chunks = (max_sub + WL_CHUNK - 1) / WL_CHUNK;
    return std::max<size_t>(1, (size_t)n_groups * chunks * EL_LEVELS) * 4;
}
void launch_walk_levels(cudaStream_t st, const int32_t* sub_idx, const unsigned long long* sub_off, uint32_t n_groups,
                        uint32_t max_sub, uint32_t* hist, int32_t* lvl_idx, uint32_t* lvl_pos) {
    if (!n_groups || !max_sub) return;
    const uint32_t chunks = (max_sub + WL_CHUNK - 1) / WL_CHUNK;
    g_launch_count += 2;
    walk_levels_hist_kernel<<<dim3(chunks, n_groups), 256, 0, st>>>(sub_off, chunks, hist);
    walk_levels_kernel<<<dim3(chunks, n_groups), 256, 0, st>>>(sub_idx, sub_off, chunks, hist, lvl_idx, lvl_pos);
}

// ---- stages and their work lists -------------------------------------------------------------------------------------
// A stage scores M consecutive levels in one launch.  Its tiles take 32-point chunks from each of the M levels over
// the same stretch of the subset: point group k (the lane's k-th point, k = 0..3) of tile j holds chunk
// j * (4 / M) + k / M of level L0 + k % M.  A level is an even 1/20 sample of the subset, so the chunks of one tile
// cover about the same elements and a tile of a 4-level stage is half as wide as four chunks of one level — which is
// what decides how many (tile, hypothesis) pairs the box cull has to let through.  Each level keeps its own
// accumulators, so the checkpoints are applied exactly as if the levels had been scored one after the other; what a
// stage gives up is the work on the later levels of hypotheses that one of its own checkpoints drops.
__host__ __device__ inline uint32_t stage_tiles(uint32_t n, int L0, int M) {
    const uint32_t per = (uint32_t)(4 / M) * 32u;  // points of one level per tile
    uint32_t tiles = 0;
    for (int r = 0; r < M; ++r) {
        const uint32_t cnt = level_begin(n, L0 + r + 1) - level_begin(n, L0 + r);
        const uint32_t t = (cnt + per - 1) / per;
        tiles = t > tiles ? t : tiles;
    }
    return tiles;
}
// Work items of one stage: (tile, 256-hypothesis chunk) over the hypotheses still alive.  The alive hypotheses are kept
// as a compacted list hl[] grouped by subset (ranges goff[g] .. goff[g + 1]); an item's hypothesis range is a range
// of list positions.
__global__ void el_work_count_kernel(const unsigned long long* __restrict__ sub_off, const uint32_t* __restrict__ goff,
                                     uint32_t n_groups, int L0, int M, uint32_t* __restrict__ n_items) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t n = (uint32_t)(sub_off[g + 1] - sub_off[g]);
    const uint32_t nh = goff[g + 1] - goff[g];
    n_items[g] = nh ? stage_tiles(n, L0, M) * ((nh + EL_HCHUNK - 1) / EL_HCHUNK) : 0u;
}
// item: sub_begin = first row of the subset, npts = its size, pad = tile index within the stage
__global__ void el_work_fill_kernel(const unsigned long long* __restrict__ sub_off, const uint32_t* __restrict__ goff,
                                    uint32_t n_groups, int L0, int M, const uint32_t* __restrict__ item_off,
                                    WorkItem* __restrict__ items) {
    const uint32_t g = blockIdx.x;
    if (g >= n_groups) return;
    const unsigned long long sb = sub_off[g];
    const uint32_t n = (uint32_t)(sub_off[g + 1] - sb);
    const uint32_t hb = goff[g], nh = goff[g + 1] - hb;
    if (!nh) return;
    const uint32_t tiles = stage_tiles(n, L0, M), chunks = (nh + EL_HCHUNK - 1) / EL_HCHUNK;
    const uint32_t base = item_off[g];
    for (uint32_t k = threadIdx.x; k < tiles * chunks; k += blockDim.x) {
        const uint32_t tile = k % tiles, chunk = k / tiles;
        WorkItem w;
        w.sub_begin = sb;
        w.npts = n;
        w.hyp_begin = hb + chunk * EL_HCHUNK;
        w.hyp_end = min(hb + nh, w.hyp_begin + EL_HCHUNK);
        w.pad = tile;
        items[base + k] = w;
    }
}
void launch_el_work_count(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* goff, uint32_t n_groups,
                          int L0, int M, uint32_t* n_items) {
    if (!n_groups) return;
    ++g_launch_count;
    el_work_count_kernel<<<(n_groups + 127) / 128, 128, 0, st>>>(sub_off, goff, n_groups, L0, M, n_items);
}
void launch_el_work_fill(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* goff, uint32_t n_groups,
                         int L0, int M, const uint32_t* item_off, WorkItem* items) {
    if (!n_groups) return;
    ++g_launch_count;
    el_work_fill_kernel<<<n_groups, 128, 0, st>>>(sub_off, goff, n_groups, L0, M, item_off, items);
}
// The alive list after a stage's checkpoints: per subset group, the hypotheses still alive, order kept.  One CTA per
// group: count, then (after a scan of the counts over the groups) an ordered compaction (warp ballots + a prefix over
// the CTA's warps).
__global__ void __launch_bounds__(256)
    el_alive_count_kernel(const uint32_t* __restrict__ hl, const uint32_t* __restrict__ goff, uint32_t n_groups,
                          const uint8_t* __restrict__ alive, uint32_t* __restrict__ cnt) {
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.x;
    uint32_t c = 0;
    for (uint32_t p = goff[g] + threadIdx.x; p < goff[g + 1]; p += blockDim.x) c += alive[hl[p]] ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) wsum[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        cnt[g] = t;
    }
}
__global__ void __launch_bounds__(256)
    el_alive_fill_kernel(const uint32_t* __restrict__ hl, const uint32_t* __restrict__ goff, uint32_t n_groups,
                         const uint8_t* __restrict__ alive, const uint32_t* __restrict__ goff_new,
                         uint32_t* __restrict__ hl_new) {
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.x;
    uint32_t out = goff_new[g];
    const uint32_t b = goff[g], e = goff[g + 1];
    for (uint32_t p0 = b; p0 < e; p0 += blockDim.x) {  // uniform trip count across the CTA
        const uint32_t p = p0 + threadIdx.x;
        const uint32_t h = p < e ? hl[p] : 0u;
        const bool keep = p < e && alive[h];
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            before += w < warp ? wsum[w] : 0u;
            total += wsum[w];
        }
        if (keep) hl_new[out + before + __popc(m & ((1u << lane) - 1u))] = h;
        out += total;
        __syncthreads();  // wsum is rewritten by the next round
    }
}
void launch_el_alive_count(cudaStream_t st, const uint32_t* hl, const uint32_t* goff, uint32_t n_groups, const uint8_t* alive,
                           uint32_t* cnt) {
    if (!n_groups) return;
    ++g_launch_count;
    el_alive_count_kernel<<<n_groups, 256, 0, st>>>(hl, goff, n_groups, alive, cnt);
}
void launch_el_alive_fill(cudaStream_t st, const uint32_t* hl, const uint32_t* goff, uint32_t n_groups, const uint8_t* alive,
                          const uint32_t* goff_new, uint32_t* hl_new) {
    if (!n_groups) return;
    ++g_launch_count;
    el_alive_fill_kernel<<<n_groups, 256, 0, st>>>(hl, goff, n_groups, alive, goff_new, hl_new);
}
uint64_t el_items_bound(uint64_t n_points, uint64_t n_hyp) {
    // (tiles x hypothesis chunks) of the largest stage: a stage holds at most 4 of the 19 levels plus the double-sized
    // last one, i.e. well under half of the subset's 128-point tiles; bounded by all of them
    return (n_points / SCORE_TILE + 4) * ((n_hyp + EL_HCHUNK - 1) / EL_HCHUNK + 1);
}

// ---- one level: counts + first reaching position per live hypothesis ------------------------------------------------
// Exact test of one hypothesis against the tile's point pairs A (point groups 0, 1) and / or B (2, 3): per point group k
// of this lane the inlier bit (cbits) and, for a reaching element, key = walk position << 1 | !inlier.  A pair the cull
// excluded is not evaluated (it holds no reaching element, so it contributes neither).  Returns false when nothing
// reaches the grid.
template <bool FUSED, bool OCC, bool DO_A, bool DO_B>
__device__ __forceinline__ bool level_eval(const ModelDev& m, const X2& e, float sq_thres, float4 r0, float4 r1, float4 r2,
                                           p2 pxA, p2 pyA, p2 pzA, p2 pxB, p2 pyB, p2 pzB, const uint32_t (&wpos)[4],
                                           uint32_t tflags, uint32_t& cbits, uint32_t (&key)[4]) {
    float x[4], y[4], z[4], vx[4], vy[4], vz[4];
    if (DO_A) {
        const p2 xA = e.row_apply(r0, pxA, pyA, pzA), yA = e.row_apply(r1, pxA, pyA, pzA), zA = e.row_apply(r2, pxA, pyA, pzA);
        const p2 vxA = e.add(e.mul(m.sx, xA), m.tx), vyA = e.add(e.mul(m.sy, yA), m.ty), vzA = e.add(e.mul(m.sz, zA), m.tz);
        x[0] = lo2(xA); x[1] = hi2(xA); y[0] = lo2(yA); y[1] = hi2(yA); z[0] = lo2(zA); z[1] = hi2(zA);
        vx[0] = lo2(vxA); vx[1] = hi2(vxA); vy[0] = lo2(vyA); vy[1] = hi2(vyA); vz[0] = lo2(vzA); vz[1] = hi2(vzA);
    }
    if (DO_B) {
        const p2 xB = e.row_apply(r0, pxB, pyB, pzB), yB = e.row_apply(r1, pxB, pyB, pzB), zB = e.row_apply(r2, pxB, pyB, pzB);
        const p2 vxB = e.add(e.mul(m.sx, xB), m.tx), vyB = e.add(e.mul(m.sy, yB), m.ty), vzB = e.add(e.mul(m.sz, zB), m.tz);
        x[2] = lo2(xB); x[3] = hi2(xB); y[2] = lo2(yB); y[3] = hi2(yB); z[2] = lo2(zB); z[3] = hi2(zB);
        vx[2] = lo2(vxB); vx[3] = hi2(vxB); vy[2] = lo2(vyB); vy[3] = hi2(vyB); vz[2] = lo2(vzB); vz[3] = hi2(vzB);
    }
    uint32_t lin[4];
    bool reach[4], look[4];
    bool any_reach = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        reach[k] = look[k] = false;
        lin[k] = 0u;
        if (k < 2 ? DO_A : DO_B) {
            // voxel_query succeeded (model.hpp:186-189) in the float domain: int(v) in [0, e) <=> -1 < v < e for every
            // finite v; NaN (masked / padding points) and +-inf are out, as with the walker
            const bool ok = (vx[k] > -1.f) & (vx[k] < m.exf) & (vy[k] > -1.f) & (vy[k] < m.eyf) & (vz[k] > -1.f) &
                            (vz[k] < m.ezf);
            const int i = (int)vx[k], j = (int)vy[k], kk = (int)vz[k];
            bool lk = ok;
            if (OCC) {  // the occupancy mask only saves the gather; the element reaches either way
                const uint32_t b = ok ? (uint32_t)(((kk >> OCC_SHIFT) * m.oby + (j >> OCC_SHIFT)) * m.obx + (i >> OCC_SHIFT))
                                      : 0u;
                lk = ok & (((__ldg(&m.occ[b >> 5]) >> (b & 31u)) & 1u) != 0u);
            }
            lin[k] = (uint32_t)((kk * m.ey + j) * m.ex + i);
            reach[k] = ok;
            look[k] = lk;
            any_reach |= ok;
        }
    }
    if (!__any_sync(0xffffffffu, any_reach)) return false;  // nothing reaches the grid
    float4 mp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((k < 2 ? DO_A : DO_B) && look[k]) {
            if (FUSED) mp[k] = __ldg(&m.vcell[lin[k]]);
            else mp[k] = __ldg(&m.cloud.pos[__ldg(&m.voxel[lin[k]])]);
        }
    }
    float sq[4];
    if (DO_A) {
        const p2 sqA = e.sqnorm(pack2(x[0] - mp[0].x, x[1] - mp[1].x), pack2(y[0] - mp[0].y, y[1] - mp[1].y),
                                pack2(z[0] - mp[0].z, z[1] - mp[1].z));
        sq[0] = lo2(sqA); sq[1] = hi2(sqA);
    }
    if (DO_B) {
        const p2 sqB = e.sqnorm(pack2(x[2] - mp[2].x, x[3] - mp[3].x), pack2(y[2] - mp[2].y, y[3] - mp[3].y),
                                pack2(z[2] - mp[2].z, z[3] - mp[3].z));
        sq[2] = lo2(sqB); sq[3] = hi2(sqB);
    }
    cbits = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        key[k] = 0xffffffffu;
        if (k < 2 ? DO_A : DO_B) {
            const uint32_t pfl = (tflags >> k) & 1u;
            const bool inl = look[k] && (sq[k] <= sq_thres) && (((pfl ^ __float_as_uint(mp[k].w)) & FLAG_TANGENT) == 0u);
            cbits |= inl ? (1u << k) : 0u;
            if (reach[k]) key[k] = (wpos[k] << 1) | (inl ? 0u : 1u);
        }
    }
    return true;
}

// bounding box of one point pair's 64 tile points -> centre / half extents; valid = the pair holds a live point
struct HalfBox {
    float cx, cy, cz, hx, hy, hz;
    bool valid;
};
__device__ __forceinline__ HalfBox half_box(float mnx, float mny, float mnz, float mxx, float mxy, float mxz) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, d));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, d));
    }
    HalfBox b;
    b.valid = mnx <= mxx;
    b.cx = 0.5f * (mnx + mxx); b.hx = 0.5f * (mxx - mnx);
    b.cy = 0.5f * (mny + mxy); b.hy = 0.5f * (mxy - mny);
    b.cz = 0.5f * (mnz + mxz); b.hz = 0.5f * (mxz - mnz);
    return b;
}
// interval test of k_score2.cu's cull: true when the transformed box provably misses the grid (NaN never culls)
__device__ __forceinline__ bool box_misses_grid(const ModelDev& m, const HalfBox& b, float4 r0, float4 r1, float4 r2) {
    const float acx = fabsf(b.cx) + b.hx, acy = fabsf(b.cy) + b.hy, acz = fabsf(b.cz) + b.hz;
    bool out = false;
#define TM_AXIS(r, S, TV, EXF)                                                                        \
    {                                                                                                 \
        float cc = r.x * b.cx + r.y * b.cy + r.z * b.cz + r.w;                                        \
        float ee = fabsf(r.x) * b.hx + fabsf(r.y) * b.hy + fabsf(r.z) * b.hz;                         \
        float mag = fabsf(r.x) * acx + fabsf(r.y) * acy + fabsf(r.z) * acz + fabsf(r.w);              \
        ee += 1e-5f * mag + 1e-30f;                                                                   \
        float sl = 1e-5f * (S * mag + fabsf(TV)) + 1e-30f;                                            \
        float lo = S * (cc - ee) + TV - sl, hi = S * (cc + ee) + TV + sl;                             \
        out = out || (lo >= EXF) || (hi <= -1.0f);                                                    \
    }
    TM_AXIS(r0, m.sx, m.tx, m.exf)
    TM_AXIS(r1, m.sy, m.ty, m.eyf)
    TM_AXIS(r2, m.sz, m.tz, m.ezf)
#undef TM_AXIS
    return out;
}

template <bool FUSED, bool OCC, int M>
__global__ void __launch_bounds__(SCORE_THREADS, SCORE_MIN_BLOCKS)
    score_level_kernel(LevelArgs a, p2 k_nz, p2 k_one) {
    static_assert(SCORE_P == 4 && (M == 1 || M == 2 || M == 4), "two point pairs per lane, 4 / M chunks per level");
    __shared__ float4 s_rows[(SCORE_THREADS / 32) * 32 * 3];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float4* my_rows = s_rows + warp * 32 * 3;
    const uint32_t item_end = *a.n_items;
    const ModelDev& m = a.model;
    X2 e;
    e.nz = k_nz; e.one = k_one; e.mone = 0ull;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= item_end) break;
        const WorkItem w = a.items[item];
        // point group k of this lane: chunk w.pad * (4 / M) + k / M of level L0 + k % M (see stage_tiles); pairs
        // A = groups (0, 1), B = (2, 3).  For M < 4 the two pairs cover different stretches of the subset, so each gets
        // its own box and is culled on its own.
        float px[4], py[4], pz[4];
        uint32_t wpos[4];     // walk position of point k (only meaningful where the point is live)
        uint32_t tflags = 0;  // bit k: tangent_mask_ of point k
        const float nanv = __int_as_float(0x7fc00000);
        float mn[2][3] = {{3.0e38f, 3.0e38f, 3.0e38f}, {3.0e38f, 3.0e38f, 3.0e38f}};
        float mx[2][3] = {{-3.0e38f, -3.0e38f, -3.0e38f}, {-3.0e38f, -3.0e38f, -3.0e38f}};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            px[k] = py[k] = pz[k] = nanv;  // masked / padding points: NaN, never reaching, never inliers
            wpos[k] = 0x7fffffffu;
            const int lvl = a.L0 + (k % M);
            const uint32_t b0 = level_begin(w.npts, lvl), cnt = level_begin(w.npts, lvl + 1) - b0;
            const uint32_t q = (w.pad * (4 / M) + (uint32_t)(k / M)) * 32u + lane;  // position within the level
            if (q < cnt) {
                const unsigned long long row = w.sub_begin + b0 + q;
                const uint32_t idx = (uint32_t)a.lvl_idx[row];
                const float4 v = a.scene.pos[idx];
                const uint32_t fl = __float_as_uint(v.w);
                if (!(fl & FLAG_MASKED)) {  // mask_ (scene.hpp:434): skipped before anything is counted
                    px[k] = v.x; py[k] = v.y; pz[k] = v.z;
                    wpos[k] = a.lvl_pos[row];
                    if (fl & FLAG_TANGENT) tflags |= 1u << k;
                    mn[k >> 1][0] = fminf(mn[k >> 1][0], v.x); mx[k >> 1][0] = fmaxf(mx[k >> 1][0], v.x);
                    mn[k >> 1][1] = fminf(mn[k >> 1][1], v.y); mx[k >> 1][1] = fmaxf(mx[k >> 1][1], v.y);
                    mn[k >> 1][2] = fminf(mn[k >> 1][2], v.z); mx[k >> 1][2] = fmaxf(mx[k >> 1][2], v.z);
                }
            }
        }
        const HalfBox bA = half_box(mn[0][0], mn[0][1], mn[0][2], mx[0][0], mx[0][1], mx[0][2]);
        const HalfBox bB = half_box(mn[1][0], mn[1][1], mn[1][2], mx[1][0], mx[1][1], mx[1][2]);
        if (!bA.valid && !bB.valid) continue;  // no live (finite, unmasked) point in this tile
        const p2 pxA = e.add(pack2(px[0], px[1]), e.nz), pyA = e.add(pack2(py[0], py[1]), e.nz),
                 pzA = e.add(pack2(pz[0], pz[1]), e.nz);
        const p2 pxB = e.add(pack2(px[2], px[3]), e.nz), pyB = e.add(pack2(py[2], py[3]), e.nz),
                 pzB = e.add(pack2(pz[2], pz[3]), e.nz);
        for (uint32_t h0 = w.hyp_begin; h0 < w.hyp_end; h0 += 32) {  // positions in the list of hypotheses still alive
            const bool live = h0 + lane < w.hyp_end;
            const uint32_t h = live ? a.hl[h0 + lane] : 0u;
            bool sA = false, sB = false;
            __syncwarp();  // readers of the previous batch's rows are done
            if (live) {
                const float4 r0 = __ldg(&a.T[3 * (size_t)h]), r1 = __ldg(&a.T[3 * (size_t)h + 1]),
                             r2 = __ldg(&a.T[3 * (size_t)h + 2]);
                my_rows[lane] = r0;
                my_rows[32 + lane] = r1;
                my_rows[64 + lane] = r2;
                // a culled pair holds no reaching element: it changes neither count nor checkpoint
                sA = bA.valid && !box_misses_grid(m, bA, r0, r1, r2);
                sB = bB.valid && !box_misses_grid(m, bB, r0, r1, r2);
            }
            const uint32_t maskA = __ballot_sync(0xffffffffu, sA);  // also orders the smem stores
            const uint32_t maskB = __ballot_sync(0xffffffffu, sB);
            uint32_t mask = maskA | maskB;
            if (a.stats) {
                const uint32_t al = __ballot_sync(0xffffffffu, live);
                if (lane == 0) {
                    atomicAdd(&a.stats[0], (unsigned long long)min(32u, w.hyp_end - h0));
                    atomicAdd(&a.stats[1], (unsigned long long)__popc(al));
                    atomicAdd(&a.stats[2], (unsigned long long)__popc(mask));
                    atomicAdd(&a.stats[4], (unsigned long long)(__popc(maskA) + __popc(maskB)));
                }
            }
            uint32_t mycnt = 0;  // level r of the stage: bits 8r .. 8r+7 (at most 128 points of a level per tile)
            uint32_t mymin[M];
#pragma unroll
            for (int r = 0; r < M; ++r) mymin[r] = 0xffffffffu;
            while (mask) {
                const int hh = __ffs(mask) - 1;
                mask &= mask - 1u;
                const float4 r0 = my_rows[hh], r1 = my_rows[32 + hh], r2 = my_rows[64 + hh];
                const bool doA = (maskA >> hh) & 1u, doB = (maskB >> hh) & 1u;  // warp-uniform
                uint32_t cbits, key[4];
                bool any;
                if (doA && doB)
                    any = level_eval<FUSED, OCC, true, true>(m, e, a.sq_thres, r0, r1, r2, pxA, pyA, pzA, pxB, pyB, pzB, wpos,
                                                             tflags, cbits, key);
                else if (doA)
                    any = level_eval<FUSED, OCC, true, false>(m, e, a.sq_thres, r0, r1, r2, pxA, pyA, pzA, pxB, pyB, pzB, wpos,
                                                              tflags, cbits, key);
                else
                    any = level_eval<FUSED, OCC, false, true>(m, e, a.sq_thres, r0, r1, r2, pxA, pyA, pzA, pxB, pyB, pzB, wpos,
                                                              tflags, cbits, key);
                if (!any) continue;
                if (a.stats && lane == 0) atomicAdd(&a.stats[3], 1ull);
#pragma unroll
                for (int r = 0; r < M; ++r) {  // the point groups k with k % M == r belong to level L0 + r
                    uint32_t c = 0, mnk = 0xffffffffu;
#pragma unroll
                    for (int k = r; k < 4; k += M) {
                        c += (cbits >> k) & 1u;
                        mnk = min(mnk, key[k]);
                    }
                    const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
                    const uint32_t wmn = __reduce_min_sync(0xffffffffu, mnk);
                    if (lane == hh) {
                        mycnt |= tot << (8 * r);
                        mymin[r] = wmn;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const uint32_t c = (mycnt >> (8 * r)) & 0xffu;
                if (c) atomicAdd(&a.lvl_cnt[(size_t)r * a.cap + h], c);
                if (mymin[r] != 0xffffffffu) atomicMin(&a.minkey[(size_t)r * a.cap + h], mymin[r]);
            }
        }
    }
}
template <bool FUSED, bool OCC>
static void launch_level_m(cudaStream_t st, const LevelArgs& a, int grid, int M) {
    const p2 nz = host_pair(-0.0f), one = host_pair(1.0f);
    if (M == 4) score_level_kernel<FUSED, OCC, 4><<<grid, SCORE_THREADS, 0, st>>>(a, nz, one);
    else if (M == 2) score_level_kernel<FUSED, OCC, 2><<<grid, SCORE_THREADS, 0, st>>>(a, nz, one);
    else score_level_kernel<FUSED, OCC, 1><<<grid, SCORE_THREADS, 0, st>>>(a, nz, one);
}
void launch_score_level(cudaStream_t st, const LevelArgs& a, int grid, bool fused, int M) {
    ++g_launch_count;
    if (fused) {
        if (a.model.occ) launch_level_m<true, true>(st, a, grid, M);
        else launch_level_m<true, false>(st, a, grid, M);
    } else {
        if (a.model.occ) launch_level_m<false, true>(st, a, grid, M);
        else launch_level_m<false, false>(st, a, grid, M);
    }
}
int score_level_max_blocks_per_sm(bool fused) {
    int nb = 0, nb2 = 0;
    if (fused) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_level_kernel<true, false, 4>, SCORE_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, score_level_kernel<true, true, 4>, SCORE_THREADS, 0);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_level_kernel<false, false, 4>, SCORE_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, score_level_kernel<false, true, 4>, SCORE_THREADS, 0);
    }
    nb = nb < nb2 ? nb : nb2;
    return nb > 0 ? nb : 1;
}

// ---- the checkpoints of a stage's levels, in order (level 0 has none; level 18 also closes the walk) ---------------
__global__ void __launch_bounds__(256)
    el_eval_kernel(EvalArgs a) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;  // position in the list of hypotheses still alive
    unsigned long long tested = 0;  // walk positions this hypothesis' walk ends with (summed per warp: one atomic)
    bool irregular = false;         // to be walked on its own (appended per warp: one atomic)
    const uint32_t h = pos < *a.n_alive ? a.hl[pos] : 0u;
    if (pos < *a.n_alive) {
        const uint32_t g = a.g_of_hyp[h];
        const uint32_t nsub = (uint32_t)(a.sub_off[g + 1] - a.sub_off[g]);
        uint32_t corrs = a.corrs[h];
        bool open = true;  // still walking
        for (int r = 0; r < a.M; ++r) {
            const size_t slot = (size_t)r * a.cap + h;
            const uint32_t cnt = a.lvl_cnt[slot], mk = a.minkey[slot];
            a.lvl_cnt[slot] = 0u;
            a.minkey[slot] = 0xffffffffu;
            if (!open) continue;  // (the accumulators of the later levels are still reset)
            const int L = a.L0 + r;
            if (L >= 1 && !(a.skip_first && r == 0)) {
                if (level_begin(nsub, L) == level_begin(nsub, L + 1) || mk == 0xffffffffu) {
                    // the level holds no reaching element: its checkpoint fires later, on an element another checkpoint
                    // may also claim — walked exactly afterwards
                    irregular = true;
                    open = false;
                    continue;
                }
                const uint32_t inl = (mk & 1u) ? 0u : 1u, tried = (mk >> 1) + 1u;
                const uint32_t c_here = corrs + inl;
                const uint32_t upper = early_drop_upper(tried, nsub, c_here);
                if ((float)upper < a.accept_bound) {  // scene.hpp:500-503
                    a.counts[h] = c_here;
                    a.dropped[h] = 1;
                    tested = tried;
                    open = false;
                    continue;
                }
            }
            if (a.probe) continue;  // the level itself is scored by the next stage
            corrs += cnt;
            if (L == EL_LEVELS - 1) {  // passed every checkpoint: the walk ends with the subset
                a.counts[h] = corrs;
                a.dropped[h] = 0;
                tested = nsub;
                open = false;
            }
        }
        if (open) a.corrs[h] = corrs;
        else a.alive[h] = 0;
    }
    const uint32_t irr = __ballot_sync(0xffffffffu, irregular);
    if (irr) {
        const int lane = threadIdx.x & 31, leader = __ffs(irr) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(a.n_irregular, (uint32_t)__popc(irr));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (irregular) a.irregular[base + __popc(irr & ((1u << lane) - 1u))] = h;
    }
    if (a.n_tests) {
#pragma unroll
        for (int d = 16; d; d >>= 1) tested += __shfl_xor_sync(0xffffffffu, tested, d);
        if ((threadIdx.x & 31) == 0 && tested) atomicAdd(a.n_tests, tested);
    }
}
void launch_el_eval(cudaStream_t st, const EvalArgs& a, uint32_t n_hyp_bound) {
    if (!n_hyp_bound) return;
    ++g_launch_count;
    el_eval_kernel<<<(n_hyp_bound + 255) / 256, 256, 0, st>>>(a);
}

// alive[h] = 1 for h < n_local, the list of alive hypotheses = all of them, per-hypothesis accumulators reset
// (lvl_cnt / minkey: EL_MAX_MERGE levels x cap)
__global__ void el_init_kernel(const uint32_t* __restrict__ n_local, uint32_t cap, uint8_t* __restrict__ alive,
                               uint32_t* __restrict__ hl, uint32_t* __restrict__ corrs, uint32_t* __restrict__ lvl_cnt,
                               uint32_t* __restrict__ minkey, uint8_t* __restrict__ dropped, uint32_t* __restrict__ counts) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= cap) return;
    alive[h] = h < *n_local ? 1 : 0;
    hl[h] = h;
    corrs[h] = 0u;
#pragma unroll
    for (int r = 0; r < EL_MAX_MERGE; ++r) {
        lvl_cnt[(size_t)r * cap + h] = 0u;
        minkey[(size_t)r * cap + h] = 0xffffffffu;
    }
    dropped[h] = 0;
    counts[h] = 0u;
}
void launch_el_init(cudaStream_t st, const uint32_t* n_local, uint32_t cap, uint8_t* alive, uint32_t* hl, uint32_t* corrs,
                    uint32_t* lvl_cnt, uint32_t* minkey, uint8_t* dropped, uint32_t* counts) {
    if (!cap) return;
    ++g_launch_count;
    el_init_kernel<<<(cap + 255) / 256, 256, 0, st>>>(n_local, cap, alive, hl, corrs, lvl_cnt, minkey, dropped, counts);
}

}  // namespace tmk
