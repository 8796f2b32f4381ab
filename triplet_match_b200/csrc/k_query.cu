// k_query.cu — device-side glue of the resident query (tm_query_run): shard
// range, per-outer hypothesis ranges, top-k selection for the ICP stage and the
// best-pose read-out.  Everything stays on the device so the whole pipeline is
// one stream-ordered sequence with no host round trip.
#include <algorithm>

#include <cstdlib>

#include "tm_kernels.cuh"

namespace tmk {

// shard = [h_begin, h_end) of the global hypothesis list owned by this rank
// bounds != null: the test-balanced split of balance_bounds_kernel ([bounds[rank], bounds[rank+1]))
__global__ void shard_range_kernel(const unsigned long long* __restrict__ hyp_off, uint64_t n_pairs,
                                   unsigned long long hyp_limit, uint32_t rank, uint32_t world,
                                   unsigned long long capacity, const unsigned long long* __restrict__ bounds,
                                   unsigned long long* shard, uint32_t* n_local, uint32_t* err) {
    unsigned long long H = hyp_off[n_pairs];
    if (hyp_limit && H > hyp_limit) H = hyp_limit;
    unsigned long long per = (H + world - 1) / world;
    unsigned long long hb = min((unsigned long long)rank * per, H);
    unsigned long long he = min(hb + per, H);
    if (bounds) {
        hb = min(bounds[rank], H);
        he = min(bounds[rank + 1], H);
    }
    if (he - hb > capacity) {
        *err = 1u;
        he = hb + capacity;
    }
    shard[0] = hb;
    shard[1] = he;
    shard[2] = H;
    *n_local = (uint32_t)(he - hb);
}
void launch_shard_range(cudaStream_t st, const unsigned long long* hyp_off, uint64_t n_pairs,
                        unsigned long long hyp_limit, uint32_t rank, uint32_t world,
                        unsigned long long capacity, const unsigned long long* bounds, unsigned long long* shard,
                        uint32_t* n_local, uint32_t* err) {
    ++g_launch_count;
    shard_range_kernel<<<1, 1, 0, st>>>(hyp_off, n_pairs, hyp_limit, rank, world, capacity, bounds, shard,
                                        n_local, err);
}

// ---- test-balanced shards -------------------------------------------------------------------------
// A shard's work is the number of hypothesis-point tests, |subset(outer)| x hypotheses(outer) summed over
// its outer samples, not its number of hypotheses.  balance_bounds_kernel cuts the global hypothesis list
// into `world` contiguous ranges of (nearly) equal tests: cum[o] = tests before outer sample o (block scan),
// rank r starts at the hypothesis where the running test count reaches total * r / world.  One CTA; every
// rank computes the same bounds from the same inputs.
__global__ void __launch_bounds__(1024)
    balance_bounds_kernel(const unsigned long long* __restrict__ hyp_off, uint64_t n_pairs,
                          unsigned long long hyp_limit, const uint32_t* __restrict__ outer_pair_off,
                          const uint32_t* __restrict__ sizes, uint32_t n_outer, uint32_t world,
                          unsigned long long* __restrict__ cum, unsigned long long* __restrict__ bounds) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    unsigned long long H = hyp_off[n_pairs];
    if (hyp_limit && H > hyp_limit) H = hyp_limit;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (uint32_t base = 0; base < n_outer; base += blockDim.x) {
        const uint32_t o = base + threadIdx.x;
        unsigned long long t = 0ull;
        if (o < n_outer) {
            const unsigned long long hb = min(hyp_off[outer_pair_off[o]], H), he = min(hyp_off[outer_pair_off[o + 1]], H);
            t = (he - hb) * (unsigned long long)sizes[o];
        }
        unsigned long long incl = t;  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long woff = 0ull;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        const unsigned long long carry = carry_s;
        if (o < n_outer) cum[o] = carry + woff + incl - t;  // exclusive
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + woff + incl;
        __syncthreads();
    }
    const unsigned long long total = carry_s;
    if (threadIdx.x == 0) cum[n_outer] = total;
    __syncthreads();
    for (uint32_t r = threadIdx.x; r <= world; r += blockDim.x) {
        unsigned long long b;
        if (r == 0) b = 0ull;
        else if (r == world || total == 0ull) b = (r == world) ? H : min(H, (H + world - 1) / world * r);
        else {
            // total * r / world without overflow: total < 2^63 / world for any realistic list; use 128-bit
            const unsigned long long hi = __umul64hi(total, (unsigned long long)r);
            const unsigned long long lo = total * (unsigned long long)r;
            unsigned long long target = hi ? ~0ull : lo / world;
            // the last outer sample o with cum[o] <= target
            uint32_t a = 0, e = n_outer;  // cum[a] <= target < cum[e] (cum[n_outer] = total > target)
            while (e - a > 1) {
                const uint32_t mid = (a + e) >> 1;
                if (cum[mid] <= target) a = mid; else e = mid;
            }
            const unsigned long long hb = min(hyp_off[outer_pair_off[a]], H), he = min(hyp_off[outer_pair_off[a + 1]], H);
            const unsigned long long sz = sizes[a];
            b = sz ? min(he, hb + (target - cum[a]) / sz) : hb;
        }
        bounds[r] = b;
    }
}
void launch_balance_bounds(cudaStream_t st, const unsigned long long* hyp_off, uint64_t n_pairs,
                           unsigned long long hyp_limit, const uint32_t* outer_pair_off, const uint32_t* sizes,
                           uint32_t n_outer, uint32_t world, unsigned long long* cum, unsigned long long* bounds) {
    ++g_launch_count;
    balance_bounds_kernel<<<1, 1024, 0, st>>>(hyp_off, n_pairs, hyp_limit, outer_pair_off, sizes, n_outer, world, cum,
                                              bounds);
}

// g_hyp[g] = local index of the first hypothesis of outer sample g (n_outer+1 entries)
__global__ void group_hyp_ranges_kernel(const unsigned long long* __restrict__ hyp_off,
                                        const uint32_t* __restrict__ outer_pair_off,
                                        uint32_t n_outer,
                                        const unsigned long long* __restrict__ shard,
                                        uint32_t* __restrict__ g_hyp) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_outer) return;
    unsigned long long h = hyp_off[outer_pair_off[g]];
    unsigned long long hb = shard[0], he = shard[1];
    h = h < hb ? hb : (h > he ? he : h);
    g_hyp[g] = (uint32_t)(h - hb);
}
void launch_group_hyp_ranges(cudaStream_t st, const unsigned long long* hyp_off,
                             const uint32_t* outer_pair_off, uint32_t n_outer,
                             const unsigned long long* shard, uint32_t* g_hyp) {
    ++g_launch_count;
    group_hyp_ranges_kernel<<<(n_outer + 1 + 127) / 128, 128, 0, st>>>(hyp_off, outer_pair_off,
                                                                       n_outer, shard, g_hyp);
}

__global__ void group_of_hyp_kernel(const uint32_t* __restrict__ g_hyp, uint32_t n_groups,
                                    uint32_t* __restrict__ g_of_hyp) {
    uint32_t g = blockIdx.x;
    if (g >= n_groups) return;
    for (uint32_t h = g_hyp[g] + threadIdx.x; h < g_hyp[g + 1]; h += blockDim.x) g_of_hyp[h] = g;
}
void launch_group_of_hyp(cudaStream_t st, const uint32_t* g_hyp, uint32_t n_groups,
                         uint32_t* g_of_hyp) {
    if (!n_groups) return;
    ++g_launch_count;
    group_of_hyp_kernel<<<n_groups, 128, 0, st>>>(g_hyp, n_groups, g_of_hyp);
}

// top-k by (inliers desc, id asc).  Keys (count << 32 | ~id) are unique, so "the largest key below
// the previous one" enumerates them in order.  Two levels: every CTA takes a slice of TOPK_SLICE
// hypotheses into shared memory and emits its own top k (phase A, whole GPU); one CTA then picks
// the global top k from those candidates (phase B).  The global top k is a subset of the union
// of the slice top k's, so the result equals the single-pass selection.
constexpr uint32_t TOPK_SLICE = 4096;
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* wbest) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o > v ? o : v;
    }
    __syncthreads();  // previous readers of wbest are done
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = wbest[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = wbest[w] > r ? wbest[w] : r;
    return r;
}
__device__ __forceinline__ void topk_slice_body(const uint32_t* __restrict__ counts, const uint8_t* __restrict__ valid,
                                                const uint8_t* __restrict__ excluded, const uint32_t* __restrict__ n_local,
                                                uint32_t k, unsigned long long* __restrict__ cand) {
    __shared__ unsigned long long keys[TOPK_SLICE];
    __shared__ unsigned long long wbest[32];
    const uint32_t n = *n_local;
    const uint32_t base = blockIdx.x * TOPK_SLICE;
    for (uint32_t t = threadIdx.x; t < TOPK_SLICE; t += blockDim.x) {
        const uint32_t i = base + t;
        unsigned long long key = 0;
        if (i < n && (!valid || valid[i]) && (!excluded || !excluded[i]) && counts[i])
            key = ((unsigned long long)counts[i] << 32) | (unsigned long long)(0xFFFFFFFFu - i);
        keys[t] = key;
    }
    __syncthreads();
    unsigned long long last = ~0ull;
    for (uint32_t r = 0; r < k; ++r) {
        unsigned long long best = 0;
        for (uint32_t t = threadIdx.x; t < TOPK_SLICE; t += blockDim.x) {
            const unsigned long long key = keys[t];
            if (key < last && key > best) best = key;
        }
        best = block_max_u64(best, wbest);
        if (threadIdx.x == 0) cand[(size_t)blockIdx.x * k + r] = best;
        last = best;
        if (best == 0ull) {  // slice exhausted (uniform across the CTA)
            for (uint32_t rr = r + 1 + threadIdx.x; rr < k; rr += blockDim.x) cand[(size_t)blockIdx.x * k + rr] = 0ull;
            break;
        }
    }
}
__device__ __forceinline__ void topk_final_body(const unsigned long long* __restrict__ cand, uint32_t n_cand, uint32_t k,
                                                uint32_t* __restrict__ topk_ids) {
    __shared__ unsigned long long wbest[32];
    unsigned long long last = ~0ull;
    for (uint32_t r = 0; r < k; ++r) {
        unsigned long long best = 0;
        for (uint32_t i = threadIdx.x; i < n_cand; i += blockDim.x) {
            const unsigned long long key = cand[i];
            if (key < last && key > best) best = key;
        }
        best = block_max_u64(best, wbest);
        // best == 0 only when no candidate is left (a real key always has ~id bits set
        // unless id == 0xFFFFFFFF, which never occurs for n < 2^32 - 1)
        if (threadIdx.x == 0) topk_ids[r] = best ? (0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull)) : 0xFFFFFFFFu;
        last = best;
        if (best == 0ull) {
            for (uint32_t rr = r + 1 + threadIdx.x; rr < k; rr += blockDim.x) topk_ids[rr] = 0xFFFFFFFFu;
            break;
        }
    }
}
// ---- fast path: threshold + compaction -----------------------------------------------------------------------------
// The k-th largest of the slice maxima is a lower bound T0 of the k-th largest key overall (those maxima are k keys
// >= T0), so the top k are among the keys >= T0 — usually a few hundred of the 2^20.  They are compacted into a short
// list and ranked there.  When the list would overflow (many equal counts) a flag routes the selection through the
// slice kernels above instead; both paths select the same k keys.
constexpr uint32_t TOPK_LIST = 4096;  // 32 KB of static shared memory in the rank kernel
__device__ __forceinline__ unsigned long long topk_key(const uint32_t* counts, const uint8_t* valid, const uint8_t* excluded,
                                                       uint32_t i, uint32_t n) {
    if (i < n && (!valid || valid[i]) && (!excluded || !excluded[i]) && counts[i])
        return ((unsigned long long)counts[i] << 32) | (unsigned long long)(0xFFFFFFFFu - i);
    return 0ull;
}
__global__ void __launch_bounds__(1024)
    topk_slice_max_kernel(const uint32_t* __restrict__ counts, const uint8_t* __restrict__ valid,
                          const uint8_t* __restrict__ excluded, const uint32_t* __restrict__ n_local,
                          unsigned long long* __restrict__ slice_max) {
    __shared__ unsigned long long wbest[32];
    const uint32_t n = *n_local, base = blockIdx.x * TOPK_SLICE;
    unsigned long long best = 0;
    for (uint32_t t = threadIdx.x; t < TOPK_SLICE; t += blockDim.x) {
        const unsigned long long key = topk_key(counts, valid, excluded, base + t, n);
        best = key > best ? key : best;
    }
    best = block_max_u64(best, wbest);
    if (threadIdx.x == 0) slice_max[blockIdx.x] = best;
}
// ctrl[0] = T0 (0 when fewer than k slices hold a key), ctrl[1] = list length, ctrl[2] = overflow flag
__global__ void __launch_bounds__(1024)
    topk_threshold_kernel(const unsigned long long* __restrict__ slice_max, uint32_t n_slices, uint32_t k,
                          unsigned long long* __restrict__ ctrl) {
    if (threadIdx.x == 0) { ctrl[0] = 0ull; ctrl[1] = 0ull; ctrl[2] = 0ull; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_slices; i += blockDim.x) {
        const unsigned long long v = slice_max[i];
        if (!v) continue;
        uint32_t rank = 0;  // keys are unique: exactly one maximum has rank k - 1
        for (uint32_t j = 0; j < n_slices; ++j) rank += slice_max[j] > v ? 1u : 0u;
        if (rank == k - 1u) ctrl[0] = v;
    }
}
__global__ void __launch_bounds__(256)
    topk_compact_kernel(const uint32_t* __restrict__ counts, const uint8_t* __restrict__ valid,
                        const uint8_t* __restrict__ excluded, const uint32_t* __restrict__ n_local,
                        unsigned long long* __restrict__ ctrl, unsigned long long* __restrict__ list) {
    const uint32_t n = *n_local;
    const unsigned long long T0 = ctrl[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = topk_key(counts, valid, excluded, i, n);
        if (key && key >= T0) {
            const unsigned long long at = atomicAdd(&ctrl[1], 1ull);
            if (at < TOPK_LIST) list[at] = key;
            else ctrl[2] = 1ull;
        }
    }
}
// rank the compacted keys; rank r < k goes to topk_ids[r]
__global__ void __launch_bounds__(1024)
    topk_rank_kernel(const unsigned long long* __restrict__ ctrl, const unsigned long long* __restrict__ list, uint32_t k,
                     uint32_t* __restrict__ topk_ids) {
    if (ctrl[2]) return;  // overflow: the slice kernels produce the result
    __shared__ unsigned long long keys[TOPK_LIST];
    const uint32_t m = (uint32_t)ctrl[1];
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) keys[i] = list[i];
    for (uint32_t r = threadIdx.x; r < k; r += blockDim.x) topk_ids[r] = 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned long long v = keys[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < m; ++j) rank += keys[j] > v ? 1u : 0u;
        if (rank < k) topk_ids[rank] = 0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFull);
    }
}
// the slice path, only when the list overflowed (uniform early exit otherwise)
__global__ void __launch_bounds__(1024)
    topk_slice_guarded_kernel(const unsigned long long* __restrict__ ctrl, const uint32_t* __restrict__ counts,
                              const uint8_t* __restrict__ valid, const uint8_t* __restrict__ excluded,
                              const uint32_t* __restrict__ n_local, uint32_t k, unsigned long long* __restrict__ cand,
                              int force);
__global__ void __launch_bounds__(1024)
    topk_final_guarded_kernel(const unsigned long long* __restrict__ ctrl, const unsigned long long* __restrict__ cand,
                              uint32_t n_cand, uint32_t k, uint32_t* __restrict__ topk_ids, int force);
static bool topk_force_slices() {  // TM_TOPK_SLICES=1: development knob, read once — always select through the slice kernels
    static const bool v = [] {
        const char* e = getenv("TM_TOPK_SLICES");
        return e && atoi(e) != 0;
    }();
    return v;
}

// capacity = upper bound of *n_local (sizes the grid); scratch_keys: topk_scratch_bytes(capacity, k)
// layout: [slices * k] slice candidates | [slices] slice maxima | [TOPK_LIST] list | [4] ctrl
size_t topk_scratch_bytes(uint64_t capacity, uint32_t k) {
    const size_t slices = (size_t)((capacity + TOPK_SLICE - 1) / TOPK_SLICE);
    return (slices * k + slices + TOPK_LIST + 4) * 8 + 8;
}
void launch_select_topk(cudaStream_t st, const uint32_t* counts, const uint8_t* valid, const uint8_t* excluded,
                        const uint32_t* n_local, uint64_t capacity, uint32_t k, uint32_t* topk_ids,
                        unsigned long long* scratch_keys) {
    if (!k) return;
    const uint32_t slices = (uint32_t)std::max<uint64_t>(1, (capacity + TOPK_SLICE - 1) / TOPK_SLICE);
    unsigned long long* cand = scratch_keys;
    unsigned long long* slice_max = cand + (size_t)slices * k;
    unsigned long long* list = slice_max + slices;
    unsigned long long* ctrl = list + TOPK_LIST;
    g_launch_count += 6;
    topk_slice_max_kernel<<<slices, 1024, 0, st>>>(counts, valid, excluded, n_local, slice_max);
    topk_threshold_kernel<<<1, 1024, 0, st>>>(slice_max, slices, k, ctrl);
    topk_compact_kernel<<<std::min<uint32_t>(slices * 4u, 1184u), 256, 0, st>>>(counts, valid, excluded, n_local, ctrl, list);
    topk_rank_kernel<<<1, 1024, 0, st>>>(ctrl, list, k, topk_ids);
    const int force = topk_force_slices() ? 1 : 0;
    topk_slice_guarded_kernel<<<slices, 1024, 0, st>>>(ctrl, counts, valid, excluded, n_local, k, cand, force);
    topk_final_guarded_kernel<<<1, 1024, 0, st>>>(ctrl, cand, slices * k, k, topk_ids, force);
}

__global__ void __launch_bounds__(1024)
    topk_slice_guarded_kernel(const unsigned long long* __restrict__ ctrl, const uint32_t* __restrict__ counts,
                              const uint8_t* __restrict__ valid, const uint8_t* __restrict__ excluded,
                              const uint32_t* __restrict__ n_local, uint32_t k, unsigned long long* __restrict__ cand,
                              int force) {
    if (!ctrl[2] && !force) return;
    topk_slice_body(counts, valid, excluded, n_local, k, cand);
}
__global__ void __launch_bounds__(1024)
    topk_final_guarded_kernel(const unsigned long long* __restrict__ ctrl, const unsigned long long* __restrict__ cand,
                              uint32_t n_cand, uint32_t k, uint32_t* __restrict__ topk_ids, int force) {
    if (!ctrl[2] && !force) return;
    topk_final_body(cand, n_cand, k, topk_ids);
}

__global__ void gather_rows_kernel(const float4* __restrict__ T, const uint32_t* __restrict__ ids,
                                   uint32_t k, float4* __restrict__ out,
                                   uint32_t* __restrict__ active) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    uint32_t id = ids[r];
    if (id == 0xFFFFFFFFu) {
        const float nanv = __int_as_float(0x7fc00000);
        out[3 * r] = out[3 * r + 1] = out[3 * r + 2] = make_float4(nanv, nanv, nanv, nanv);
        active[r] = 0;
    } else {
        out[3 * r] = T[3 * (size_t)id];
        out[3 * r + 1] = T[3 * (size_t)id + 1];
        out[3 * r + 2] = T[3 * (size_t)id + 2];
        active[r] = 1;
    }
}
void launch_gather_rows(cudaStream_t st, const float4* T, const uint32_t* ids, uint32_t k,
                        float4* out, uint32_t* active) {
    if (!k) return;
    ++g_launch_count;
    gather_rows_kernel<<<(k + 63) / 64, 64, 0, st>>>(T, ids, k, out, active);
}

// after the (all-)reduce: if the winning hypothesis lives in this shard, export
// its pose (column-major) and normalised score
__global__ void finalize_best_kernel(const unsigned long long* __restrict__ best,
                                     const unsigned long long* __restrict__ shard,
                                     const float4* __restrict__ T,
                                     const unsigned long long* __restrict__ scores,
                                     const unsigned long long* __restrict__ lazy_acc,
                                     uint32_t model_n, float* __restrict__ best_T16,
                                     double* __restrict__ best_score) {
    unsigned long long key = *best;
    if (!key) return;
    unsigned long long gid = 0xFFFFFFFFull - (key & 0xFFFFFFFFull);
    if (gid < shard[0] || gid >= shard[1]) return;
    size_t l = (size_t)(gid - shard[0]);
    for (int r = 0; r < 3; ++r) {
        float4 v = T[3 * l + r];
        best_T16[r] = v.x;
        best_T16[4 + r] = v.y;
        best_T16[8 + r] = v.z;
        best_T16[12 + r] = v.w;
    }
    best_T16[3] = best_T16[7] = best_T16[11] = 0.f;
    best_T16[15] = 1.f;
    // lazy_acc: the selected pose's score summed by score_best_kernel (count-only bulk scorer)
    if (lazy_acc) *best_score = (double)lazy_acc[0] / SCORE_SCALE / (double)model_n;
    else if (scores) *best_score = (double)scores[l] / SCORE_SCALE / (double)model_n;
}
void launch_finalize_best(cudaStream_t st, const unsigned long long* best,
                          const unsigned long long* shard, const float4* T,
                          const unsigned long long* scores, const unsigned long long* lazy_acc, uint32_t model_n,
                          float* best_T16, double* best_score) {
    ++g_launch_count;
    finalize_best_kernel<<<1, 1, 0, st>>>(best, shard, T, scores, lazy_acc, model_n, best_T16, best_score);
}

}  // namespace tmk
