// k_util.cu — layout / utility kernels of the search path:
//   pack_cloud      strided host layout (packed xyz or pcl::PointSurfel AoS) -> SoA float4 + flags
//   set_mask        mask_ updates (include/impl/scene.hpp:87-90)
//   exclusive scan  CSR offsets (hit counts, subset counts)
//   ball subsets    radius subset of find_in_subset (include/impl/scene.hpp:273)
//   voxel_fill      exact 1-NN grid fill (include/impl/model.hpp:81-94)
//   traits_project  per-point closed forms (cylinder/plane/plane2/identity _traits::project)
#include <algorithm>

#include "tm_kernels.cuh"
#include "../../include/triplet_match/tm_voxel_centre.h"

namespace tmk {

// ------------------------------------------------------------------ pack_cloud
// raw: the three strided arrays were copied verbatim into one device buffer
// (n*stride floats each, or one shared AoS buffer).  flags: per-point byte
// (scene: tangent_mask_) or null (model: computed ||tangent|| > 0.7,
// include/impl/scene.hpp:470).
__global__ void pack_cloud_kernel(const float* __restrict__ pos, const float* __restrict__ nrm,
                                  const float* __restrict__ tgt, uint32_t stride, uint32_t n,
                                  const uint8_t* __restrict__ flags, int model_mode,
                                  float4* __restrict__ opos, float4* __restrict__ onrm,
                                  float4* __restrict__ otgt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t b = (size_t)i * stride;
    f3 p = {pos[b], pos[b + 1], pos[b + 2]};
    f3 nn = {nrm[b], nrm[b + 1], nrm[b + 2]};
    f3 t = {tgt[b], tgt[b + 1], tgt[b + 2]};
    uint32_t fl = 0;
    if (model_mode) {
        if (norm3(t) > 0.7f) fl |= FLAG_TANGENT;
    } else if (flags && flags[i]) {
        fl |= FLAG_TANGENT;
    }
    opos[i] = make_float4(p.x, p.y, p.z, __uint_as_float(fl));
    onrm[i] = make_float4(nn.x, nn.y, nn.z, 0.f);
    otgt[i] = make_float4(t.x, t.y, t.z, 0.f);
}

void launch_pack_cloud(cudaStream_t st, const float* pos, const float* nrm, const float* tgt,
                       uint32_t stride, uint32_t n, const uint8_t* flags, int model_mode,
                       float4* opos, float4* onrm, float4* otgt) {
    if (!n) return;
    ++g_launch_count;
    pack_cloud_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, nrm, tgt, stride, n, flags, model_mode,
                                                       opos, onrm, otgt);
}

__global__ void set_mask_kernel(float4* __restrict__ pos, uint32_t n,
                                const uint8_t* __restrict__ mask) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t fl = __float_as_uint(pos[i].w) & ~FLAG_MASKED;
    if (mask && mask[i]) fl |= FLAG_MASKED;
    pos[i].w = __uint_as_float(fl);
}
void launch_set_mask(cudaStream_t st, float4* pos, uint32_t n, const uint8_t* mask) {
    if (!n) return;
    ++g_launch_count;
    set_mask_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos, n, mask);
}

// ------------------------------------------------------------ exclusive scan
// Single-CTA tiled scan with carry (inputs here are at most a few million
// counters; the launch is latency-, not bandwidth-critical).  out has n+1
// entries; out[n] = total.
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
    exclusive_scan_kernel(const uint32_t* __restrict__ in, TOut* __restrict__ out, uint64_t n) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += (uint64_t)SCAN_THREADS * SCAN_ITEMS) {
        uint64_t i0 = base + (uint64_t)threadIdx.x * SCAN_ITEMS;
        unsigned long long v[SCAN_ITEMS], tsum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
            tsum += v[k];
        }
        unsigned long long incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sums[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sums[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        unsigned long long carry = carry_s;
        unsigned long long excl = carry + warp_sums[warp] + (incl - tsum);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (i0 + k < n) out[i0 + k] = (TOut)excl;
            excl += v[k];
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl;  // total so far
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = (TOut)carry_s;
}
// The same scan over many CTAs for long inputs (the hit counts of an 8-GPU recorded list: 2.6e5 pairs, 64 rounds
// of the single-CTA loop above).  Chained scan: a CTA takes the next tile by ticket (so tiles start in order and a
// waiting CTA's predecessor is always running or done), scans its 4096 items, waits for the predecessor's inclusive
// total, publishes its own.  scratch: [0] ticket, [1] unused, then per tile {ready flag, inclusive total}.
template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
    chained_scan_kernel(const uint32_t* __restrict__ in, TOut* __restrict__ out, uint64_t n,
                        unsigned long long* __restrict__ scratch) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry_s;
    __shared__ uint32_t tile_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = (uint32_t)((n + (uint64_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((uint64_t)SCAN_THREADS * SCAN_ITEMS));
    volatile unsigned long long* flag = scratch + 2;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) tile_s = (uint32_t)atomicAdd(&scratch[0], 1ull);
        __syncthreads();
        const uint32_t tile = tile_s;
        if (tile >= n_tiles) break;
        const uint64_t i0 = (uint64_t)tile * SCAN_THREADS * SCAN_ITEMS + (uint64_t)threadIdx.x * SCAN_ITEMS;
        unsigned long long v[SCAN_ITEMS], tsum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
            tsum += v[k];
        }
        unsigned long long incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sums[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sums[lane] = wi - w;  // exclusive
            if (lane == 31) {          // wi = the tile's total: chain it
                unsigned long long prev = 0ull;
                if (tile) {
                    while (flag[2 * (tile - 1)] == 0ull) {}
                    __threadfence();
                    prev = flag[2 * (tile - 1) + 1];
                }
                carry_s = prev;
                flag[2 * tile + 1] = prev + wi;
                __threadfence();
                flag[2 * tile] = 1ull;
                if (tile == n_tiles - 1) out[n] = (TOut)(prev + wi);
            }
        }
        __syncthreads();
        unsigned long long excl = carry_s + warp_sums[warp] + (incl - tsum);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (i0 + k < n) out[i0 + k] = (TOut)excl;
            excl += v[k];
        }
    }
}
size_t scan_scratch_bytes(uint64_t n) {
    const uint64_t tiles = (n + (uint64_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((uint64_t)SCAN_THREADS * SCAN_ITEMS);
    return (size_t)(2 + 2 * tiles) * 8;
}
// scratch == nullptr or a short input: the single-CTA kernel
void launch_exclusive_scan_u64_chained(cudaStream_t st, const uint32_t* in, unsigned long long* out, uint64_t n,
                                       unsigned long long* scratch, int max_ctas) {
    if (!scratch || n <= (uint64_t)SCAN_THREADS * SCAN_ITEMS * 2) {
        launch_exclusive_scan_u64(st, in, out, n);
        return;
    }
    ++g_launch_count;
    cudaMemsetAsync(scratch, 0, scan_scratch_bytes(n), st);
    const uint64_t tiles = (n + (uint64_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((uint64_t)SCAN_THREADS * SCAN_ITEMS);
    const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)std::max(1, max_ctas));
    chained_scan_kernel<unsigned long long><<<grid, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
}
void launch_exclusive_scan_u32_chained(cudaStream_t st, const uint32_t* in, uint32_t* out, uint64_t n,
                                       unsigned long long* scratch, int max_ctas) {
    if (!scratch || n <= (uint64_t)SCAN_THREADS * SCAN_ITEMS * 2) {
        launch_exclusive_scan_u32(st, in, out, n);
        return;
    }
    ++g_launch_count;
    cudaMemsetAsync(scratch, 0, scan_scratch_bytes(n), st);
    const uint64_t tiles = (n + (uint64_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((uint64_t)SCAN_THREADS * SCAN_ITEMS);
    const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)std::max(1, max_ctas));
    chained_scan_kernel<uint32_t><<<grid, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
}
void launch_exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint64_t n) {
    ++g_launch_count;
    exclusive_scan_kernel<uint32_t><<<1, SCAN_THREADS, 0, st>>>(in, out, n);
}
void launch_exclusive_scan_u64(cudaStream_t st, const uint32_t* in, unsigned long long* out,
                               uint64_t n) {
    ++g_launch_count;
    exclusive_scan_kernel<unsigned long long><<<1, SCAN_THREADS, 0, st>>>(in, out, n);
}

// -------------------------------------------------------------- ball subsets
// One warp owns a segment of BALL_SEG consecutive scene points; lane l visits points
// seg*BALL_SEG + k*32 + l, so ballot order is ascending index order and the compaction is
// deterministic.  Predicate (FLANN L2_Simple order): (dx*dx + dy*dy) + dz*dz < r^2.
// Centres are walked 32 at a time: lane l decides whether centre c0+l can touch the segment at
// all — the centre is active (has hypotheses in this rank's shard) and the segment's bounding
// box (computed once per scene, seg_bbox_kernel) comes within r of it, with a 1e-5 relative
// guard that dominates the rounding of the exact predicate — and only the surviving
// (segment, centre) pairs run the per-point loop.  On Morton-ordered scenes a ball touches a
// few percent of the segments.  counts layout: [centre][segment], zero-initialised by the host.
__global__ void __launch_bounds__(256)
    seg_bbox_kernel(const float4* __restrict__ pos, uint32_t n, uint32_t n_seg, float4* __restrict__ lo,
                    float4* __restrict__ hi) {
    const int lane = threadIdx.x & 31;
    const uint32_t seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= n_seg) return;
    const uint32_t base = seg * BALL_SEG;
    float mnx = 3.0e38f, mny = 3.0e38f, mnz = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f, mxz = -3.0e38f;
    for (uint32_t k = 0; k < BALL_SEG / 32; ++k) {
        const uint32_t i = base + k * 32 + lane;
        if (i < n) {
            const float4 p = pos[i];
            if (p.x == p.x && p.y == p.y && p.z == p.z) {  // NaN points never pass the predicate
                mnx = fminf(mnx, p.x); mxx = fmaxf(mxx, p.x);
                mny = fminf(mny, p.y); mxy = fmaxf(mxy, p.y);
                mnz = fminf(mnz, p.z); mxz = fmaxf(mxz, p.z);
            }
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
        lo[seg] = make_float4(mnx, mny, mnz, 0.f);
        hi[seg] = make_float4(mxx, mxy, mxz, 0.f);
    }
}
void launch_seg_bbox(cudaStream_t st, const float4* pos, uint32_t n, float4* lo, float4* hi) {
    const uint32_t n_seg = (n + BALL_SEG - 1) / BALL_SEG;
    if (!n_seg) return;
    ++g_launch_count;
    seg_bbox_kernel<<<(n_seg + 7) / 8, 256, 0, st>>>(pos, n, n_seg, lo, hi);
}

template <bool FILL>
__global__ void __launch_bounds__(256)
    ball_kernel(const float4* __restrict__ pos, uint32_t n, const float4* __restrict__ seg_lo,
                const float4* __restrict__ seg_hi, const uint32_t* __restrict__ centres, uint32_t n_centres,
                const uint32_t* __restrict__ active_ranges, float r2, uint32_t n_seg,
                uint32_t* __restrict__ counts, const unsigned long long* __restrict__ row_off,
                int32_t* __restrict__ indices) {
    const int lane = threadIdx.x & 31;
    const uint32_t seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= n_seg) return;
    const uint32_t base = seg * BALL_SEG;
    float4 blo = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f), bhi = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f);
    if (seg_lo) {
        blo = seg_lo[seg];
        bhi = seg_hi[seg];
    }
    // blockIdx.y selects a batch of 32 centres: (segment, batch) pairs spread over the whole GPU
    {
        const uint32_t c0 = blockIdx.y * 32u;
        const uint32_t cl = c0 + lane;
        bool cand = false;
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cl < n_centres && (!active_ranges || active_ranges[cl + 1] > active_ranges[cl])) {
            mine = pos[centres[cl]];
            // squared distance from the centre to the segment box (0 inside); NaN never culls
            const float dx = fmaxf(fmaxf(blo.x - mine.x, mine.x - bhi.x), 0.f);
            const float dy = fmaxf(fmaxf(blo.y - mine.y, mine.y - bhi.y), 0.f);
            const float dz = fmaxf(fmaxf(blo.z - mine.z, mine.z - bhi.z), 0.f);
            const float d2 = dx * dx + dy * dy + dz * dz;
            cand = !(d2 > r2 * 1.00002f + 1e-30f);
        }
        uint32_t todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t c = c0 + src;
            const float cx = __shfl_sync(0xffffffffu, mine.x, src), cy = __shfl_sync(0xffffffffu, mine.y, src),
                        cz = __shfl_sync(0xffffffffu, mine.z, src);
            unsigned long long off = 0ull;
            if (FILL) off = row_off[c] + counts[(size_t)c * n_seg + seg];  // counts hold in-row offsets here
            uint32_t cnt = 0;
#pragma unroll 4
            for (uint32_t k = 0; k < BALL_SEG / 32; ++k) {
                const uint32_t i = base + k * 32 + lane;
                bool in = false;
                if (i < n) {
                    const float4 p = pos[i];
                    const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
                    in = ((dx * dx + dy * dy) + dz * dz) < r2;
                }
                const uint32_t b = __ballot_sync(0xffffffffu, in);
                if (FILL) {
                    if (in) indices[off + cnt + __popc(b & ((1u << lane) - 1u))] = (int32_t)i;
                }
                cnt += __popc(b);
            }
            if (!FILL && lane == 0 && cnt) counts[(size_t)c * n_seg + seg] = cnt;
        }
    }
}
void launch_ball_count(cudaStream_t st, const CloudDev& scene, const uint32_t* centres, uint32_t n_centres,
                       const uint32_t* active_ranges, float r2, uint32_t n_seg, uint32_t* counts) {
    if (!n_seg || !n_centres) return;
    cudaMemsetAsync(counts, 0, (size_t)n_centres * n_seg * 4, st);
    ++g_launch_count;
    ball_kernel<false><<<dim3((n_seg + 7) / 8, (n_centres + 31) / 32), 256, 0, st>>>(scene.pos, scene.n, scene.seg_lo, scene.seg_hi, centres,
                                                       n_centres, active_ranges, r2, n_seg, counts, nullptr, nullptr);
}
void launch_ball_fill(cudaStream_t st, const CloudDev& scene, const uint32_t* centres, uint32_t n_centres,
                      const uint32_t* active_ranges, float r2, uint32_t n_seg, const uint32_t* seg_local_off,
                      const unsigned long long* row_off, int32_t* indices) {
    if (!n_seg || !n_centres) return;
    ++g_launch_count;
    ball_kernel<true><<<dim3((n_seg + 7) / 8, (n_centres + 31) / 32), 256, 0, st>>>(scene.pos, scene.n, scene.seg_lo, scene.seg_hi, centres,
                                                      n_centres, active_ranges, r2, n_seg,
                                                      const_cast<uint32_t*>(seg_local_off), row_off, indices);
}
// per-centre exclusive scan of the segment counts, in place (one CTA per centre), + row totals
__global__ void __launch_bounds__(1024)
    ball_seg_scan_kernel(uint32_t* __restrict__ counts, uint32_t n_seg, uint32_t* __restrict__ row_total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* row = counts + (size_t)blockIdx.x * n_seg;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n_seg; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_seg ? row[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = warp_sums[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = carry_s + warp_sums[warp] + (incl - v);
        if (i < n_seg) row[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_total[blockIdx.x] = carry_s;
}
void launch_ball_seg_scan(cudaStream_t st, uint32_t* counts, uint32_t n_centres, uint32_t n_seg,
                          uint32_t* row_total) {
    if (!n_centres) return;
    ++g_launch_count;
    ball_seg_scan_kernel<<<n_centres, 1024, 0, st>>>(counts, n_seg, row_total);
}

// ---------------------------------------------------------------- voxel_fill
// One thread per voxel; model points streamed through shared memory in tiles.
// centre = inverse(to_voxel_) * (i,j,k,1) as the reference's Matrix4f::inverse() leaves it
// (tm_voxel_centre.h: a*index + b per axis); squared distance (dx*dx + dy*dy) + dz*dz; the
// lowest point index wins ties (strict '<' while scanning ascending).
constexpr int VF_TILE = 1024;
__global__ void __launch_bounds__(256)
    voxel_fill_kernel(const float4* __restrict__ mpos, uint32_t n, int ex, int ey, int ez, tm_centre_map cm,
                      uint32_t* __restrict__ voxel) {
    __shared__ float4 tile[VF_TILE];
    const size_t total = (size_t)ex * ey * ez;
    size_t lin = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = lin < total;
    int i = 0, j = 0, k = 0;
    if (live) {
        i = (int)(lin % ex);
        j = (int)((lin / ex) % ey);
        k = (int)(lin / ((size_t)ex * ey));
    }
    float qx = tm_voxel_centre(cm.a[0], cm.b[0], i), qy = tm_voxel_centre(cm.a[1], cm.b[1], j),
          qz = tm_voxel_centre(cm.a[2], cm.b[2], k);
    float best = 3.402823466e+38f;
    uint32_t bi = 0;
    for (uint32_t base = 0; base < n; base += VF_TILE) {
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < VF_TILE; t += blockDim.x)
            if (base + t < n) tile[t] = mpos[base + t];
        __syncthreads();
        uint32_t m = min((uint32_t)VF_TILE, n - base);
        for (uint32_t t = 0; t < m; ++t) {
            float4 p = tile[t];
            float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            float d = (dx * dx + dy * dy) + dz * dz;
            if (d < best) {
                best = d;
                bi = base + t;
            }
        }
    }
    if (live) voxel[lin] = bi;
}
// Pruned fill for larger grids.  Cells are grouped in 8x8x8 blocks.  Pass A: the distance d_b from
// every block centre c_b to its nearest model point (one warp per block, brute force).  For a cell
// centre q of the block, with h = the largest |q - c_b|, the nearest point of q is no farther than
// |q - p_b| <= h + d_b, so every point that can be nearest to (or tie at) any cell of the block
// lies within R = 2h + d_b of c_b.  Pass B: one CTA per block streams the model through shared
// memory, keeps (order-preserving ballot compaction) the points inside that ball — with a 1e-4
// relative guard that dominates every float rounding — and each thread runs the exact reference
// scan over the survivors for its two cells: same arithmetic, same ascending order, same strict
// '<', hence the same winners as the brute-force kernel.
constexpr int VB = 8;  // block edge in cells
__global__ void __launch_bounds__(256)
    voxel_block_radius_kernel(const float4* __restrict__ mpos, uint32_t n, int ex, int ey, int ez, float sx, float sy,
                              float sz, float tx, float ty, float tz, int bx, int by, int bz,
                              float* __restrict__ radius2) {
    const int lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nb = (uint32_t)bx * by * bz;
    if (b >= nb) return;
    const int bi = b % bx, bj = (b / bx) % by, bk = b / (bx * by);
    // centre of the block's lattice of cell centres (cells bi*8 .. bi*8+7, clipped to the grid)
    const float i0 = (float)(bi * VB), i1 = (float)min(bi * VB + VB - 1, ex - 1);
    const float j0 = (float)(bj * VB), j1 = (float)min(bj * VB + VB - 1, ey - 1);
    const float k0 = (float)(bk * VB), k1 = (float)min(bk * VB + VB - 1, ez - 1);
    const float cx = (0.5f * (i0 + i1) - tx) / sx, cy = (0.5f * (j0 + j1) - ty) / sy, cz = (0.5f * (k0 + k1) - tz) / sz;
    const float hx = 0.5f * (i1 - i0) / sx, hy = 0.5f * (j1 - j0) / sy, hz = 0.5f * (k1 - k0) / sz;
    float best = 3.0e38f;
    for (uint32_t t = lane; t < n; t += 32) {
        const float4 p = mpos[t];
        const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best) best = d;  // NaN never wins
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 0) {
        const float h = sqrtf(hx * hx + hy * hy + hz * hz);
        const float R = (2.f * h + sqrtf(best)) * 1.0001f + 1e-30f;
        radius2[b] = best < 3.0e38f ? R * R : 3.0e38f;  // no finite point: keep everything
    }
}
constexpr int VF2_TILE = 1024;
__global__ void __launch_bounds__(256)
    voxel_fill_pruned_kernel(const float4* __restrict__ mpos, uint32_t n, int ex, int ey, int ez, float sx, float sy,
                             float sz, float tx, float ty, float tz, tm_centre_map cm, int bx, int by,
                             const float* __restrict__ radius2, uint32_t* __restrict__ voxel) {
    __shared__ float4 cand[VF2_TILE];  // xyz + index bits in w
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t n_cand_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x;
    const int bi = b % bx, bj = (b / bx) % by, bk = b / (bx * by);
    const float i0 = (float)(bi * VB), i1 = (float)min(bi * VB + VB - 1, ex - 1);
    const float j0 = (float)(bj * VB), j1 = (float)min(bj * VB + VB - 1, ey - 1);
    const float k0 = (float)(bk * VB), k1 = (float)min(bk * VB + VB - 1, ez - 1);
    const float cx = (0.5f * (i0 + i1) - tx) / sx, cy = (0.5f * (j0 + j1) - ty) / sy, cz = (0.5f * (k0 + k1) - tz) / sz;
    const float R2 = radius2[b];
    // this thread's two cells: local ids threadIdx.x and threadIdx.x + 256 of the 8x8x8 block
    float qx[2], qy[2], qz[2], best[2];
    uint32_t bidx[2];
    bool live[2];
    size_t lin[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int l = threadIdx.x + 256 * c;
        const int i = bi * VB + (l & 7), j = bj * VB + ((l >> 3) & 7), k = bk * VB + (l >> 6);
        live[c] = i < ex && j < ey && k < ez;
        lin[c] = ((size_t)k * ey + j) * ex + i;
        qx[c] = tm_voxel_centre(cm.a[0], cm.b[0], i);  // exact reference centres (the block ball above is a guarded bound)
        qy[c] = tm_voxel_centre(cm.a[1], cm.b[1], j);
        qz[c] = tm_voxel_centre(cm.a[2], cm.b[2], k);
        best[c] = 3.402823466e+38f;
        bidx[c] = 0;
    }
    for (uint32_t base = 0; base < n; base += VF2_TILE) {
        __syncthreads();  // previous tile's candidates are consumed
        uint32_t tile_count = 0;
        // 4 rounds of 256 points: order-preserving compaction into cand[]
        for (int r = 0; r < VF2_TILE / 256; ++r) {
            const uint32_t t = base + r * 256 + threadIdx.x;
            bool in = false;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < n) {
                p = mpos[t];
                const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
                in = !(dx * dx + dy * dy + dz * dz > R2);  // NaN points stay (they never win anyway)
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, in);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            uint32_t off = tile_count;
            for (int w = 0; w < warp; ++w) off += warp_cnt[w];
            uint32_t round_total = 0;
            for (int w = 0; w < 8; ++w) round_total += warp_cnt[w];
            if (in) {
                p.w = __uint_as_float(t);
                cand[off + __popc(bal & ((1u << lane) - 1u))] = p;
            }
            tile_count += round_total;
            __syncthreads();
        }
        // exact scan (reference arithmetic) over the surviving points, ascending index
        for (uint32_t t = 0; t < tile_count; ++t) {
            const float4 p = cand[t];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float dx = p.x - qx[c], dy = p.y - qy[c], dz = p.z - qz[c];
                const float d = (dx * dx + dy * dy) + dz * dz;
                if (d < best[c]) {
                    best[c] = d;
                    bidx[c] = __float_as_uint(p.w);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
        if (live[c]) voxel[lin[c]] = bidx[c];
    (void)n_cand_s;
}
void launch_voxel_fill(cudaStream_t st, const float4* mpos, uint32_t n, int ex, int ey, int ez,
                       float sx, float sy, float sz, float tx, float ty, float tz,
                       uint32_t* voxel, float* block_scratch) {
    size_t total = (size_t)ex * ey * ez;
    if (!total) return;
    const float s3[3] = {sx, sy, sz}, t3[3] = {tx, ty, tz};
    const tm_centre_map cm = tm_voxel_centre_map(s3, t3);  // model.hpp:63
    if (!block_scratch) {  // brute force (small grids, and the cross-check of the pruned kernel)
        ++g_launch_count;
        voxel_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mpos, n, ex, ey, ez, cm, voxel);
        return;
    }
    const int bx = (ex + VB - 1) / VB, by = (ey + VB - 1) / VB, bz = (ez + VB - 1) / VB;
    const uint32_t nb = (uint32_t)bx * by * bz;
    g_launch_count += 2;
    voxel_block_radius_kernel<<<(nb + 7) / 8, 256, 0, st>>>(mpos, n, ex, ey, ez, sx, sy, sz, tx, ty, tz, bx, by, bz,
                                                           block_scratch);
    voxel_fill_pruned_kernel<<<nb, 256, 0, st>>>(mpos, n, ex, ey, ez, sx, sy, sz, tx, ty, tz, cm, bx, by,
                                                 block_scratch, voxel);
}
size_t voxel_fill_scratch_bytes(int ex, int ey, int ez) {
    return (size_t)((ex + VB - 1) / VB) * ((ey + VB - 1) / VB) * ((ez + VB - 1) / VB) * sizeof(float);
}

// Block-occupancy mask for one distance threshold.  A position p that voxel_query maps to cell c has
// voxel coordinates within one cell of c's fill point q_c per axis (truncation toward zero makes cell
// 0 two cells wide), i.e. |p - q_c| <= D = the cell diagonal, so |p - nn(c)| >= |q_c - nn(c)| - D:
// cell c can hold an inlier only if |q_c - nn(c)| <= thres + D.  reach2 = that bound squared (with a
// relative guard).  A block's bit is set when any of its cells can.
__global__ void __launch_bounds__(256)
    occupancy_kernel(const uint32_t* __restrict__ voxel, const float4* __restrict__ mpos, int ex, int ey, int ez,
                     float sx, float sy, float sz, float tx, float ty, float tz, float reach2, int obx, int oby,
                     uint32_t* __restrict__ occ) {
    const size_t total = (size_t)ex * ey * ez;
    const size_t lin = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lin >= total) return;
    const int i = (int)(lin % ex), j = (int)((lin / ex) % ey), k = (int)(lin / ((size_t)ex * ey));
    const float qx = ((float)i - tx) / sx, qy = ((float)j - ty) / sy, qz = ((float)k - tz) / sz;
    const float4 p = mpos[voxel[lin]];
    const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
    const float d2 = dx * dx + dy * dy + dz * dz;
    if (!(d2 > reach2)) {  // NaN keeps the block
        const uint32_t b = (uint32_t)(((k >> OCC_SHIFT) * oby + (j >> OCC_SHIFT)) * obx + (i >> OCC_SHIFT));
        atomicOr(&occ[b >> 5], 1u << (b & 31u));
    }
}
void launch_occupancy(cudaStream_t st, const uint32_t* voxel, const float4* mpos, int ex, int ey, int ez, float sx,
                      float sy, float sz, float tx, float ty, float tz, float reach2, int obx, int oby, uint32_t* occ) {
    const size_t total = (size_t)ex * ey * ez;
    if (!total) return;
    ++g_launch_count;
    occupancy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(voxel, mpos, ex, ey, ez, sx, sy, sz, tx, ty, tz,
                                                                    reach2, obx, oby, occ);
}

// ---- L2 gather micro-benchmark: the roofline the scorer is measured against --------------------
// Every lane reads independent pseudo-random 16-byte cells (one 32-byte sector each, like the fused
// grid gathers) from a working set that fits L2; 8 loads in flight per lane.
__global__ void __launch_bounds__(256)
    l2_gather_kernel(const float4* __restrict__ buf, uint32_t n_cells_mask, uint32_t iters, float* __restrict__ out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (uint32_t it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;
            v[u] = __ldg(&buf[(s >> 7) & n_cells_mask]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;  // never true: keeps the loads alive
}
void launch_l2_gather(cudaStream_t st, const float4* buf, uint32_t n_cells_mask, uint32_t iters, float* out, int grid) {
    ++g_launch_count;
    l2_gather_kernel<<<grid, 256, 0, st>>>(buf, n_cells_mask, iters, out);
}

// fused grid: cell -> (nearest model point xyz, (index << 1) | class flag), so scoring needs one gather
// per test instead of the dependent voxel -> point chain; the ref vector of an inlier's model point
// comes from a compact per-point array (model_ref_kernel) that stays L1-resident, not from a second
// per-cell grid (measured: 44.3 -> 42.8 ms on C2, and half the fused-grid memory)
__global__ void fuse_grid_kernel(const uint32_t* __restrict__ voxel, size_t total,
                                 const float4* __restrict__ mpos, float4* __restrict__ vcell) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t mi = voxel[i];
    float4 p = mpos[mi];
    p.w = __uint_as_float((mi << 1) | (__float_as_uint(p.w) & FLAG_TANGENT));
    vcell[i] = p;
}
void launch_fuse_grid(cudaStream_t st, const uint32_t* voxel, size_t total, const float4* mpos, float4* vcell) {
    if (!total) return;
    ++g_launch_count;
    fuse_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(voxel, total, mpos, vcell);
}
__global__ void model_ref_kernel(const float4* __restrict__ mpos, const float4* __restrict__ mnrm,
                                 const float4* __restrict__ mtgt, uint32_t n, float4* __restrict__ mref) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mref[i] = (__float_as_uint(mpos[i].w) & FLAG_TANGENT) ? mtgt[i] : mnrm[i];  // scene.hpp:469-483
}
void launch_model_ref(cudaStream_t st, const float4* mpos, const float4* mnrm, const float4* mtgt, uint32_t n,
                      float4* mref) {
    if (!n) return;
    ++g_launch_count;
    model_ref_kernel<<<(n + 255) / 256, 256, 0, st>>>(mpos, mnrm, mtgt, n, mref);
}

// ------------------------------------------------------------ traits project
// g2l rows r0..r3 (row 3 unused).  kind: 0 cylinder (impl/cylinder_traits.hpp:102-114),
// 1 plane (impl/plane_traits.hpp:66-72), 2 plane2 (impl/plane2_traits.hpp:86-89),
// 3 identity (impl/identity_traits.hpp:33-36).
__global__ void traits_project_kernel(int kind, float4 r0, float4 r1, float4 r2, float radius,
                                      float threshold, const float* __restrict__ xyz, uint64_t n,
                                      float* __restrict__ uvw, uint8_t* __restrict__ ok) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    float u = x, v = y, w = z;
    bool good = true;
    if (kind != 3) {
        float lx = row_apply(r0, x, y, z), ly = row_apply(r1, x, y, z), lz = row_apply(r2, x, y, z);
        if (kind == 0) {
            float height = sqrtf(lx * lx + ly * ly) - radius;
            if (fabsf(height) > threshold) {
                good = false;
                u = v = w = 0.f;
            } else {
                float ang = atan2f_full(ly, lx);
                if (ang < 0.f) ang = (float)((double)ang + 2.0 * 3.14159265358979323846);
                u = ang * radius;
                v = lz;
                w = height / radius;
            }
        } else {
            u = lx; v = ly; w = lz;
            if (kind == 1 && fabsf(lz) > threshold) {
                good = false;
                u = v = w = 0.f;
            }
        }
    }
    uvw[3 * i] = u;
    uvw[3 * i + 1] = v;
    uvw[3 * i + 2] = w;
    ok[i] = good ? 1 : 0;
}
void launch_traits_project(cudaStream_t st, int kind, float4 r0, float4 r1, float4 r2, float radius,
                           float threshold, const float* xyz, uint64_t n, float* uvw, uint8_t* ok) {
    if (!n) return;
    ++g_launch_count;
    traits_project_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kind, r0, r1, r2, radius,
                                                                      threshold, xyz, n, uvw, ok);
}

// ------------------------------------------------------------------ L2 flush
__global__ void flush_kernel(float4* buf, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(v, v, v, v);
}
void launch_flush(cudaStream_t st, float4* buf, size_t n, float v) {
    ++g_launch_count;
    flush_kernel<<<148 * 8, 256, 0, st>>>(buf, n, v);
}

}  // namespace tmk
