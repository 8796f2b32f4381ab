// k_score2.cu — the count-only bulk scorer (v8) and the lazy score pass.
//
//   score_count_x2_kernel   project_ with early_out = false (include/impl/scene.hpp:411-510 through
//                           finish_find semantics), inlier COUNTS only.  Same work decomposition as
//                           score_full_kernel (k_score.cu): warp-granular (tile, hypothesis chunk)
//                           items, box cull per 32 hypotheses, exact test for the survivors.  What is
//                           new is the arithmetic unit: the reference's no-FMA sequence is evaluated
//                           two points at a time on the packed FP32 pipe (FFMA2) WITHOUT giving up a
//                           single rounding:
//                               a * b  ==  fma(a, b, -0)      (one rounding of the exact product;
//                                                              (+0) + (-0) = +0, (-0) + (-0) = -0)
//                               a + b  ==  fma(a,  1,  b)     (one rounding of the exact sum)
//                               a - b  ==  fma(b, -1,  a)
//                           The constants -0, 1, -1 arrive as kernel arguments, so ptxas cannot fold
//                           fma(a, b, -0) back into a multiply and contract it with the add that
//                           follows (which it does for mul.rn.f32x2 + add.rn.f32x2 even under
//                           -fmad=false).  One FFMA2 issue slot then does the work of two FMUL or two
//                           FADD; the scorer is issue-bound, not FP32-pipe-bound.
//   score_best_kernel       Σ|ref·ref_n| (scene.hpp:461,479-483) of the selected pose only: the
//                           per-hypothesis score is not part of the selection (argmax / top-k use
//                           counts), so the bulk pass does not pay an LDS.128 + LDG.128 + ~30
//                           instructions per inlier for it.
#include "tm_kernels.cuh"
#include "tm_x2.cuh"

namespace tmk {

// bounding box of a set of tile points -> centre / half extents; valid = the set holds a live point
struct TileBox {
    float cx, cy, cz, hx, hy, hz;
    bool valid;
};
__device__ __forceinline__ TileBox warp_box(float mnx, float mny, float mnz, float mxx, float mxy, float mxz) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, d));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
        mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, d));
    }
    TileBox b;
    b.valid = mnx <= mxx;
    b.cx = 0.5f * (mnx + mxx); b.hx = 0.5f * (mxx - mnx);
    b.cy = 0.5f * (mny + mxy); b.hy = 0.5f * (mxy - mny);
    b.cz = 0.5f * (mnz + mxz); b.hz = 0.5f * (mxz - mnz);
    return b;
}

// Can any point of the box, moved by the rows r0..r2, be an inlier?  Interval test against the grid (identical to
// score_full_kernel's: NaN never culls) and, with SPHERE, the distance-field test below.
template <bool FUSED, bool SPHERE>
__device__ __forceinline__ bool box_survives(const ModelDev& m, const TileBox& b, float4 r0, float4 r1, float4 r2,
                                             float thres, float cell_reach) {
    const float cx = b.cx, cy = b.cy, cz = b.cz, hx = b.hx, hy = b.hy, hz = b.hz;
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
    bool survive = !out;
    if (SPHERE && survive) {
        // Sphere cull for grids that are mostly empty space around a surface (the instantiations that also use the
        // occupancy mask).  The nearest-neighbour grid doubles as a distance field: with x_c = T c (c = centre of the box),
        // z the position the grid was filled for in x_c's cell (the one with integer voxel coordinates, model.hpp:81-94)
        // and q* that cell's nearest model point, every model point m has
        // |x_c - m| >= |z - q*| - |x_c - z| >= |x_c - q*| - 2 |x_c - z|, and every point p of the box has
        // |T p - x_c| <= rho.  So if |x_c - q*| - 2 |x_c - z| - rho > thres no point of the box is within the threshold
        // of ANY model point: it holds no inlier and is skipped.  |x_c - z| is the fractional part of x_c's voxel
        // coordinates in model units (the (-1, 1) band of cell 0 included); rho comes from the interval arithmetic above
        // (no rigidity assumed): the transformed box lies in cc +- ee.  NaN anywhere makes the comparison false (no cull).
        const float xc = r0.x * cx + r0.y * cy + r0.z * cz + r0.w, yc = r1.x * cx + r1.y * cy + r1.z * cz + r1.w,
                    zc = r2.x * cx + r2.y * cy + r2.z * cz + r2.w;
        const float ex_ = fabsf(r0.x) * hx + fabsf(r0.y) * hy + fabsf(r0.z) * hz,
                    ey_ = fabsf(r1.x) * hx + fabsf(r1.y) * hy + fabsf(r1.z) * hz,
                    ez_ = fabsf(r2.x) * hx + fabsf(r2.y) * hy + fabsf(r2.z) * hz;
        const float vx = m.sx * xc + m.tx, vy = m.sy * yc + m.ty, vz = m.sz * zc + m.tz;
        if ((vx > -1.f) & (vx < m.exf) & (vy > -1.f) & (vy < m.eyf) & (vz > -1.f) & (vz < m.ezf)) {
            const int ci = (int)vx, cj = (int)vy, ck = (int)vz;
            const uint32_t lc = (uint32_t)((ck * m.ey + cj) * m.ex + ci);
            const float fx = (vx - (float)ci) / m.sx, fy = (vy - (float)cj) / m.sy, fz = (vz - (float)ck) / m.sz;
            const float dcell = fminf(sqrtf(fx * fx + fy * fy + fz * fz), cell_reach);
            const float4 qn = FUSED ? __ldg(&m.vcell[lc]) : __ldg(&m.cloud.pos[__ldg(&m.voxel[lc])]);
            const float dx = xc - qn.x, dy = yc - qn.y, dz = zc - qn.z;
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
            float rho = sqrtf(ex_ * ex_ + ey_ * ey_ + ez_ * ez_);
            {   // a (numerically) orthonormal R maps the box's circumscribed sphere onto a sphere of the same radius:
                // |R v| <= sqrt(1 + 3 delta) |v| with delta = max |R^T R - I| (Gershgorin) — tighter than the box of the
                // rotated box; transforms that are not rigid keep the box bound
                const float g00 = r0.x * r0.x + r1.x * r1.x + r2.x * r2.x, g11 = r0.y * r0.y + r1.y * r1.y + r2.y * r2.y,
                            g22 = r0.z * r0.z + r1.z * r1.z + r2.z * r2.z, g01 = r0.x * r0.y + r1.x * r1.y + r2.x * r2.y,
                            g02 = r0.x * r0.z + r1.x * r1.z + r2.x * r2.z, g12 = r0.y * r0.z + r1.y * r1.z + r2.y * r2.z;
                const float delta = fmaxf(fmaxf(fmaxf(fabsf(g00 - 1.f), fabsf(g11 - 1.f)), fabsf(g22 - 1.f)),
                                          fmaxf(fmaxf(fabsf(g01), fabsf(g02)), fabsf(g12)));
                if (delta < 0.01f) rho = fminf(rho, sqrtf(1.f + 3.f * delta) * sqrtf(hx * hx + hy * hy + hz * hz));
            }
            const float mag = fabsf(xc) + fabsf(yc) + fabsf(zc) + fabsf(qn.x) + fabsf(qn.y) + fabsf(qn.z);
            // slack: rounding of everything above, of the grid's own cell positions included (1e-5 relative of the
            // magnitudes involved, 0.1 % of |x_c - z|), 0.1 % of the threshold
            if (dist - 2.002f * dcell - rho * 1.0001f - 1e-5f * mag > thres * 1.001f) survive = false;
        }
    }
    return survive;
}

// Exact test of one hypothesis against the tile's point pairs A (tile points 0..63) and / or B (64..127): the inlier
// count of this lane's points.  A pair the cull excluded holds no inlier and is not evaluated.
template <bool FUSED, bool OCC, bool DO_A, bool DO_B>
__device__ __forceinline__ uint32_t count_eval(const ModelDev& m, const X2& e, float sq_thres, float4 r0, float4 r1, float4 r2,
                                               p2 pxA, p2 pyA, p2 pzA, p2 pxB, p2 pyB, p2 pzB, uint32_t tflags, bool& any) {
    float x[4], y[4], z[4], vx[4], vy[4], vz[4];
    // ---- pos = t*pos (scene.hpp:444) and to_voxel*pos (model.hpp:182), two points per FFMA2
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
    bool in[4];
    bool any_in = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        in[k] = false;
        lin[k] = 0u;
        if (k < 2 ? DO_A : DO_B) {
            // ijk = trunc(v); the reference's own test on the truncated integers (model.hpp:186-189):
            // cvt.rzi saturates, a NaN converts to 0 and is rejected by the distance test (sq is NaN)
            const int i = (int)vx[k], j = (int)vy[k], kk = (int)vz[k];
            bool ok = ((uint32_t)i < (uint32_t)m.ex) & ((uint32_t)j < (uint32_t)m.ey) & ((uint32_t)kk < (uint32_t)m.ez);
            if (OCC) {  // branch-free: out-of-grid lanes look at block 0
                const uint32_t b = ok ? (uint32_t)(((kk >> OCC_SHIFT) * m.oby + (j >> OCC_SHIFT)) * m.obx + (i >> OCC_SHIFT))
                                      : 0u;
                ok = ok & (((__ldg(&m.occ[b >> 5]) >> (b & 31u)) & 1u) != 0u);
            }
            lin[k] = (uint32_t)((kk * m.ey + j) * m.ex + i);  // only dereferenced when in[k]
            in[k] = ok;
            any_in |= ok;
        }
    }
    any = __any_sync(0xffffffffu, any_in);
    if (!any) return 0u;  // nothing reaches the grid
    // ---- the cell gathers in flight; out-of-grid lanes issue no request (predicated loads)
    float4 mp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((k < 2 ? DO_A : DO_B) && in[k]) {
            if (FUSED) mp[k] = __ldg(&m.vcell[lin[k]]);
            else mp[k] = __ldg(&m.cloud.pos[__ldg(&m.voxel[lin[k]])]);
        }
    }
    // ---- dist > thres (scene.hpp:464-467), class agreement (:469-478)
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
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < 2 ? DO_A : DO_B) {
            const uint32_t pfl = (tflags >> k) & 1u;
            const bool inl = in[k] && (sq[k] <= sq_thres) &&  // NaN: not an inlier
                             (((pfl ^ __float_as_uint(mp[k].w)) & FLAG_TANGENT) == 0u);
            c += inl ? 1u : 0u;
        }
    }
    return c;
}

template <bool FUSED, bool OCC, bool STATS>
__global__ void __launch_bounds__(SCORE_THREADS, SCORE_MIN_BLOCKS)
    score_count_x2_kernel(ScoreArgs a, p2 k_nz, p2 k_one, p2 k_mone) {
    static_assert(SCORE_P == 4, "two point pairs per lane");
    __shared__ float4 s_rows[(SCORE_THREADS / 32) * 32 * 3];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float4* my_rows = s_rows + warp * 32 * 3;
    const uint32_t n_items = *a.n_items;
    const ModelDev& m = a.model;
    X2 e;
    e.nz = k_nz; e.one = k_one; e.mone = k_mone;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const WorkItem w = a.items[item];
        // points k = 0..3 of this lane: pairs A = (0, 1) = tile points 0..63, B = (2, 3) = tile points 64..127
        float px[4], py[4], pz[4];
        uint32_t tflags = 0;  // bit k: tangent_mask_ of point k
        const float nanv = __int_as_float(0x7fc00000);
        float mn[2][3] = {{3.0e38f, 3.0e38f, 3.0e38f}, {3.0e38f, 3.0e38f, 3.0e38f}};
        float mx[2][3] = {{-3.0e38f, -3.0e38f, -3.0e38f}, {-3.0e38f, -3.0e38f, -3.0e38f}};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t q = k * 32 + lane;
            px[k] = py[k] = pz[k] = nanv;  // a NaN point fails every test
            if (q < w.npts) {
                const uint32_t idx = a.sub_idx ? (uint32_t)a.sub_idx[w.sub_begin + q] : (uint32_t)(w.sub_begin + q);
                const float4 v = a.scene.pos[idx];
                const uint32_t fl = __float_as_uint(v.w);
                if (!(fl & FLAG_MASKED)) {  // mask_ (scene.hpp:434)
                    px[k] = v.x; py[k] = v.y; pz[k] = v.z;
                    if (fl & FLAG_TANGENT) tflags |= 1u << k;
                    mn[k >> 1][0] = fminf(mn[k >> 1][0], v.x); mx[k >> 1][0] = fmaxf(mx[k >> 1][0], v.x);
                    mn[k >> 1][1] = fminf(mn[k >> 1][1], v.y); mx[k >> 1][1] = fmaxf(mx[k >> 1][1], v.y);
                    mn[k >> 1][2] = fminf(mn[k >> 1][2], v.z); mx[k >> 1][2] = fmaxf(mx[k >> 1][2], v.z);
                }
            }
        }
        // One box per tile; with the distance-field cull (OCC: surface models in a mostly empty grid) what decides is the
        // radius of the box, so there each 64-point pair gets its own box and is culled and evaluated on its own.
        TileBox bA, bB;
        if (OCC) {
            bA = warp_box(mn[0][0], mn[0][1], mn[0][2], mx[0][0], mx[0][1], mx[0][2]);
            bB = warp_box(mn[1][0], mn[1][1], mn[1][2], mx[1][0], mx[1][1], mx[1][2]);
        } else {
            bA = warp_box(fminf(mn[0][0], mn[1][0]), fminf(mn[0][1], mn[1][1]), fminf(mn[0][2], mn[1][2]),
                          fmaxf(mx[0][0], mx[1][0]), fmaxf(mx[0][1], mx[1][1]), fmaxf(mx[0][2], mx[1][2]));
            bB = bA;
        }
        if (!bA.valid && !bB.valid) continue;  // no live (finite, unmasked) point in this tile
        uint32_t n_live = 0;  // STATS only
        if (STATS) {
#pragma unroll
            for (int k = 0; k < 4; ++k) n_live += (px[k] == px[k]) ? 1u : 0u;
            n_live = __reduce_add_sync(0xffffffffu, n_live);
        }
        // v * 1 + (-0) == v for every v (signed zeros and NaN included): the pairs become results of an FFMA2, which
        // pins them to aligned register pairs for the whole hypothesis loop (ptxas otherwise re-packs per use)
        const p2 pxA = e.add(pack2(px[0], px[1]), e.nz), pyA = e.add(pack2(py[0], py[1]), e.nz),
                 pzA = e.add(pack2(pz[0], pz[1]), e.nz);
        const p2 pxB = e.add(pack2(px[2], px[3]), e.nz), pyB = e.add(pack2(py[2], py[3]), e.nz),
                 pzB = e.add(pack2(pz[2], pz[3]), e.nz);
        for (uint32_t h0 = w.hyp_begin; h0 < w.hyp_end; h0 += 32) {
            const uint32_t h = h0 + lane;
            bool sA = false, sB = false;
            __syncwarp();  // readers of the previous batch's rows are done
            if (h < w.hyp_end) {
                const float4 r0 = __ldg(&a.T[3 * (size_t)h]), r1 = __ldg(&a.T[3 * (size_t)h + 1]),
                             r2 = __ldg(&a.T[3 * (size_t)h + 2]);
                my_rows[lane] = r0;
                my_rows[32 + lane] = r1;
                my_rows[64 + lane] = r2;
                sA = bA.valid && box_survives<FUSED, OCC>(m, bA, r0, r1, r2, a.thres, a.cell_reach);
                if (OCC) sB = bB.valid && box_survives<FUSED, OCC>(m, bB, r0, r1, r2, a.thres, a.cell_reach);
            }
            const uint32_t maskA = __ballot_sync(0xffffffffu, sA);  // also orders the smem stores
            const uint32_t maskB = OCC ? __ballot_sync(0xffffffffu, sB) : maskA;
            uint32_t mask = maskA | maskB;
            if (STATS && lane == 0) {
                atomicAdd(&a.stats[0], (unsigned long long)min(32u, w.hyp_end - h0));
                atomicAdd(&a.stats[1], (unsigned long long)__popc(mask));
                atomicAdd(&a.stats[6], (unsigned long long)(__popc(maskA) + __popc(maskB)));
            }
            uint32_t mycnt = 0;
            while (mask) {
                const int hh = __ffs(mask) - 1;
                mask &= mask - 1u;
                const float4 r0 = my_rows[hh], r1 = my_rows[32 + hh], r2 = my_rows[64 + hh];
                bool any;
                uint32_t c;
                if (!OCC || (((maskA & maskB) >> hh) & 1u))
                    c = count_eval<FUSED, OCC, true, true>(m, e, a.sq_thres, r0, r1, r2, pxA, pyA, pzA, pxB, pyB, pzB, tflags, any);
                else if ((maskA >> hh) & 1u)
                    c = count_eval<FUSED, OCC, true, false>(m, e, a.sq_thres, r0, r1, r2, pxA, pyA, pzA, pxB, pyB, pzB, tflags, any);
                else
                    c = count_eval<FUSED, OCC, false, true>(m, e, a.sq_thres, r0, r1, r2, pxA, pyA, pzA, pxB, pyB, pzB, tflags, any);
                if (!any) continue;
                const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
                if (STATS && tot && lane == 0) {
                    atomicAdd(&a.stats[2], 1ull);
                    if (tot == n_live) atomicAdd(&a.stats[3], 1ull);           // every live point of the tile is an inlier
                    else if (tot * 10u >= n_live * 9u) atomicAdd(&a.stats[4], 1ull);  // >= 90 %
                    atomicAdd(&a.stats[5], (unsigned long long)tot);
                }
                if (lane == hh) mycnt = tot;
            }
            if (mycnt) atomicAdd(&a.counts[h], mycnt);
        }
    }
}

// ---- lazy score of the selected pose ----------------------------------------------------------------
// score = Σ|ref·ref_n| over the inliers of ONE hypothesis (scene.hpp:461,479-483), the hypothesis named by the
// packed best key; same fixed-point sum as the bulk scorer's WITH_SCORE path, so the value is identical.
// acc[0] += score (2^-36 quanta), acc[1] += inliers (cross-check against the key).  The caller zeroes acc.
template <bool FUSED>
__global__ void __launch_bounds__(256)
    score_best_kernel(CloudDev scene, ModelDev m, const int32_t* __restrict__ sub_idx,
                      const unsigned long long* __restrict__ sub_off, const uint32_t* __restrict__ g_hyp,
                      uint32_t n_groups, const float4* __restrict__ T, const unsigned long long* __restrict__ best_key,
                      const unsigned long long* __restrict__ shard, float sq_thres, unsigned long long* acc) {
    const unsigned long long key = *best_key;
    if (!key) return;
    const unsigned long long gid = 0xFFFFFFFFull - (key & 0xFFFFFFFFull);
    if (gid < shard[0] || gid >= shard[1]) return;  // another rank owns the winner
    const uint32_t l = (uint32_t)(gid - shard[0]);
    // subset row: the last g with g_hyp[g] <= l (rows without hypotheses have g_hyp[g] == g_hyp[g+1])
    uint32_t lo = 0, hi = n_groups;  // invariant: g_hyp[lo] <= l < g_hyp[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (g_hyp[mid] <= l) lo = mid; else hi = mid;
    }
    const unsigned long long sb = sub_off[lo];
    const uint32_t nsub = (uint32_t)(sub_off[lo + 1] - sb);
    const float4 r0 = T[3 * (size_t)l], r1 = T[3 * (size_t)l + 1], r2 = T[3 * (size_t)l + 2];
    unsigned long long sc = 0;
    uint32_t cnt = 0;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nsub; q += gridDim.x * blockDim.x) {
        const uint32_t idx = sub_idx ? (uint32_t)sub_idx[sb + q] : (uint32_t)(sb + q);
        const float4 v = scene.pos[idx];
        const uint32_t fl = __float_as_uint(v.w);
        if (fl & FLAG_MASKED) continue;
        const float x = row_apply(r0, v.x, v.y, v.z), y = row_apply(r1, v.x, v.y, v.z), z = row_apply(r2, v.x, v.y, v.z);
        const float vx = m.sx * x + m.tx, vy = m.sy * y + m.ty, vz = m.sz * z + m.tz;
        if (!((vx > -1.f) & (vx < m.exf) & (vy > -1.f) & (vy < m.eyf) & (vz > -1.f) & (vz < m.ezf))) continue;
        const uint32_t lin = (uint32_t)(((int)vz * m.ey + (int)vy) * m.ex + (int)vx);
        const uint32_t mi = __ldg(&m.voxel[lin]);
        const float4 mp = FUSED ? __ldg(&m.vcell[lin]) : __ldg(&m.cloud.pos[mi]);
        const float dx = x - mp.x, dy = y - mp.y, dz = z - mp.z;
        const float sq = sum3(dx * dx, dy * dy, dz * dz);
        if (!(sq <= sq_thres)) continue;
        if (((fl ^ __float_as_uint(mp.w)) & FLAG_TANGENT) != 0u) continue;
        const f3 ref = mk3((fl & FLAG_TANGENT) ? scene.tgt[idx] : scene.nrm[idx]);
        const f3 rn = mk3(__ldg(&m.mref[mi]));
        const f3 rr = {row_rot(r0, ref), row_rot(r1, ref), row_rot(r2, ref)};
        sc += score_fixed(fabsf(dot3(rr, rn)));
        ++cnt;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, d);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&acc[0], sc);
        atomicAdd(&acc[1], (unsigned long long)cnt);
    }
}
void launch_score_best(cudaStream_t st, const CloudDev& scene, const ModelDev& m, const int32_t* sub_idx,
                       const unsigned long long* sub_off, const uint32_t* g_hyp, uint32_t n_groups, const float4* T,
                       const unsigned long long* best_key, const unsigned long long* shard, float sq_thres,
                       unsigned long long* acc, bool fused) {
    if (!n_groups) return;
    ++g_launch_count;
    if (fused) score_best_kernel<true><<<64, 256, 0, st>>>(scene, m, sub_idx, sub_off, g_hyp, n_groups, T, best_key,
                                                           shard, sq_thres, acc);
    else score_best_kernel<false><<<64, 256, 0, st>>>(scene, m, sub_idx, sub_off, g_hyp, n_groups, T, best_key, shard,
                                                      sq_thres, acc);
}

template <bool FUSED, bool OCC>
static void launch_x2_v(cudaStream_t st, const ScoreArgs& a, int grid) {
    const p2 nz = host_pair(-0.0f), one = host_pair(1.0f), mone = host_pair(-1.0f);
    if (a.stats) score_count_x2_kernel<FUSED, OCC, true><<<grid, SCORE_THREADS, 0, st>>>(a, nz, one, mone);
    else score_count_x2_kernel<FUSED, OCC, false><<<grid, SCORE_THREADS, 0, st>>>(a, nz, one, mone);
}
void launch_score_count_x2(cudaStream_t st, const ScoreArgs& a, int grid, bool fused) {
    ++g_launch_count;
    if (fused) {
        if (a.model.occ) launch_x2_v<true, true>(st, a, grid);
        else launch_x2_v<true, false>(st, a, grid);
    } else {
        if (a.model.occ) launch_x2_v<false, true>(st, a, grid);
        else launch_x2_v<false, false>(st, a, grid);
    }
}
int score_count_x2_max_blocks_per_sm(bool fused) {
    int nb = 0, nb2 = 0;
    if (fused) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_count_x2_kernel<true, false, false>, SCORE_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, score_count_x2_kernel<true, true, false>, SCORE_THREADS, 0);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, score_count_x2_kernel<false, false, false>, SCORE_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, score_count_x2_kernel<false, true, false>, SCORE_THREADS, 0);
    }
    nb = nb < nb2 ? nb : nb2;
    return nb > 0 ? nb : 1;
}

}  // namespace tmk
