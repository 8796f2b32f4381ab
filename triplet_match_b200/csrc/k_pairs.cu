// k_pairs.cu — stages 1, 2 and 3a of the search path:
//   pair_features_probe  pair filter + feature + valid + discretize_feature + hash probe
//                        (include/impl/scene.hpp:290-311, impl/feature.hpp:5-88,
//                         src/discretize.cpp:19-30, impl/discretize.hpp:10-45, impl/model.hpp:169-178)
//   hypotheses           base_transform_ + force_up (include/impl/scene.hpp:312-319, 538-567)
#include "tm_kernels.cuh"

namespace tmk {

__device__ __forceinline__ float feature_angle(f3 a, f3 b) {  // impl/feature.hpp:5-8
    float y = norm3(cross3(a, b));
    float x = fabsf(dot3(a, b));
    return atan2f_q1(y, x);
}

// One thread per recorded pair.  outer may be null (then pair_i is the scene
// index itself); with outer, pair_i[k] indexes outer[].
__global__ void __launch_bounds__(128)
    pair_features_probe_kernel(CloudDev scene, ModelDev model, const uint32_t* __restrict__ outer,
                               const uint32_t* __restrict__ pair_i,
                               const uint32_t* __restrict__ pair_j, uint64_t n, float lower,
                               float upper, uint32_t limit, float* __restrict__ feats,
                               uint4* __restrict__ keys, uint8_t* __restrict__ valid,
                               uint32_t* __restrict__ hit_begin, uint32_t* __restrict__ hit_count,
                               unsigned long long* __restrict__ n_valid) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t i = outer ? outer[pair_i[q]] : pair_i[q];
    uint32_t j = pair_j[q];
    bool ok = (i < scene.n) && (j < scene.n);
    float f0 = 0.f, f1 = 0.f, f2 = 0.f;
    uint4 key = make_uint4(0, 0, 0, 0);
    uint32_t hb = 0, hc = 0;
    if (ok) {
        float4 p1 = scene.pos[i], p2 = scene.pos[j];
        uint32_t fj = __float_as_uint(p2.w);
        // scene.hpp:290: tangent_mask_[j] && !mask_[j] && i != j
        ok = (fj & FLAG_TANGENT) && !(fj & FLAG_MASKED) && (i != j);
        if (ok) {
            f3 t1 = mk3(scene.tgt[i]), t2 = mk3(scene.tgt[j]);
            f3 d0 = sub3(mk3(p2), mk3(p1));
            float sqn0 = sqnorm3(d0);
            f3 d0n = normalized3(d0);
            if (sqn0 < lower || sqn0 > upper) ok = false;                  // :296
            if (ok && (1.f - fabsf(dot3(d0n, t1)) < 0.01f)) ok = false;    // :297
            if (ok) {
                f0 = sqrtf(sqn0);  // feature.hpp:27 d0.norm()
                f1 = feature_angle(d0, t1);
                f2 = feature_angle(d0, t2);
                // valid() feature.hpp:48-88
                const float pi = 3.14159274101257324219f;  // static_cast<float>(M_PI)
                if (f0 < model.fb_min0 || f0 > model.fb_max0) ok = false;
                if (!((f1 >= 0.f && f1 <= pi) && (f2 >= 0.f && f2 <= pi))) ok = false;
            }
            if (ok) {
                float diag0 = model.fb_max0 - model.fb_min0;
                uint32_t k0 = discretize_range(f0, model.fb_min0, diag0, model.dist_steps);
                uint32_t k1 = discretize_step(f1, model.angle_step);
                uint32_t k2 = discretize_step(f2, model.angle_step);
                key = make_uint4(k0, k1, k2, k0);  // f[3] == f[0] (feature.hpp:31)
                if (model.slots) {
                    uint32_t h = murmur4(k0, k1, k2, k0) & model.slot_mask;
                    for (;;) {
                        HashSlot s = model.slots[h];
                        if (s.count == 0u) break;
                        if (s.k[0] == k0 && s.k[1] == k1 && s.k[2] == k2 && s.k[3] == k0) {
                            hb = s.begin;
                            hc = (limit && s.count > limit) ? limit : s.count;
                            break;
                        }
                        h = (h + 1u) & model.slot_mask;
                    }
                }
            }
        }
    }
    if (feats) {
        feats[4 * q + 0] = ok ? f0 : 0.f;
        feats[4 * q + 1] = ok ? f1 : 0.f;
        feats[4 * q + 2] = ok ? f2 : 0.f;
        feats[4 * q + 3] = ok ? f0 : 0.f;
    }
    if (keys) keys[q] = ok ? key : make_uint4(0, 0, 0, 0);
    if (valid) valid[q] = ok ? 1 : 0;
    if (hit_begin) hit_begin[q] = hb;
    if (hit_count) hit_count[q] = hc;
    if (n_valid && ok) atomicAdd(n_valid, 1ull);
}

void launch_pair_features_probe(cudaStream_t st, const CloudDev& scene, const ModelDev& model,
                                const uint32_t* outer, const uint32_t* pair_i,
                                const uint32_t* pair_j, uint64_t n, float lower, float upper,
                                uint32_t limit, float* feats, uint4* keys, uint8_t* valid,
                                uint32_t* hit_begin, uint32_t* hit_count,
                                unsigned long long* n_valid) {
    if (!n) return;
    ++g_launch_count;
    pair_features_probe_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
        scene, model, outer, pair_i, pair_j, n, lower, upper, limit, feats, keys, valid, hit_begin,
        hit_count, n_valid);
}

// probe only (tm_probe): keys given by the caller
__global__ void probe_kernel(ModelDev model, const uint4* __restrict__ keys,
                             const uint8_t* __restrict__ valid, uint64_t n, uint32_t limit,
                             uint32_t* __restrict__ hit_begin, uint32_t* __restrict__ hit_count) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t hb = 0, hc = 0;
    if ((!valid || valid[q]) && model.slots) {
        uint4 k = keys[q];
        uint32_t h = murmur4(k.x, k.y, k.z, k.w) & model.slot_mask;
        for (;;) {
            HashSlot s = model.slots[h];
            if (s.count == 0u) break;
            if (s.k[0] == k.x && s.k[1] == k.y && s.k[2] == k.z && s.k[3] == k.w) {
                hb = s.begin;
                hc = (limit && s.count > limit) ? limit : s.count;
                break;
            }
            h = (h + 1u) & model.slot_mask;
        }
    }
    hit_begin[q] = hb;
    hit_count[q] = hc;
}
void launch_probe(cudaStream_t st, const ModelDev& model, const uint4* keys, const uint8_t* valid,
                  uint64_t n, uint32_t limit, uint32_t* hit_begin, uint32_t* hit_count) {
    if (!n) return;
    ++g_launch_count;
    probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(model, keys, valid, n, limit,
                                                             hit_begin, hit_count);
}

// copy the (m_i, m_j) hit lists into the caller-facing CSR (tm_probe)
__global__ void gather_hits_kernel(ModelDev model, const uint32_t* __restrict__ hit_begin,
                                   const unsigned long long* __restrict__ offsets, uint64_t n,
                                   uint2* __restrict__ out) {
    uint64_t q = blockIdx.x;
    if (q >= n) return;
    unsigned long long o = offsets[q], cnt = offsets[q + 1] - o;
    uint32_t b = hit_begin[q];
    for (uint32_t t = threadIdx.x; t < cnt; t += blockDim.x) out[o + t] = model.hits[b + t];
}
void launch_gather_hits(cudaStream_t st, const ModelDev& model, const uint32_t* hit_begin,
                        const unsigned long long* offsets, uint64_t n, uint2* out) {
    if (!n) return;
    ++g_launch_count;
    gather_hits_kernel<<<(unsigned)n, 64, 0, st>>>(model, hit_begin, offsets, n, out);
}

// ---------------------------------------------------------------- hypotheses
// include/impl/scene.hpp:538-567 (scale_invariant == false)
__device__ __forceinline__ Rows base_transform(f3 src_i, f3 src_j, f3 src_t, f3 tgt_i, f3 tgt_j,
                                               f3 tgt_t) {
    f3 u_a = normalized3(sub3(src_j, src_i));
    f3 u_b = normalized3(sub3(tgt_j, tgt_i));
    float da = dot3(src_t, u_a), db = dot3(tgt_t, u_b);
    f3 v_a = {src_t.x - da * u_a.x, src_t.y - da * u_a.y, src_t.z - da * u_a.z};
    f3 v_b = {tgt_t.x - db * u_b.x, tgt_t.y - db * u_b.y, tgt_t.z - db * u_b.z};
    v_a = normalized3(v_a);
    v_b = normalized3(v_b);
    f3 w_a = normalized3(cross3(u_a, v_a));
    f3 w_b = normalized3(cross3(u_b, v_b));
    // base_a columns (u_a, v_a, w_a): A[r][c]
    float A[3][3] = {{u_a.x, v_a.x, w_a.x}, {u_a.y, v_a.y, w_a.y}, {u_a.z, v_a.z, w_a.z}};
    float B[3][3] = {{u_b.x, v_b.x, w_b.x}, {u_b.y, v_b.y, w_b.y}, {u_b.z, v_b.z, w_b.z}};
    // Eigen compute_inverse_size3: cofactor_3x3<i,j>(m) = m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1)
#define TM_COF(i, j)                                                                   \
    (A[((i) + 1) % 3][((j) + 1) % 3] * A[((i) + 2) % 3][((j) + 2) % 3] -               \
     A[((i) + 1) % 3][((j) + 2) % 3] * A[((i) + 2) % 3][((j) + 1) % 3])
    float c0 = TM_COF(0, 0), c1 = TM_COF(1, 0), c2 = TM_COF(2, 0);
    float det = sum3(c0 * A[0][0], c1 * A[1][0], c2 * A[2][0]);
    float invdet = 1.f / det;
    float I[3][3];
    I[0][0] = c0 * invdet;
    I[0][1] = c1 * invdet;
    I[0][2] = c2 * invdet;
    I[1][0] = TM_COF(0, 1) * invdet;
    I[1][1] = TM_COF(1, 1) * invdet;
    I[1][2] = TM_COF(2, 1) * invdet;
    I[2][0] = TM_COF(0, 2) * invdet;
    I[2][1] = TM_COF(1, 2) * invdet;
    I[2][2] = TM_COF(2, 2) * invdet;
#undef TM_COF
    float R[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[r][c] = sum3(B[r][0] * I[0][c], B[r][1] * I[1][c], B[r][2] * I[2][c]);
    float rx = sum3(R[0][0] * src_i.x, R[0][1] * src_i.y, R[0][2] * src_i.z);
    float ry = sum3(R[1][0] * src_i.x, R[1][1] * src_i.y, R[1][2] * src_i.z);
    float rz = sum3(R[2][0] * src_i.x, R[2][1] * src_i.y, R[2][2] * src_i.z);
    Rows t;
    t.r0 = make_float4(R[0][0], R[0][1], R[0][2], tgt_i.x - rx);
    t.r1 = make_float4(R[1][0], R[1][1], R[1][2], tgt_i.y - ry);
    t.r2 = make_float4(R[2][0], R[2][1], R[2][2], tgt_i.z - rz);
    return t;
}

// One CTA per pair, one thread per hash hit (<= query_limit).  Hypotheses
// outside [h_begin, h_end) (other shards) are skipped.  Rejected hypotheses
// (force_up) keep valid = 0 and NaN rows so that they can never score.
__global__ void __launch_bounds__(256)
    hypotheses_kernel(CloudDev scene, ModelDev model, const uint32_t* __restrict__ outer,
                      const uint32_t* __restrict__ pair_i, const uint32_t* __restrict__ pair_j,
                      uint64_t n_pairs, const unsigned long long* __restrict__ hyp_off,
                      const uint32_t* __restrict__ hit_begin, const uint2* __restrict__ hits,
                      int force_up, const unsigned long long* __restrict__ shard,
                      float4* __restrict__ T, uint8_t* __restrict__ hyp_valid,
                      uint32_t* __restrict__ hyp_pair) {
    uint64_t q = blockIdx.x;
    if (q >= n_pairs) return;
    const unsigned long long h_begin = shard[0], h_end = shard[1];
    unsigned long long o = hyp_off[q];
    uint32_t cnt = (uint32_t)(hyp_off[q + 1] - o);
    if (!cnt) return;
    uint32_t i = outer ? outer[pair_i[q]] : pair_i[q];
    uint32_t j = pair_j[q];
    f3 p1 = mk3(scene.pos[i]), p2 = mk3(scene.pos[j]), t1 = mk3(scene.tgt[i]);
    const uint2* hl = hits + (hit_begin ? hit_begin[q] : o);
    for (uint32_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        unsigned long long h = o + t;
        if (h < h_begin || h >= h_end) continue;
        uint2 mm = hl[t];
        f3 pmi = mk3(model.cloud.pos[mm.x]), pmj = mk3(model.cloud.pos[mm.y]);
        f3 tmi = mk3(model.cloud.tgt[mm.x]);
        Rows r = base_transform(p1, p2, t1, pmi, pmj, tmi);
        bool ok = true;
        if (force_up && fabsf(1.f - r.r2.z) > 0.01f) ok = false;  // scene.hpp:317
        size_t l = (size_t)(h - h_begin);
        if (!ok) {
            float nanv = __int_as_float(0x7fc00000);
            r.r0 = r.r1 = r.r2 = make_float4(nanv, nanv, nanv, nanv);
        }
        T[3 * l + 0] = r.r0;
        T[3 * l + 1] = r.r1;
        T[3 * l + 2] = r.r2;
        if (hyp_valid) hyp_valid[l] = ok ? 1 : 0;
        if (hyp_pair) hyp_pair[l] = (uint32_t)q;
    }
}
void launch_hypotheses(cudaStream_t st, const CloudDev& scene, const ModelDev& model,
                       const uint32_t* outer, const uint32_t* pair_i, const uint32_t* pair_j,
                       uint64_t n_pairs, const unsigned long long* hyp_off,
                       const uint32_t* hit_begin, const uint2* hits, int force_up,
                       const unsigned long long* shard, float4* T, uint8_t* hyp_valid,
                       uint32_t* hyp_pair) {
    if (!n_pairs) return;
    ++g_launch_count;
    hypotheses_kernel<<<(unsigned)n_pairs, 64, 0, st>>>(scene, model, outer, pair_i, pair_j, n_pairs,
                                                       hyp_off, hit_begin, hits, force_up, shard, T,
                                                       hyp_valid, hyp_pair);
}

// column-major float[16] <-> 3 row float4
__global__ void rows_from_colmajor_kernel(const float* __restrict__ T16, uint64_t n,
                                          float4* __restrict__ rows) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    const float* t = T16 + 16 * h;
#pragma unroll
    for (int r = 0; r < 3; ++r) rows[3 * h + r] = make_float4(t[r], t[4 + r], t[8 + r], t[12 + r]);
}
__global__ void colmajor_from_rows_kernel(const float4* __restrict__ rows, uint64_t n,
                                          float* __restrict__ T16) {
    uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    float* t = T16 + 16 * h;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float4 v = rows[3 * h + r];
        t[r] = v.x;
        t[4 + r] = v.y;
        t[8 + r] = v.z;
        t[12 + r] = v.w;
    }
    t[3] = 0.f;
    t[7] = 0.f;
    t[11] = 0.f;
    t[15] = 1.f;
}
void launch_rows_from_colmajor(cudaStream_t st, const float* T16, uint64_t n, float4* rows) {
    if (!n) return;
    ++g_launch_count;
    rows_from_colmajor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T16, n, rows);
}
void launch_colmajor_from_rows(cudaStream_t st, const float4* rows, uint64_t n, float* T16) {
    if (!n) return;
    ++g_launch_count;
    colmajor_from_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows, n, T16);
}

}  // namespace tmk
