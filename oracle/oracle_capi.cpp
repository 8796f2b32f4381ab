// oracle/oracle_capi.cpp — extern "C" surface of the CPU ORACLE (test
// infrastructure only; see oracle.hpp header).  Loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Transforms cross
// this boundary as column-major float[16] (Eigen mat4f_t layout).
#include "oracle.hpp"

#include <atomic>
#include <map>

using namespace orc;

namespace {

struct model_h {
    std::vector<float> pos, nrm, tgt;
    model m;
};
struct scene_h {
    std::vector<float> pos, nrm, tgt;
    scene s;
};

m4 from_colmajor(const float* t) {
    m4 r;
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) r.m[rr][c] = t[c * 4 + rr];
    return r;
}
void to_colmajor(const m4& r, float* t) {
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) t[c * 4 + rr] = r.m[rr][c];
}

// table in deterministic key order (lexicographic), values in equal_range order
typedef std::map<std::array<uint32_t, 4>, std::vector<std::pair<uint32_t, uint32_t>>> table_t;
table_t export_table(const model& m, uint32_t cap) {
    table_t t;
    for (auto it = m.map.begin(); it != m.map.end();) {
        auto range = m.map.equal_range(it->first);
        std::array<uint32_t, 4> k = {it->first.k[0], it->first.k[1], it->first.k[2], it->first.k[3]};
        auto& v = t[k];
        for (auto e = range.first; e != range.second; ++e) {
            if (cap && v.size() >= cap) break;
            v.push_back(e->second);
        }
        it = range.second;
    }
    return t;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ primitives
uint32_t orc_murmur4(const uint32_t* key) { return murmur4(key); }
uint32_t orc_discretize_range(float v, float mn, float range, uint32_t steps) {
    return discretize(v, mn, range, steps);
}
uint32_t orc_discretize_step(float v, float step) { return discretize(v, step); }
float orc_atan2f_q1(float y, float x) { return atan2f_q1(y, x); }
float orc_atan2f_full(float y, float x) { return atan2f_full(y, x); }
float orc_libm_atan2f(float y, float x) { return ::atan2f(y, x); }
void orc_atan2f_q1_batch(const float* y, const float* x, uint64_t n, float* out, float* out_libm) {
    for (uint64_t i = 0; i < n; ++i) {
        out[i] = atan2f_q1(y[i], x[i]);
        if (out_libm) out_libm[i] = ::atan2f(y[i], x[i]);
    }
}
// in: p0,t0,p1,t1 (12 floats) -> f[4]
// tm_sincosf.h (pcl::eigen33's cos / sin) next to the host libm's binary32 routines
void orc_sincosf_small_batch(const float* x, uint64_t n, float* s, float* c, float* s_libm, float* c_libm) {
    for (uint64_t i = 0; i < n; ++i) {
        s[i] = tm_math::sinf_small(x[i]);
        c[i] = tm_math::cosf_small(x[i]);
        s_libm[i] = ::sinf(x[i]);
        c_libm[i] = ::cosf(x[i]);
    }
}
void orc_feature(const float* in, int use_libm, float* f) {
    feature({in[0], in[1], in[2]}, {in[3], in[4], in[5]}, {in[6], in[7], in[8]},
            {in[9], in[10], in[11]}, f, use_libm != 0);
}
// in: src_i, src_j, src_t, tgt_i, tgt_j, tgt_t (18 floats) -> column-major 4x4
void orc_base_transform(const float* in, float* out16) {
    m4 t = base_transform({in[0], in[1], in[2]}, {in[3], in[4], in[5]}, {in[6], in[7], in[8]},
                          {in[9], in[10], in[11]}, {in[12], in[13], in[14]},
                          {in[15], in[16], in[17]});
    to_colmajor(t, out16);
}
uint32_t orc_early_drop_upper(uint32_t tried, uint32_t nsub, uint32_t corrs) {
    return early_drop_upper(tried, nsub, corrs);
}
void orc_early_drop_tests(uint64_t nsub, uint32_t* out18) {
    auto t = early_drop_tests(nsub);
    for (size_t i = 0; i < t.size(); ++i) out18[i] = t[i];
}
void orc_voxel_centre_map(const float* scale3, const float* trans3, float* ia3, float* ib3) {
    voxel_centre_map(scale3, trans3, ia3, ib3);
}
float orc_resolution(const float* pos, uint32_t n) {
    cloud c{pos, nullptr, nullptr, n};
    return resolution(c);
}
void orc_umeyama(const float* src, const float* dst, uint32_t n, float* out16) {
    std::vector<v3> s(n), d(n);
    for (uint32_t i = 0; i < n; ++i) {
        s[i] = ld3(src, i);
        d[i] = ld3(dst, i);
    }
    to_colmajor(umeyama(s, d), out16);
}
// traits project: kind 0 cylinder, 1 plane, 2 plane2, 3 identity.  g2l column-major.
int orc_traits_project(int kind, const float* g2l16, float radius, float threshold,
                       const float* xyz, float* uvw) {
    m4 g = from_colmajor(g2l16);
    v3 p = {xyz[0], xyz[1], xyz[2]};
    switch (kind) {
        case 0: return cylinder_project(g, radius, threshold, p, uvw) ? 1 : 0;
        case 1: return plane_project(g, threshold, p, uvw) ? 1 : 0;
        case 2: return plane2_project(g, p, uvw) ? 1 : 0;
        default: return identity_project(p, uvw) ? 1 : 0;
    }
}
// k-NN (inclusive), principal curvatures and the tangent criterion of scene.hpp:50 / model.hpp:98
void orc_knn(const float* pos, uint32_t n, const uint32_t* query, uint32_t n_query, uint32_t k, int32_t* out_idx,
             float* out_d2) {
    cloud c{pos, pos, pos, n};
    std::vector<int32_t> idx;
    std::vector<float> d2;
    for (uint32_t w = 0; w < n_query; ++w) {
        knn_inclusive(c, query[w], k, idx, d2);
        for (uint32_t j = 0; j < k; ++j) {
            out_idx[(size_t)w * k + j] = idx[j];
            if (out_d2) out_d2[(size_t)w * k + j] = d2[j];
        }
    }
}
void orc_curvature(const float* pos, const float* nrm, uint32_t n, const uint32_t* query, uint32_t n_query, uint32_t k,
                   const int32_t* nbr, float* pc_min, float* pc_max, float* cov9) {
    cloud c{pos, nrm, nrm, n};
    for (uint32_t w = 0; w < n_query; ++w) {
        const int32_t* my = nbr + (size_t)w * k;
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < k; ++j) cnt += my[j] >= 0;
        float cov[3][3];
        principal_curvatures(c, query[w], my, cnt, cov, pc_min[w], pc_max[w]);
        if (cov9)
            for (int i = 0; i < 9; ++i) cov9[(size_t)w * 9 + i] = cov[i / 3][i % 3];
    }
}
void orc_eigen33(const float* cov9, float* evals) {
    float cov[3][3];
    for (int i = 0; i < 9; ++i) cov[i / 3][i % 3] = cov9[i];
    pcl_eigen33(cov, evals);
}
// opencl/icp.cl restated over all work-items; returns the number of emitted correspondences
uint32_t orc_cl_icp_projection(int projector, const float* pnts4, int n, const float* image4, const int32_t* img_size,
                               const int32_t* img_margin, const float* mat_align, const float* mat_uvw,
                               const float* mat_proj, const float* mat_norm, float max_corr_dist,
                               float* out_positions4, int32_t* model_indices, int32_t* scene_indices) {
    uint32_t c = 0;
    for (int i = 0; i < n; ++i) {
        cl_icp_projection(projector, pnts4, i, image4, img_size, img_margin, mat_align, mat_uvw, mat_proj, mat_norm,
                          max_corr_dist, out_positions4, model_indices, scene_indices);
        c += model_indices[i] >= 0;
    }
    return c;
}
// per-correspondence float16 records + their sum in double (index order)
void orc_cl_icp_correlation(const float* scene4, const float* model4, const int32_t* is, const int32_t* im, int n,
                            const float* cs, const float* cm, float* records16, double* cov9) {
    for (int k = 0; k < 9; ++k) cov9[k] = 0.0;
    for (int i = 0; i < n; ++i) {
        float o[16];
        cl_icp_correlation(scene4, model4, is, im, n, i, cs, cm, o);
        if (records16) std::memcpy(records16 + 16 * (size_t)i, o, 64);
        for (int k = 0; k < 9; ++k) cov9[k] += (double)o[k];
    }
}
uint32_t orc_get_octant(const float* center, const float* pos) {
    return get_octant({center[0], center[1], center[2]}, {pos[0], pos[1], pos[2]});
}

// ----------------------------------------------------------------------- model
void* orc_model_create(const float* pos, const float* nrm, const float* tgt, uint32_t n,
                       const uint8_t* curv_ok, float distance_step_count, float angle_step,
                       float min_diameter_factor, float max_diameter_factor, float resolution_) {
    auto* h = new model_h();
    h->pos.assign(pos, pos + 3 * (size_t)n);
    h->nrm.assign(nrm, nrm + 3 * (size_t)n);
    h->tgt.assign(tgt, tgt + 3 * (size_t)n);
    cloud c{h->pos.data(), h->nrm.data(), h->tgt.data(), n};
    discretization_params dp{distance_step_count, angle_step, 10.f};
    sample_parameters sp{min_diameter_factor, max_diameter_factor, false};
    h->m.init(c, dp, sp, curv_ok, resolution_);
    return h;
}
// the same with the voxel grid supplied (extents product entries, reference linearisation) instead of filled
void* orc_model_create_with_grid(const float* pos, const float* nrm, const float* tgt, uint32_t n,
                                 const uint8_t* curv_ok, float distance_step_count, float angle_step,
                                 float min_diameter_factor, float max_diameter_factor, float resolution_,
                                 const uint32_t* voxel_in) {
    auto* h = new model_h();
    h->pos.assign(pos, pos + 3 * (size_t)n);
    h->nrm.assign(nrm, nrm + 3 * (size_t)n);
    h->tgt.assign(tgt, tgt + 3 * (size_t)n);
    cloud c{h->pos.data(), h->nrm.data(), h->tgt.data(), n};
    discretization_params dp{distance_step_count, angle_step, 10.f};
    sample_parameters sp{min_diameter_factor, max_diameter_factor, false};
    h->m.init(c, dp, sp, curv_ok, resolution_, voxel_in);
    return h;
}
// model::init(subset, params): in_subset[i] != 0 marks the caller's subset
void* orc_model_create_subset(const float* pos, const float* nrm, const float* tgt, uint32_t n,
                              const uint8_t* in_subset, const uint8_t* curv_ok, float distance_step_count,
                              float angle_step, float min_diameter_factor, float max_diameter_factor, float resolution_) {
    auto* h = new model_h();
    h->pos.assign(pos, pos + 3 * (size_t)n);
    h->nrm.assign(nrm, nrm + 3 * (size_t)n);
    h->tgt.assign(tgt, tgt + 3 * (size_t)n);
    cloud c{h->pos.data(), h->nrm.data(), h->tgt.data(), n};
    discretization_params dp{distance_step_count, angle_step, 10.f};
    sample_parameters sp{min_diameter_factor, max_diameter_factor, false};
    h->m.init(c, dp, sp, curv_ok, resolution_, nullptr, in_subset);
    return h;
}
// model.hpp:87-88 for n cells given as (i, j, k) triples: the exact brute-force nearest point
void orc_model_cell_nearest(void* p, const int32_t* ijk, uint64_t n, uint32_t* out) {
    const model& m = static_cast<model_h*>(p)->m;
    for (uint64_t q = 0; q < n; ++q) out[q] = m.cell_nearest(ijk[3 * q], ijk[3 * q + 1], ijk[3 * q + 2]);
}
void orc_model_destroy(void* p) { delete static_cast<model_h*>(p); }
// floats: resolution, diameter, scale[3], trans[3], fb.mn[4], fb.mx[4]  (16)
// ints: extents[3], margin; counts: n_subset, n_entries, n_keys
void orc_model_info(void* p, float* f16, int* i4, uint64_t* c3) {
    const model& m = static_cast<model_h*>(p)->m;
    f16[0] = m.resolution_;
    f16[1] = m.diameter;
    for (int k = 0; k < 3; ++k) {
        f16[2 + k] = m.scale[k];
        f16[5 + k] = m.trans[k];
        i4[k] = m.extents[k];
    }
    i4[3] = m.margin;
    for (int k = 0; k < 4; ++k) {
        f16[8 + k] = m.fb.mn[k];
        f16[12 + k] = m.fb.mx[k];
    }
    c3[0] = m.subset.size();
    c3[1] = m.map.size();
    uint64_t nk = 0;
    for (auto it = m.map.begin(); it != m.map.end(); it = m.map.equal_range(it->first).second) ++nk;
    c3[2] = nk;
}
void orc_model_voxels(void* p, uint32_t* out) {
    const model& m = static_cast<model_h*>(p)->m;
    std::copy(m.voxel.begin(), m.voxel.end(), out);
}
void orc_model_subset(void* p, uint32_t* out) {
    const model& m = static_cast<model_h*>(p)->m;
    std::copy(m.subset.begin(), m.subset.end(), out);
}
// CSR export: keys lexicographically sorted, values in equal_range order,
// capped at `cap` per key (0 = uncapped).  Null outputs => returns total only.
uint64_t orc_model_table(void* p, uint32_t cap, uint32_t* keys, uint32_t* offsets, uint32_t* pairs) {
    const model& m = static_cast<model_h*>(p)->m;
    table_t t = export_table(m, cap);
    uint64_t total = 0, ki = 0;
    for (auto& kv : t) {
        if (keys) {
            for (int a = 0; a < 4; ++a) keys[4 * ki + a] = kv.first[a];
            offsets[ki] = (uint32_t)total;
        }
        for (auto& e : kv.second) {
            if (pairs) {
                pairs[2 * total] = e.first;
                pairs[2 * total + 1] = e.second;
            }
            ++total;
        }
        ++ki;
    }
    if (keys) offsets[ki] = (uint32_t)total;
    return total;
}
// include/impl/model.hpp:169-178 + the caller's query_limit (scene.hpp:310)
uint32_t orc_model_query(void* p, const float* f, uint32_t limit, uint32_t* pairs_out) {
    const model& m = static_cast<model_h*>(p)->m;
    key4 k;
    discretize_feature(f, m.fb, m.dp, k.k);
    auto range = m.map.equal_range(k);
    uint32_t cnt = 0;
    for (auto it = range.first; it != range.second; ++it) {
        if (limit && cnt >= limit) break;
        if (pairs_out) {
            pairs_out[2 * cnt] = it->second.first;
            pairs_out[2 * cnt + 1] = it->second.second;
        }
        ++cnt;
    }
    return cnt;
}
int orc_model_voxel_query(void* p, const float* pos4, uint32_t* out) {
    const model& m = static_cast<model_h*>(p)->m;
    return m.voxel_query(pos4, *out) ? 1 : 0;
}

// ----------------------------------------------------------------------- scene
void* orc_scene_create(const float* pos, const float* nrm, const float* tgt, uint32_t n,
                       const uint8_t* tangent_mask, const uint8_t* mask) {
    auto* h = new scene_h();
    h->pos.assign(pos, pos + 3 * (size_t)n);
    h->nrm.assign(nrm, nrm + 3 * (size_t)n);
    h->tgt.assign(tgt, tgt + 3 * (size_t)n);
    h->s.c = cloud{h->pos.data(), h->nrm.data(), h->tgt.data(), n};
    h->s.mask.assign(n, 0);
    h->s.tangent_mask.assign(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        h->s.tangent_mask[i] = tangent_mask ? tangent_mask[i] : 0;
        h->s.mask[i] = mask ? mask[i] : 0;
    }
    return h;
}
void orc_scene_destroy(void* p) { delete static_cast<scene_h*>(p); }
void orc_scene_set_mask(void* p, const uint8_t* mask) {
    scene& s = static_cast<scene_h*>(p)->s;
    for (uint32_t i = 0; i < s.c.n; ++i) s.mask[i] = mask[i];
}
// ball subset around scene point `idx` (ascending); out may be null (count only)
uint64_t orc_ball_subset(void* sp, uint32_t idx, float radius, int* out) {
    const scene& s = static_cast<scene_h*>(sp)->s;
    auto v = ball_subset(s.c, ld3(s.c.pos, idx), radius);
    if (out) std::copy(v.begin(), v.end(), out);
    return v.size();
}
// pair stage (scene.hpp:290-304): filters + feature + valid + key
void orc_pair_features(void* sp, void* mp, const uint32_t* pi, const uint32_t* pj, uint64_t n,
                       float min_diameter_factor, float max_diameter_factor, float* feats,
                       uint32_t* keys, uint8_t* valid_out) {
    const scene& s = static_cast<scene_h*>(sp)->s;
    const model& m = static_cast<model_h*>(mp)->m;
    float lower = m.diameter * min_diameter_factor, upper = m.diameter * max_diameter_factor;
    lower *= lower;
    upper *= upper;
    for (uint64_t q = 0; q < n; ++q) {
        float f[4] = {0, 0, 0, 0};
        bool ok = scene_pair_feature(s, m, pi[q], pj[q], lower, upper, f);
        uint32_t k[4] = {0, 0, 0, 0};
        if (ok) discretize_feature(f, m.fb, m.dp, k);
        for (int a = 0; a < 4; ++a) {
            if (feats) feats[4 * q + a] = f[a];
            keys[4 * q + a] = k[a];
        }
        valid_out[q] = ok ? 1 : 0;
    }
}
// one project_ call; corrs outputs sized nsub by the caller (may be null)
uint32_t orc_project(void* sp, void* mp, const int* subset, uint64_t nsub, const float* T16,
                     float accept_prob, float dist_thres, int early_out, uint32_t* scene_corrs,
                     uint32_t* model_corrs, double* score, uint32_t* saved, int* dropped) {
    const scene& s = static_cast<scene_h*>(sp)->s;
    const model& m = static_cast<model_h*>(mp)->m;
    project_result r =
        project(s, m, subset, nsub, from_colmajor(T16), accept_prob, dist_thres, early_out != 0);
    if (scene_corrs) std::copy(r.scene_corrs.begin(), r.scene_corrs.end(), scene_corrs);
    if (model_corrs) std::copy(r.model_corrs.begin(), r.model_corrs.end(), model_corrs);
    if (score) *score = r.score;
    if (saved) *saved = r.saved;
    if (dropped) *dropped = r.dropped ? 1 : 0;
    return (uint32_t)r.scene_corrs.size();
}
// batch of project_ calls, `nthreads` std::threads over hypotheses (the
// reference fans out with std::async, scene.hpp:146-166).  hyp_sub[h] selects
// the subset (CSR sub_off/sub_idx); hyp_sub == null => all scene points.
void orc_score_batch(void* sp, void* mp, const float* T16s, uint64_t n_hyp, const uint32_t* hyp_sub,
                     const uint64_t* sub_off, const int* sub_idx, float accept_prob,
                     float dist_thres, int early_out, int nthreads, uint32_t* counts,
                     double* scores, uint8_t* dropped) {
    const scene& s = static_cast<scene_h*>(sp)->s;
    const model& m = static_cast<model_h*>(mp)->m;
    std::vector<int> all;
    if (!hyp_sub) {
        all.resize(s.c.n);
        for (uint32_t i = 0; i < s.c.n; ++i) all[i] = (int)i;
    }
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (;;) {
            uint64_t h0 = next.fetch_add(16);
            if (h0 >= n_hyp) break;
            for (uint64_t h = h0; h < std::min(h0 + 16, n_hyp); ++h) {
                const int* sub = hyp_sub ? sub_idx + sub_off[hyp_sub[h]] : all.data();
                uint64_t ns = hyp_sub ? sub_off[hyp_sub[h] + 1] - sub_off[hyp_sub[h]] : all.size();
                project_result r = project(s, m, sub, ns, from_colmajor(T16s + 16 * h), accept_prob,
                                           dist_thres, early_out != 0);
                counts[h] = (uint32_t)r.scene_corrs.size();
                if (scores) scores[h] = r.score;
                if (dropped) dropped[h] = r.dropped ? 1 : 0;
            }
        }
    };
    if (nthreads <= 1) {
        work();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
}
// hypotheses of the recorded pair list: for each valid pair, up to `limit` hash
// hits in equal_range order, each turned into a transform by base_transform_
// (scene.hpp:304-319).  Outputs sized by a first call with T16s == null.
// hyp_pair[h] = pair index; hyp_valid[h] = 0 when force_up rejects it.
uint64_t orc_hypotheses(void* sp, void* mp, const uint32_t* pi, const uint32_t* pj, uint64_t n,
                        float min_diameter_factor, float max_diameter_factor, uint32_t limit,
                        int force_up, float* T16s, uint32_t* hyp_pair, uint32_t* hyp_mi,
                        uint32_t* hyp_mj, uint8_t* hyp_valid) {
    const scene& s = static_cast<scene_h*>(sp)->s;
    const model& m = static_cast<model_h*>(mp)->m;
    float lower = m.diameter * min_diameter_factor, upper = m.diameter * max_diameter_factor;
    lower *= lower;
    upper *= upper;
    uint64_t nh = 0;
    for (uint64_t q = 0; q < n; ++q) {
        float f[4];
        if (!scene_pair_feature(s, m, pi[q], pj[q], lower, upper, f)) continue;
        key4 k;
        discretize_feature(f, m.fb, m.dp, k.k);
        auto range = m.map.equal_range(k);
        uint32_t query = 0;
        for (auto it = range.first; it != range.second; ++it) {
            if (limit && (++query) > limit) break;
            if (T16s) {
                uint32_t mi = it->second.first, mj = it->second.second;
                m4 t = base_transform(ld3(s.c.pos, pi[q]), ld3(s.c.pos, pj[q]), ld3(s.c.tgt, pi[q]),
                                      ld3(m.c.pos, mi), ld3(m.c.pos, mj), ld3(m.c.tgt, mi));
                to_colmajor(t, T16s + 16 * nh);
                hyp_pair[nh] = (uint32_t)q;
                hyp_mi[nh] = mi;
                hyp_mj[nh] = mj;
                hyp_valid[nh] = (force_up && fabsf(1.f - t.m[2][2]) > 0.01f) ? 0 : 1;
            }
            ++nh;
        }
    }
    return nh;
}
// icp_ (scene.hpp:369-404) of one start transform
uint32_t orc_icp(void* sp, void* mp, const float* T16_in, uint32_t max_iterations, float dist_thres,
                 float accept_prob, float* T16_out, double* score, uint32_t* iters) {
    const scene& s = static_cast<scene_h*>(sp)->s;
    const model& m = static_cast<model_h*>(mp)->m;
    match start{from_colmajor(T16_in), {}, {}, 0.0};
    // max_iterations == 0: icp_ returns the incoming match unchanged (scene.hpp:371); the
    // incoming match is finish_find(t, dist_thres) (scene.hpp:361-364)
    match r = max_iterations == 0 ? finish_find(s, m, start.transform, accept_prob, dist_thres)
                                  : icp(s, m, start, max_iterations, dist_thres, accept_prob, iters);
    if (max_iterations == 0 && iters) *iters = 0;
    to_colmajor(r.transform, T16_out);
    if (score) *score = r.score;
    return (uint32_t)r.scene_corrs.size();
}

}  // extern "C"
