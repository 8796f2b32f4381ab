// oracle/ref_shim_capi.cpp — TEST INFRASTRUCTURE.  Compiles the REFERENCE's own sources where
// they lie under /root/reference (nothing is copied into this repository):
//   src/discretize.cpp + include/impl/discretize.hpp   discretize x2, murmur, std::hash        (a3, a4)
//   include/impl/feature.hpp                           angle, feature, valid, discretize_feature (a1-a3)
//   include/impl/model.hpp                             model::init, query, voxel_query          (a5, a9)
//   include/impl/scene.hpp                             base_transform_, project_, finish_find, icp_ (a6, a10-a12)
//   include/impl/pointcloud.hpp                        resolution(), curvature()
// against the header stand-ins in oracle/shim/ (Eigen, PCL/FLANN, boost, fmt and range-v3 are
// absent from this image and there is no network).  What the stand-ins decide — and the
// reference therefore does NOT pin — is listed in oracle/oracle.hpp ("parity unpinned"):
// Eigen's evaluation orders, the 4x4 inverse, umeyama/SVD, kd-tree order and ties, sampling.
// Everything else (control flow, thresholds, casts, filters, container order, early-drop
// arithmetic incl. its uint32 casts as gcc/x86-64 compiles them) is the reference's own code.
// Output: oracle/_ref/libtm_ref.so (git-ignored), used only by tests/test_oracle_vs_ref.py.
#include <atomic>
#include <optional>
#include <thread>
#include <variant>

#include "oracle.hpp"             // only for the rigid solve behind the Eigen::umeyama stand-in

#include <common>                 // /root/reference/include/common
#include <src/discretize.cpp>     // /root/reference/src/discretize.cpp (+ discretize, impl/discretize.hpp)
#include <scene>                  // /root/reference/include/scene (-> model, feature, pointcloud, octree)
#include <impl/pointcloud.hpp>
#include <impl/feature.hpp>
#include <impl/model.hpp>
#include <impl/scene.hpp>
#include <numeric>
#include <cylinder_traits>        // row a14: the traits' closed forms (project / unproject / tangent / normal /
#include <plane_traits>           // intrinsic_distance); their init_* fits (PCL MSAC, SVD) are out of scope
#include <plane2_traits>
#include <identity_traits>
#include <impl/cylinder_traits.hpp>
#include <impl/plane_traits.hpp>
#include <impl/plane2_traits.hpp>
#include <impl/identity_traits.hpp>
#include <octree>                 // row a16: octree build + the five traversals (octree.ipp)
#include <impl/octree.hpp>

namespace Eigen {
Matrix4f umeyama(const Matrix<float, 3, Dynamic>& src, const Matrix<float, 3, Dynamic>& dst, bool) {
    std::vector<orc::v3> s(src.cols()), d(dst.cols());
    for (int i = 0; i < src.cols(); ++i) {
        s[i] = {src(0, i), src(1, i), src(2, i)};
        d[i] = {dst(0, i), dst(1, i), dst(2, i)};
    }
    orc::m4 t = orc::umeyama(s, d);
    Matrix4f r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r(i, j) = t.m[i][j];
    return r;
}
}  // namespace Eigen

namespace tr = triplet_match;
typedef pcl::PointSurfel point_t;
typedef tr::pointcloud<point_t> cloud_t;

// scene<Point>::impl is a protected nested type: reach it through a derived class
struct scene_access : tr::scene<point_t> {
    typedef tr::scene<point_t>::impl impl_t;
};

static point_t make_point(const float* p, const float* t) {
    point_t q;
    q.x = p[0]; q.y = p[1]; q.z = p[2];
    q.data_c[1] = t[0]; q.data_c[2] = t[1]; q.data_c[3] = t[2];
    return q;
}
static cloud_t::Ptr make_cloud(const float* pos, const float* nrm, const float* tgt, uint32_t n) {
    cloud_t::Ptr c = cloud_t::empty();
    for (uint32_t i = 0; i < n; ++i) {
        point_t q = make_point(pos + 3 * i, tgt + 3 * i);
        q.normal_x = nrm[3 * i]; q.normal_y = nrm[3 * i + 1]; q.normal_z = nrm[3 * i + 2];
        c->push_back(q);
    }
    return c;
}
static tr::feature_bounds_t make_bounds(const float* mn, const float* mx) {
    tr::feature_bounds_t b;
    for (int i = 0; i < 4; ++i) { b.min()[i] = mn[i]; b.max()[i] = mx[i]; }
    return b;
}
static tr::mat4f_t mat_from(const float* t16) {
    tr::mat4f_t m;
    for (int i = 0; i < 16; ++i) m.data()[i] = t16[i];  // both column-major
    return m;
}

struct ref_model {
    cloud_t::Ptr cloud;
    std::unique_ptr<tr::model<point_t>> m;
};
struct ref_scene {
    cloud_t::Ptr cloud;
    std::unique_ptr<scene_access::impl_t> impl;
};

// ---- traits (a14): kind 0 cylinder, 1 plane, 2 plane2, 3 identity; g2l column-major -------------
template <typename Tr>
static std::shared_ptr<typename Tr::state_t> make_state(const float* g2l16, const float* l2g16, float radius, float threshold) {
    auto h = std::make_shared<typename Tr::state_t>();
    if constexpr (!std::is_same<Tr, tr::identity_traits<point_t>>::value) {
        h->g2l = mat_from(g2l16);
        h->l2g = mat_from(l2g16);  // given by the caller: Matrix4f::inverse() is a stand-in in the shim
        h->threshold = threshold;
        if constexpr (std::is_same<Tr, tr::cylinder_traits<point_t>>::value) h->radius = radius;
    }
    return h;
}
template <typename Tr>
static int traits_ops(const float* g2l16, const float* l2g16, float radius, float threshold, const float* xyz, const float* pnt_n,
                      const float* pnt_t, float* out15) {
    typename Tr::const_handle_t h = make_state<Tr>(g2l16, l2g16, radius, threshold);
    tr::vec3f_t p(xyz[0], xyz[1], xyz[2]);
    auto uvw = Tr::project(h, p);
    int ok = uvw ? 1 : 0;
    tr::vec3f_t u = uvw ? *uvw : tr::vec3f_t(0.f, 0.f, 0.f);
    tr::vec3f_t back = Tr::unproject(h, u);
    point_t q = make_point(xyz, pnt_t);
    q.normal_x = pnt_n[0]; q.normal_y = pnt_n[1]; q.normal_z = pnt_n[2];
    tr::vec3f_t tg = Tr::tangent(h, q), nm = Tr::normal(h, q);
    float dist = Tr::intrinsic_distance(h, u, tr::vec3f_t(pnt_t[0], pnt_t[1], pnt_t[2]));
    for (int i = 0; i < 3; ++i) { out15[i] = u[i]; out15[3 + i] = back[i]; out15[6 + i] = tg[i]; out15[9 + i] = nm[i]; }
    out15[12] = dist;
    return ok;
}

// ---- octree (a16): node rows {depth, is_leaf, n_points, bbox min3, bbox max3, sum idx, first idx, last idx} ----
static void octree_row(const tr::node& nd, double* row) {
    const tr::base_node* b = tr::as_base_node(nd);
    row[0] = b->depth;
    const tr::leaf_node* lf = std::get_if<tr::leaf_node>(&nd);
    row[1] = lf ? 1.0 : 0.0;
    row[2] = lf ? (double)lf->points.size() : 0.0;
    for (int i = 0; i < 3; ++i) { row[3 + i] = b->bbox.min()[i]; row[6 + i] = b->bbox.max()[i]; }
    double sum = 0.0;
    if (lf) for (uint32_t i : lf->points) sum += (double)i;
    row[9] = sum;
    row[10] = lf && !lf->points.empty() ? (double)lf->points.front() : -1.0;
    row[11] = lf && !lf->points.empty() ? (double)lf->points.back() : -1.0;
}
template <typename Trav>
static uint32_t octree_walk(Trav t, double* rows, uint32_t cap) {
    uint32_t n = 0;
    while (!t.equal(ranges::default_sentinel{})) {
        if (n < cap) octree_row(t.read(), rows + 12 * (size_t)n);
        ++n;
        t.next();
    }
    return n;
}

// ---- traits init_from_samples (cylinder impl:55-98, plane impl:46-62, plane2 impl:50-82) ----
// samples: 3 x {pos3, normal3}; out21: ok | g2l (16, column-major) | radius | origin3
template <typename Tr>
static void traits_init_one(const float* samples, float threshold, float* out21) {
    auto mh = std::make_shared<typename Tr::state_t>();
    mh->threshold = threshold;
    typename Tr::const_handle_t cmh = mh;
    point_t pt[3];
    const float z[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < 3; ++i) {
        pt[i] = make_point(samples + 6 * i, z);
        pt[i].normal_x = samples[6 * i + 3]; pt[i].normal_y = samples[6 * i + 4]; pt[i].normal_z = samples[6 * i + 5];
    }
    typename Tr::handle_t h;
    if constexpr (Tr::sample_count == 1) h = Tr::init_from_samples(cmh, typename Tr::samples_t(pt[0]));
    else if constexpr (Tr::sample_count == 2) h = Tr::init_from_samples(cmh, typename Tr::samples_t(pt[0], pt[1]));
    else h = Tr::init_from_samples(cmh, typename Tr::samples_t(pt[0], pt[1], pt[2]));
    for (int i = 0; i < 21; ++i) out21[i] = 0.f;
    if (!h) return;
    out21[0] = 1.f;
    for (int i = 0; i < 16; ++i) out21[1 + i] = h->g2l.data()[i];
    if constexpr (std::is_same<Tr, tr::cylinder_traits<point_t>>::value) out21[17] = h->radius;
    for (int i = 0; i < 3; ++i) out21[18 + i] = h->origin[i];
}

extern "C" {

uint32_t ref_murmur4(const uint32_t* k) {
    tr::discrete_feature_t key;
    for (int i = 0; i < 4; ++i) key[i] = k[i];
    return tr::detail::murmur<4>(key);
}
uint64_t ref_std_hash4(const uint32_t* k) {
    tr::discrete_feature_t key;
    for (int i = 0; i < 4; ++i) key[i] = k[i];
    return std::hash<tr::discrete_feature_t>()(key);
}
uint32_t ref_discretize_range(float v, float mn, float range, uint32_t steps) { return tr::discretize(v, mn, range, steps); }
uint32_t ref_discretize_step(float v, float step) { return tr::discretize(v, step); }
// in: p0, t0, p1, t1 (12 floats)
void ref_feature(const float* in, float* f) {
    point_t a = make_point(in, in + 3), b = make_point(in + 6, in + 9);
    tr::curv_info_t<point_t> c{};
    auto r = tr::feature<point_t>(a, b, c, c);
    for (int i = 0; i < 4; ++i) f[i] = (*r)[i];
}
int ref_valid(const float* f, const float* mn, const float* mx) {
    tr::feature_t ff;
    for (int i = 0; i < 4; ++i) ff[i] = f[i];
    return tr::valid<point_t>(ff, make_bounds(mn, mx)) ? 1 : 0;
}
void ref_discretize_feature(const float* f, const float* mn, const float* mx, float dist_steps, float angle_step, uint32_t* key) {
    tr::feature_t ff;
    for (int i = 0; i < 4; ++i) ff[i] = f[i];
    tr::discretization_params dp{dist_steps, angle_step, 10.f};
    tr::discrete_feature_t df = tr::discretize_feature<point_t>(ff, make_bounds(mn, mx), dp);
    for (int i = 0; i < 4; ++i) key[i] = df[i];
}
void ref_valid_bounds(const float* mn, const float* mx, float min_rel, float max_rel, float* omn, float* omx) {
    tr::feature_bounds_t b = tr::valid_bounds(make_bounds(mn, mx), 0.f, 0.f, min_rel, max_rel);
    for (int i = 0; i < 4; ++i) { omn[i] = b.min()[i]; omx[i] = b.max()[i]; }
}

// ---- model: the reference's model<PointSurfel>::init / query / voxel_query ----------------
void* ref_model_create(const float* pos, const float* nrm, const float* tgt, uint32_t n,
                       float dist_steps, float angle_step, float min_df, float max_df) {
    auto* h = new ref_model();
    h->cloud = make_cloud(pos, nrm, tgt, n);
    tr::discretization_params dp{dist_steps, angle_step, 10.f};
    h->m.reset(new tr::model<point_t>(h->cloud, dp));
    tr::sample_parameters sp{0.f, 0.f, 1.f, 1.f, min_df, max_df, 0.f, 1.f, false};
    h->m->init(sp);
    return h;
}
// model::init(subset, params) (model.hpp:16-22)
void* ref_model_create_subset(const float* pos, const float* nrm, const float* tgt, uint32_t n, const uint32_t* subset,
                              uint32_t n_subset, float dist_steps, float angle_step, float min_df, float max_df) {
    auto* h = new ref_model();
    h->cloud = make_cloud(pos, nrm, tgt, n);
    tr::discretization_params dp{dist_steps, angle_step, 10.f};
    h->m.reset(new tr::model<point_t>(h->cloud, dp));
    tr::sample_parameters sp{0.f, 0.f, 1.f, 1.f, min_df, max_df, 0.f, 1.f, false};
    tr::subset_t sub(subset, subset + n_subset);
    h->m->init(sub, sp);
    return h;
}
void ref_model_destroy(void* p) { delete static_cast<ref_model*>(p); }
// f: resolution, diameter, feat_min[4], feat_max[4] (10); to_voxel16 column-major; ints: extents[3], margin, point_count
void ref_model_info(void* p, float* f10, float* to_voxel16, int* i5) {
    auto* h = static_cast<ref_model*>(p);
    f10[0] = h->cloud->resolution();
    f10[1] = h->m->diameter();
    for (int k = 0; k < 4; ++k) {
        f10[2 + k] = h->m->feature_bounds().min()[k];
        f10[6 + k] = h->m->feature_bounds().max()[k];
    }
    for (int k = 0; k < 16; ++k) to_voxel16[k] = h->m->voxel_transform().data()[k];
    for (int k = 0; k < 3; ++k) i5[k] = h->m->extents()[k];
    i5[3] = h->m->margin();
    i5[4] = (int)h->m->point_count();
}
// Matrix4f::inverse() as the reference build evaluates it (model.hpp:63), column-major in and out
void ref_matrix4f_inverse(const float* m16, float* out16) {
    tr::mat4f_t m;
    for (int k = 0; k < 16; ++k) m.data()[k] = m16[k];
    tr::mat4f_t inv = m.inverse();
    for (int k = 0; k < 16; ++k) out16[k] = inv.data()[k];
}
// the whole voxel grid of the reference-built model (voxel_data_ is private: every cell is read back through
// the public voxel_query at a position a quarter cell inside it).  out: extents product entries, reference
// linearisation k*ex*ey + j*ex + i.  Returns the number of cells whose query did not land in the cell (0).
uint64_t ref_model_voxels(void* p, uint32_t* out) {
    auto* h = static_cast<ref_model*>(p);
    const auto ext = h->m->extents();
    const float* tv = h->m->voxel_transform().data();
    uint64_t bad = 0;
    for (int k = 0; k < ext[2]; ++k)
        for (int j = 0; j < ext[1]; ++j)
            for (int i = 0; i < ext[0]; ++i) {
                const float x = ((float)i + 0.25f - tv[12]) / tv[0], y = ((float)j + 0.25f - tv[13]) / tv[5],
                            z = ((float)k + 0.25f - tv[14]) / tv[10];
                const tr::vec4f_t pos(x, y, z, 1.f);
                const tr::vec4f_t v = h->m->voxel_transform() * pos;
                if ((int)v[0] != i || (int)v[1] != j || (int)v[2] != k) ++bad;
                auto r = h->m->voxel_query(pos);
                const size_t lin = ((size_t)k * ext[1] + j) * ext[0] + i;
                out[lin] = r ? (*r)[0] : 0xffffffffu;
            }
    return bad;
}
int ref_model_voxel_query(void* p, const float* pos4, uint32_t* out) {
    auto* h = static_cast<ref_model*>(p);
    auto r = h->m->voxel_query(tr::vec4f_t(pos4[0], pos4[1], pos4[2], pos4[3]));
    if (!r) return 0;
    *out = (*r)[0];
    return 1;
}
// model::query + the caller's query_limit loop (scene.hpp:304-311)
uint32_t ref_model_query(void* p, const float* f, uint32_t limit, uint32_t* pairs_out) {
    auto* h = static_cast<ref_model*>(p);
    tr::feature_t ff;
    for (int i = 0; i < 4; ++i) ff[i] = f[i];
    auto range = h->m->query(ff);
    uint32_t query = 0, cnt = 0;
    for (auto it = range.first; it != range.second; ++it) {
        if (limit > 0 && (++query) > limit) break;
        auto&& [m_i, m_j] = it->second;
        pairs_out[2 * cnt] = m_i;
        pairs_out[2 * cnt + 1] = m_j;
        ++cnt;
    }
    return cnt;
}

// ---- scene: the reference's scene<PointSurfel>::impl ---------------------------------------
void* ref_scene_create(const float* pos, const float* nrm, const float* tgt, uint32_t n,
                       const uint8_t* tangent_mask, const uint8_t* mask) {
    auto* h = new ref_scene();
    h->cloud = make_cloud(pos, nrm, tgt, n);
    h->impl.reset(new scene_access::impl_t(h->cloud));
    h->impl->mask_.assign(n, 0);
    h->impl->fp_mask_.assign(n, 0);
    h->impl->tangent_mask_.assign(n, 0);
    for (uint32_t i = 0; i < n; ++i) {
        h->impl->tangent_mask_[i] = tangent_mask[i];
        h->impl->mask_[i] = mask ? mask[i] : 0;
    }
    return h;
}
void ref_scene_destroy(void* p) { delete static_cast<ref_scene*>(p); }
// in: src_i, src_j, src_t, tgt_i, tgt_j, tgt_t (18 floats) -> column-major 4x4
void ref_base_transform(void* sp, const float* in, float* out16) {
    auto* h = static_cast<ref_scene*>(sp);
    auto v = [&](int k) { return tr::vec3f_t(in[3 * k], in[3 * k + 1], in[3 * k + 2]); };
    tr::mat4f_t t = h->impl->base_transform_(v(0), v(1), v(2), v(3), v(4), v(5));
    for (int i = 0; i < 16; ++i) out16[i] = t.data()[i];
}
uint32_t ref_project(void* sp, void* mp, const int* subset, uint64_t nsub, const float* T16,
                     float accept_prob, float dist_thres, int early_out, uint32_t* scene_corrs,
                     uint32_t* model_corrs, double* score, uint32_t* saved) {
    auto* h = static_cast<ref_scene*>(sp);
    auto* m = static_cast<ref_model*>(mp);
    std::vector<int> sub(subset, subset + nsub);
    uint32_t sv = 0;
    auto r = h->impl->project_(sv, *m->m, sub, mat_from(T16), accept_prob, dist_thres, early_out != 0);
    const auto& sc = std::get<0>(r);
    const auto& mc = std::get<1>(r);
    for (size_t i = 0; i < sc.size(); ++i) { scene_corrs[i] = sc[i]; model_corrs[i] = mc[i]; }
    *score = std::get<2>(r);
    *saved = sv;
    return (uint32_t)sc.size();
}
uint32_t ref_icp(void* sp, void* mp, const float* T16_in, uint32_t max_iterations, float dist_thres,
                 float accept_prob, float* T16_out, double* score) {
    auto* h = static_cast<ref_scene*>(sp);
    auto* m = static_cast<ref_model*>(mp);
    // the incoming match of icp_ is finish_find(t, dist_thres) (scene.hpp:361-364)
    auto start = h->impl->finish_find(*m->m, mat_from(T16_in), accept_prob, dist_thres);
    auto r = h->impl->icp_(*m->m, start, max_iterations, dist_thres, accept_prob);
    for (int i = 0; i < 16; ++i) T16_out[i] = r.transform.data()[i];
    *score = r.signed_score;
    return (uint32_t)r.scene_corrs.size();
}
// bench.py --impl reference: the reference's project_ (scene.hpp:411-510) for a batch of
// hypotheses over their recorded radius subsets, fanned out over std::threads the way
// find_parallel fans find_in_subset out over std::async tasks (scene.hpp:146-166)
void ref_project_batch(void* sp, void* mp, const float* T16s, uint64_t n_hyp, const uint32_t* hyp_sub,
                       const uint64_t* sub_off, const int* sub_idx, uint32_t n_sub, float accept_prob,
                       float dist_thres, int early_out, int nthreads, uint32_t* counts, double* scores) {
    auto* h = static_cast<ref_scene*>(sp);
    auto* m = static_cast<ref_model*>(mp);
    std::vector<std::vector<int>> subs(n_sub);
    for (uint32_t g = 0; g < n_sub; ++g) subs[g].assign(sub_idx + sub_off[g], sub_idx + sub_off[g + 1]);
    (void)m->cloud->resolution();  // computed once, as in a real run (cached behind its mutex afterwards)
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (;;) {
            uint64_t b = next.fetch_add(16);
            if (b >= n_hyp) break;
            for (uint64_t q = b; q < std::min<uint64_t>(b + 16, n_hyp); ++q) {
                uint32_t sv = 0;
                auto r = h->impl->project_(sv, *m->m, subs[hyp_sub[q]], mat_from(T16s + 16 * q), accept_prob,
                                           dist_thres, early_out != 0);
                counts[q] = (uint32_t)std::get<0>(r).size();
                scores[q] = std::get<2>(r);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, nthreads); ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}
// the per-pair part of find_in_subset's loop body for pairs that passed the caller's filters:
// feature -> valid -> query (<= limit hits) -> base_transform_ (scene.hpp:299-315).  Returns the
// number of hypotheses written (column-major 4x4 each); hyp_pair[k] = index of the producing pair.
uint64_t ref_hypotheses_batch(void* sp, void* mp, const uint32_t* pair_i, const uint32_t* pair_j, uint64_t n_pairs,
                              uint32_t limit, uint64_t cap, float* T16s, uint32_t* hyp_pair) {
    auto* h = static_cast<ref_scene*>(sp);
    auto* m = static_cast<ref_model*>(mp);
    uint64_t n = 0;
    tr::curv_info_t<point_t> c{};
    for (uint64_t k = 0; k < n_pairs; ++k) {
        const point_t& p1 = h->cloud->points[pair_i[k]];
        const point_t& p2 = h->cloud->points[pair_j[k]];
        auto f = tr::feature<point_t>(p1, p2, c, c);
        if (!f || !tr::valid<point_t>(*f, m->m->feature_bounds())) continue;
        auto range = m->m->query(*f);
        uint32_t query = 0;
        for (auto it = range.first; it != range.second; ++it) {
            if (limit > 0 && (++query) > limit) break;
            auto&& [m_i, m_j] = it->second;
            if (n >= cap) return n;
            const point_t& q1 = m->cloud->points[m_i];
            const point_t& q2 = m->cloud->points[m_j];
            tr::mat4f_t t = h->impl->base_transform_(p1.getVector3fMap(), p2.getVector3fMap(), tr::tangent(p1),
                                                     q1.getVector3fMap(), q2.getVector3fMap(), tr::tangent(q1));
            for (int e = 0; e < 16; ++e) T16s[16 * n + e] = t.data()[e];
            hyp_pair[n] = (uint32_t)k;
            ++n;
        }
    }
    return n;
}
// out15: uvw (3) | unproject(uvw) (3) | tangent(pnt) (3) | normal(pnt) (3) | intrinsic_distance(uvw, pnt_t) (1)
int ref_traits(int kind, const float* g2l16, const float* l2g16, float radius, float threshold, const float* xyz, const float* pnt_n,
               const float* pnt_t, float* out15) {
    switch (kind) {
        case 0: return traits_ops<tr::cylinder_traits<point_t>>(g2l16, l2g16, radius, threshold, xyz, pnt_n, pnt_t, out15);
        case 1: return traits_ops<tr::plane_traits<point_t>>(g2l16, l2g16, radius, threshold, xyz, pnt_n, pnt_t, out15);
        case 2: return traits_ops<tr::plane2_traits<point_t>>(g2l16, l2g16, radius, threshold, xyz, pnt_n, pnt_t, out15);
        default: return traits_ops<tr::identity_traits<point_t>>(g2l16, l2g16, radius, threshold, xyz, pnt_n, pnt_t, out15);
    }
}
void ref_traits_init(int kind, const float* samples, uint32_t n_cases, float threshold, float* out21) {
    for (uint32_t c = 0; c < n_cases; ++c) {
        const float* sm = samples + 18 * (size_t)c;
        float* o = out21 + 21 * (size_t)c;
        if (kind == 0) traits_init_one<tr::cylinder_traits<point_t>>(sm, threshold, o);
        else if (kind == 1) traits_init_one<tr::plane_traits<point_t>>(sm, threshold, o);
        else traits_init_one<tr::plane2_traits<point_t>>(sm, threshold, o);
    }
}
// crit_kind 0 min_voxel_size, 1 max_voxel_size, 2 max_point_count; traversal 0 depth, 1 breadth, 2 leaf,
// 3 branch, 4 level(level); returns the number of visited nodes (rows beyond cap are counted, not written)
uint32_t ref_octree(const float* pos, uint32_t n, const uint32_t* subset, uint32_t n_subset, uint32_t max_depth,
                    int crit_kind, float crit_value, int traversal, uint32_t level, double* rows, uint32_t cap,
                    uint32_t* depth_out) {
    std::vector<float> z(3 * (size_t)n, 0.f);
    cloud_t::Ptr c = make_cloud(pos, z.data(), z.data(), n);
    tr::subdivision_criterion_t crit;
    if (crit_kind == 0) crit = tr::min_voxel_size{crit_value};
    else if (crit_kind == 1) crit = tr::max_voxel_size{crit_value};
    else crit = tr::max_point_count{(uint32_t)crit_value};
    std::optional<tr::subset_t> sub;
    if (subset) sub = tr::subset_t(subset, subset + n_subset);
    typedef tr::octree<point_t> tree_t;
    auto tree = tree_t::from_pointcloud(c, max_depth, crit, sub);
    if (depth_out) *depth_out = tree->depth();
    switch (traversal) {
        case 0: return octree_walk(tree_t::depth_traverse(tree->root()), rows, cap);
        case 1: return octree_walk(tree_t::breadth_traverse(tree->root()), rows, cap);
        case 2: return octree_walk(tree_t::leaf_traverse(tree->root()), rows, cap);
        case 3: return octree_walk(tree_t::branch_traverse(tree->root()), rows, cap);
        default: return octree_walk(tree_t::level_traverse(tree->root(), (uint8_t)level), rows, cap);
    }
}
// pointcloud::curvature(k, idx) = principal_curvatures(knn_inclusive(k, idx)) (pointcloud.hpp:200-204)
void ref_curvature(const float* pos, const float* nrm, uint32_t n, const uint32_t* query, uint32_t n_query, uint32_t k,
                   float* pc_min, float* pc_max, int32_t* nbr_out) {
    std::vector<float> z(3 * (size_t)n, 0.f);
    cloud_t::Ptr c = make_cloud(pos, nrm, z.data(), n);
    struct mode_guard { mode_guard() { pcl::eigen33_mode() = 1; } ~mode_guard() { pcl::eigen33_mode() = 0; } } guard;
    for (uint32_t w = 0; w < n_query; ++w) {
        auto ci = c->curvature(k, query[w]);
        pc_min[w] = ci.pc_min;
        pc_max[w] = ci.pc_max;
        if (nbr_out) {
            auto nn = c->knn_inclusive(k, query[w]).first;
            for (uint32_t j = 0; j < k; ++j) nbr_out[(size_t)w * k + j] = j < nn.size() ? nn[j] : -1;
        }
    }
}
float ref_resolution(const float* pos, uint32_t n) {
    std::vector<float> z(3 * (size_t)n, 0.f);
    return make_cloud(pos, z.data(), z.data(), n)->resolution();
}
}  // extern "C"
