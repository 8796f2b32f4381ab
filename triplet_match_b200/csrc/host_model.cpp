// host_model.cpp — host side of model<Point>::init (include/impl/model.hpp:16-167):
// finite filter, bbox / diameter, voxel-grid geometry, tangent subset, the two
// passes over all ordered tangent pairs (feature bounds, then the hash multimap)
// and the flattening of the multimap into the CSR the device probes.  The
// per-voxel 1-NN (model.hpp:81-94, the reference's OpenMP + kd-tree loop) runs on
// the GPU (tm_voxel_fill) when a context is given, else on an exact host grid
// search.  Compiled with -ffp-contract=off: same no-FMA float results as the
// reference build.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/triplet_match/tm_voxel_centre.h"

#include "../../include/tm_b200.h"
#include "../../include/tm_b200_host.h"
#include "../../include/triplet_match/tm_atan2f.h"

namespace {

struct v3 {
    float x, y, z;
};
inline float sum3(float a, float b, float c) { return a + (b + c); }
inline v3 sub(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(v3 a, v3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float norm(v3 a) { return sqrtf(dot(a, a)); }
inline v3 cross(v3 a, v3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float sqdist_seq(v3 a, v3 b) {  // FLANN L2_Simple accumulation order
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

// atan2f of the reference platform's libm, restated (include/triplet_match/tm_atan2f.h)
float atan2f_q1(float y, float x) { return tm_math::atan2f_libm(y, x); }
float angle(v3 a, v3 b) { return atan2f_q1(norm(cross(a, b)), fabsf(dot(a, b))); }

uint32_t discretize_range(float value, float mn, float range, uint32_t steps) {
    float nval = (value - mn) / range;
    if (nval < 0.f) return 0;
    if (nval >= 1.f) return steps - 1;
    return static_cast<uint32_t>(nval * steps);
}
uint32_t discretize_step(float value, float step) { return static_cast<uint32_t>(value / step); }



// exact nearest neighbour on a uniform bucket grid; lowest index wins ties
struct NNGrid {
    const float* pos;
    uint32_t stride, n;
    float lo[3], cs;
    int dim[3];
    std::vector<uint32_t> cell_off, cell_pts;

    v3 at(uint32_t i) const {
        const float* p = pos + (size_t)i * stride;
        return {p[0], p[1], p[2]};
    }
    void build(const float* p, uint32_t st, uint32_t nn, float cell) {
        pos = p; stride = st; n = nn;
        float hi[3];
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::numeric_limits<float>::max();
            hi[k] = std::numeric_limits<float>::lowest();
        }
        for (uint32_t i = 0; i < n; ++i) {
            v3 q = at(i);
            float c[3] = {q.x, q.y, q.z};
            for (int k = 0; k < 3; ++k)
                if (std::isfinite(c[k])) {
                    lo[k] = std::min(lo[k], c[k]);
                    hi[k] = std::max(hi[k], c[k]);
                }
        }
        cs = cell > 0 ? cell : 1.f;
        for (;;) {  // keep the bucket count bounded
            double cells = 1;
            for (int k = 0; k < 3; ++k) {
                dim[k] = std::max(1, (int)std::floor((hi[k] - lo[k]) / cs) + 1);
                cells *= dim[k];
            }
            if (cells <= 6.4e7) break;
            cs *= 1.5f;
        }
        size_t nc = (size_t)dim[0] * dim[1] * dim[2];
        cell_off.assign(nc + 1, 0);
        std::vector<uint32_t> cid(n);
        for (uint32_t i = 0; i < n; ++i) {
            cid[i] = cell_of(at(i));
            if (cid[i] != 0xffffffffu) ++cell_off[cid[i] + 1];
        }
        for (size_t c = 0; c < nc; ++c) cell_off[c + 1] += cell_off[c];
        cell_pts.resize(cell_off[nc]);
        std::vector<uint32_t> cur(cell_off.begin(), cell_off.end() - 1);
        for (uint32_t i = 0; i < n; ++i)
            if (cid[i] != 0xffffffffu) cell_pts[cur[cid[i]]++] = i;  // ascending inside a cell
    }
    int coord(float v, int k) const { return (int)std::floor((v - lo[k]) / cs); }
    uint32_t cell_of(v3 q) const {
        if (!std::isfinite(q.x) || !std::isfinite(q.y) || !std::isfinite(q.z)) return 0xffffffffu;
        int c[3] = {coord(q.x, 0), coord(q.y, 1), coord(q.z, 2)};
        for (int k = 0; k < 3; ++k) c[k] = std::min(std::max(c[k], 0), dim[k] - 1);
        return (uint32_t)(((size_t)c[2] * dim[1] + c[1]) * dim[0] + c[0]);
    }
    // nearest point to q, excluding index `skip` (0xffffffff = none)
    void nearest(v3 q, uint32_t skip, float& best, uint32_t& bi) const {
        best = std::numeric_limits<float>::max();
        bi = 0;
        int c[3] = {coord(q.x, 0), coord(q.y, 1), coord(q.z, 2)};
        int maxr = 0;
        for (int k = 0; k < 3; ++k) maxr = std::max(maxr, std::max(std::abs(c[k]), std::abs(c[k] - (dim[k] - 1))) + 1);
        for (int r = 0; r <= maxr; ++r) {
            for (int dz = -r; dz <= r; ++dz) {
                int z = c[2] + dz;
                if (z < 0 || z >= dim[2]) continue;
                for (int dy = -r; dy <= r; ++dy) {
                    int y = c[1] + dy;
                    if (y < 0 || y >= dim[1]) continue;
                    bool shell_yz = std::abs(dz) == r || std::abs(dy) == r;
                    int step = shell_yz ? 1 : 2 * r;
                    if (step == 0) step = 1;
                    for (int dx = -r; dx <= r; dx += step) {
                        int x = c[0] + dx;
                        if (x < 0 || x >= dim[0]) continue;
                        size_t cell = ((size_t)z * dim[1] + y) * dim[0] + x;
                        for (uint32_t t = cell_off[cell]; t < cell_off[cell + 1]; ++t) {
                            uint32_t i = cell_pts[t];
                            if (i == skip) continue;
                            float d = sqdist_seq(at(i), q);
                            if (d < best || (d == best && i < bi)) {
                                best = d;
                                bi = i;
                            }
                        }
                    }
                }
            }
            // everything outside ring r is at least r*cs away (with a rounding guard)
            double bound = (double)r * cs;
            if (best < std::numeric_limits<float>::max() && (double)best * 1.00001 < bound * bound) break;
        }
    }
};

}  // namespace

struct tm_hostmodel {
    std::vector<uint32_t> voxel;
    int32_t extents[3];
    float to_voxel[16];
    float resolution, diameter;
    float feat_min[4], feat_max[4];
    float distance_step_count, angle_step;
    std::vector<uint32_t> subset;
    // every multimap insertion in order (drop-in model::query rebuilds its own multimap)
    std::vector<uint32_t> entry_keys;   // 4 per entry
    std::vector<uint32_t> entry_pairs;  // 2 per entry
    // CSR, equal_range order, capped
    std::vector<uint32_t> keys, offsets, pairs;
    uint64_t n_entries = 0;
};

static thread_local std::string g_host_err;

extern "C" {

const char* tm_host_last_error(void) { return g_host_err.c_str(); }

float tm_host_resolution(const tm_cloud_view* cloud) {
    // include/impl/pointcloud.hpp:66-82: running mean (include/common:104-115) of the 1-NN distance
    if (!cloud || cloud->n < 2) return 0.f;
    // bucket size from the bbox volume / area heuristics: ~2 points per bucket edge
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t i = 0; i < cloud->n; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = cloud->pos[(size_t)i * cloud->stride + k];
            if (std::isfinite(v)) {
                lo[k] = std::min(lo[k], v);
                hi[k] = std::max(hi[k], v);
            }
        }
    double ext[3] = {std::max(1e-9f, hi[0] - lo[0]), std::max(1e-9f, hi[1] - lo[1]), std::max(1e-9f, hi[2] - lo[2])};
    std::sort(ext, ext + 3);
    // surface-like clouds: area ~ two largest extents
    float cell = (float)std::sqrt(ext[1] * ext[2] / std::max(1u, cloud->n)) * 2.f;
    NNGrid g;
    g.build(cloud->pos, cloud->stride, cloud->n, cell);
    float accum = 0.f;
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < cloud->n; ++i) {
        float best;
        uint32_t bi;
        g.nearest(g.at(i), i, best, bi);
        float val = sqrtf(best);
        accum = accum + (val - accum) / (++cnt);
    }
    return accum;
}

int tm_hostmodel_build(tm_ctx* ctx, const tm_cloud_view* c, const uint8_t* curv_ok,
                       float distance_step_count, float angle_step, float min_diameter_factor,
                       float max_diameter_factor, float resolution, uint32_t cap,
                       tm_hostmodel** out) {
    return tm_hostmodel_build_subset(ctx, c, nullptr, curv_ok, distance_step_count, angle_step, min_diameter_factor,
                                     max_diameter_factor, resolution, cap, out);
}

int tm_hostmodel_build_subset(tm_ctx* ctx, const tm_cloud_view* c, const uint8_t* in_subset, const uint8_t* curv_ok,
                              float distance_step_count, float angle_step, float min_diameter_factor,
                              float max_diameter_factor, float resolution, uint32_t cap, tm_hostmodel** out) {
    if (!c || !out || !c->pos || !c->nrm || !c->tgt || c->n == 0) {
        g_host_err = "tm_hostmodel_build: bad cloud";
        return TM_ERR_INVALID;
    }
    // the pair keys are packed k0 | k1 << 24 | k2 << 44 (host multimap order and the device table): distinct keys
    // must stay distinct
    if (!(distance_step_count >= 1.f) || distance_step_count > 16777216.f || !(angle_step > 0.f) ||
        3.14159274f / angle_step >= 1048576.f) {
        g_host_err = "tm_hostmodel_build: distance_step_count must be in [1, 2^24] and pi / angle_step below 2^20";
        return TM_ERR_INVALID;
    }
    auto P = [&](uint32_t i) { const float* p = c->pos + (size_t)i * c->stride; return v3{p[0], p[1], p[2]}; };
    auto N = [&](uint32_t i) { const float* p = c->nrm + (size_t)i * c->stride; return v3{p[0], p[1], p[2]}; };
    auto T = [&](uint32_t i) { const float* p = c->tgt + (size_t)i * c->stride; return v3{p[0], p[1], p[2]}; };
    tm_hostmodel* m = new tm_hostmodel();
    m->distance_step_count = distance_step_count;
    m->angle_step = angle_step;
    std::vector<uint32_t> all;
    for (uint32_t i = 0; i < c->n; ++i) {  // model.hpp:17-30: subset_ (all points when empty), finite ones
        if (in_subset && !in_subset[i]) continue;
        v3 p = P(i), n = N(i), t = T(i);
        bool fin = std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z) &&
                   std::isfinite(n.x) && std::isfinite(n.y) && std::isfinite(n.z) &&
                   std::isfinite(t.x) && std::isfinite(t.y) && std::isfinite(t.z);
        if (fin) all.push_back(i);
    }
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = std::numeric_limits<float>::max();
        hi[k] = std::numeric_limits<float>::lowest();
    }
    for (uint32_t i : all) {  // model.hpp:34-38
        v3 p = P(i);
        float q[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], q[k]);
            hi[k] = std::max(hi[k], q[k]);
        }
    }
    v3 range = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    m->diameter = norm(range);  // model.hpp:39
    m->resolution = resolution > 0.f ? resolution : tm_host_resolution(c);
    float half_res = 0.5f * m->resolution;  // model.hpp:45-46
    float ext[3] = {std::max(range.x / half_res, 1.f), std::max(range.y / half_res, 1.f),
                    std::max(range.z / half_res, 1.f)};
    float rg[3] = {range.x, range.y, range.z};
    const int margin = 5;
    float scale[3], trans[3];
    for (int k = 0; k < 3; ++k) {
        m->extents[k] = static_cast<int>(ext[k] + 2.f * margin);  // model.hpp:50
        scale[k] = rg[k] < 1e-5f ? 1.f : ext[k] / rg[k];          // model.hpp:52-55
        trans[k] = (scale[k] * (-lo[k]) + static_cast<float>(margin)) - 0.5f;  // :58-61
    }
    memset(m->to_voxel, 0, sizeof(m->to_voxel));
    for (int k = 0; k < 3; ++k) {
        m->to_voxel[k * 4 + k] = scale[k];
        m->to_voxel[12 + k] = trans[k];
    }
    m->to_voxel[15] = 1.f;
    const size_t cells = (size_t)m->extents[0] * m->extents[1] * m->extents[2];
    if (cells == 0 || cells >= (1ull << 31)) {
        g_host_err = "voxel grid has " + std::to_string(cells) + " cells (limit 2^31)";
        delete m;
        return TM_ERR_INVALID;
    }
    m->voxel.assign(cells, 0u);
    if (ctx) {  // model.hpp:81-94 on the GPU
        int rc = tm_voxel_fill(ctx, c, m->extents, m->to_voxel, m->voxel.data());
        if (rc) {
            g_host_err = std::string("tm_voxel_fill: ") + tm_last_error();
            delete m;
            return rc;
        }
    } else {
        NNGrid g;
        g.build(c->pos, c->stride, c->n, 2.f * m->resolution);
        const tm_centre_map cm = tm_voxel_centre_map(scale, trans);  // model.hpp:63 inverse(), :87 centre
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < m->extents[2]; ++k)
            for (int j = 0; j < m->extents[1]; ++j)
                for (int i = 0; i < m->extents[0]; ++i) {
                    v3 q = {tm_voxel_centre(cm.a[0], cm.b[0], i), tm_voxel_centre(cm.a[1], cm.b[1], j),
                            tm_voxel_centre(cm.a[2], cm.b[2], k)};
                    float best;
                    uint32_t bi;
                    g.nearest(q, 0xffffffffu, best, bi);
                    m->voxel[((size_t)k * m->extents[1] + j) * m->extents[0] + i] = bi;
                }
    }
    // tangent subset, model.hpp:96-99 (curv_ok replaces the PCL curvature ratio test)
    for (uint32_t i : all)
        if (norm(T(i)) > 0.7f && (!curv_ok || curv_ok[i])) m->subset.push_back(i);
    float lower_bound = m->diameter * min_diameter_factor;  // model.hpp:101-102
    float upper_bound = m->diameter * max_diameter_factor;
    float fmn[4], fmx[4];
    for (int k = 0; k < 4; ++k) {
        fmn[k] = std::numeric_limits<float>::max();
        fmx[k] = std::numeric_limits<float>::lowest();
    }
    auto pair_feature = [&](uint32_t i, uint32_t j, float f[4]) {  // model.hpp:105-113
        if (i == j) return false;
        v3 d1 = sub(P(j), P(i));
        float dist1 = norm(d1);
        v3 dn = {d1.x / dist1, d1.y / dist1, d1.z / dist1};
        if (dist1 < lower_bound || dist1 > upper_bound) return false;
        if (1.f - fabsf(dot(dn, T(i))) < 0.01f) return false;
        f[0] = dist1;  // feature.hpp:27 (same squaredNorm + sqrt)
        f[1] = angle(d1, T(i));
        f[2] = angle(d1, T(j));
        f[3] = f[0];
        return true;
    };
    // The O(T^2) pair enumeration runs on the device when a context is given (k_model.cu: same
    // filters, feature and discretisation; packed key per pair in insertion order); the host then
    // only does the integer bookkeeping.  TM_MODEL_PAIRS_HOST=1 forces the host loops (cross-check).
    const uint32_t Tn = (uint32_t)m->subset.size();
    const char* host_env = getenv("TM_MODEL_PAIRS_HOST");
    const uint32_t steps = static_cast<uint32_t>(distance_step_count);
    const bool on_device = ctx && Tn > 0 && (uint64_t)Tn * Tn <= (1ull << 28) && steps < (1u << 24) &&
                           !(host_env && atoi(host_env) != 0);
    std::vector<float> sp3, st3;
    if (on_device) {
        sp3.resize((size_t)Tn * 3);
        st3.resize((size_t)Tn * 3);
        for (uint32_t a = 0; a < Tn; ++a) {
            v3 p = P(m->subset[a]), t = T(m->subset[a]);
            sp3[3 * a] = p.x; sp3[3 * a + 1] = p.y; sp3[3 * a + 2] = p.z;
            st3[3 * a] = t.x; st3[3 * a + 1] = t.y; st3[3 * a + 2] = t.z;
        }
        float mn3[3], mx3[3];
        uint64_t n_pass = 0;
        int rc = tm_model_pair_bounds(ctx, sp3.data(), st3.data(), Tn, lower_bound, upper_bound, mn3, mx3, &n_pass);
        if (rc) {
            g_host_err = std::string("tm_model_pair_bounds: ") + tm_last_error();
            delete m;
            return rc;
        }
        for (int k = 0; k < 3; ++k) { fmn[k] = mn3[k]; fmx[k] = mx3[k]; }
        fmn[3] = fmn[0]; fmx[3] = fmx[0];
    } else {
        for (uint32_t i : m->subset)
            for (uint32_t j : m->subset) {
                float f[4];
                if (!pair_feature(i, j, f)) continue;
                for (int k = 0; k < 4; ++k) {
                    fmn[k] = std::min(fmn[k], f[k]);
                    fmx[k] = std::max(fmx[k], f[k]);
                }
            }
    }
    // valid_bounds(bounds, ., ., 0, 1), feature.hpp:90-114
    float d0 = fmx[0] - fmn[0], d3 = fmx[3] - fmn[3];
    float nmn0 = fmn[0] + 0.0f * d0, nmx0 = fmn[0] + 1.f * d0;
    float nmn3 = fmn[3] + 0.0f * d3, nmx3 = fmn[3] + 1.f * d3;
    fmn[0] = nmn0; fmx[0] = nmx0; fmn[3] = nmn3; fmx[3] = nmx3;
    memcpy(m->feat_min, fmn, 16);
    memcpy(m->feat_max, fmx, 16);
    const float pi = static_cast<float>(M_PI);
    // per key: the (i, j) values in insertion order.  equal_range of libstdc++'s unordered_multimap
    // walks equal keys newest-first (each insert lands at the front of its group), so the table keeps
    // the LAST `cap` insertions of a key, reversed — checked against the real container in the tests.
    std::unordered_map<uint64_t, uint32_t> key_id;
    std::vector<std::array<uint32_t, 4>> uk;
    std::vector<std::vector<uint32_t>> vals;  // flattened (i, j) per key
    auto add_entry = [&](uint32_t k0, uint32_t k1, uint32_t k2, uint32_t i, uint32_t j) {
        const uint64_t pk = (uint64_t)k0 | ((uint64_t)k1 << 24) | ((uint64_t)k2 << 44);
        auto it = key_id.find(pk);
        uint32_t id;
        if (it == key_id.end()) {
            id = (uint32_t)uk.size();
            key_id.emplace(pk, id);
            uk.push_back({k0, k1, k2, k0});
            vals.emplace_back();
        } else {
            id = it->second;
        }
        vals[id].push_back(i);
        vals[id].push_back(j);
        const uint32_t k4[4] = {k0, k1, k2, k0};
        m->entry_keys.insert(m->entry_keys.end(), k4, k4 + 4);
        m->entry_pairs.push_back(i);
        m->entry_pairs.push_back(j);
    };
    if (on_device) {
        std::vector<uint64_t> pk((size_t)Tn * Tn);
        int rc = tm_model_pair_keys(ctx, sp3.data(), st3.data(), Tn, lower_bound, upper_bound, fmn[0], fmx[0], steps,
                                    angle_step, pk.data());
        if (rc) {
            g_host_err = std::string("tm_model_pair_keys: ") + tm_last_error();
            delete m;
            return rc;
        }
        for (uint32_t a = 0; a < Tn; ++a)
            for (uint32_t b = 0; b < Tn; ++b) {
                const uint64_t k = pk[(size_t)a * Tn + b];
                if (k == ~0ull) continue;
                add_entry((uint32_t)(k & 0xffffffu), (uint32_t)((k >> 24) & 0xfffffu), (uint32_t)(k >> 44), m->subset[a],
                          m->subset[b]);
            }
    } else {
        for (uint32_t i : m->subset)  // model.hpp:125-149
            for (uint32_t j : m->subset) {
                float f[4];
                if (!pair_feature(i, j, f)) continue;
                if (f[0] < fmn[0] || f[0] > fmx[0]) continue;  // valid(), feature.hpp:48-88
                if (!((f[1] >= 0.f && f[1] <= pi) && (f[2] >= 0.f && f[2] <= pi))) continue;
                float diag0 = fmx[0] - fmn[0];
                const uint32_t k0 = discretize_range(f[0], fmn[0], diag0, steps);
                // k3 = discretize_range(f[3], ...) with f[3] == f[0] and identical bounds: equals k0
                add_entry(k0, discretize_step(f[1], angle_step), discretize_step(f[2], angle_step), i, j);
            }
    }
    m->n_entries = m->entry_pairs.size() / 2;
    // flatten: keys in lexicographic order (deterministic), values newest-first, capped
    std::vector<uint32_t> order(uk.size());
    for (uint32_t k = 0; k < order.size(); ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return uk[a] < uk[b]; });
    m->offsets.push_back(0);
    for (uint32_t id : order) {
        const std::vector<uint32_t>& v = vals[id];
        const size_t n_val = v.size() / 2;
        const size_t take = cap ? std::min<size_t>(cap, n_val) : n_val;
        for (size_t e = 0; e < take; ++e) {
            m->pairs.push_back(v[2 * (n_val - 1 - e)]);
            m->pairs.push_back(v[2 * (n_val - 1 - e) + 1]);
        }
        m->keys.insert(m->keys.end(), uk[id].begin(), uk[id].end());
        m->offsets.push_back((uint32_t)(m->pairs.size() / 2));
    }
    *out = m;
    return TM_OK;
}

void tm_hostmodel_destroy(tm_hostmodel* m) { delete m; }

// ---- model blob: what model::init produced, serialised (the reference rebuilds the hash table and
// the grid on every run, SURVEY section 5 "checkpoint / resume: none").  Layout, little endian:
//   "TMB200M\0" | u32 version = 1 | u32 n_cloud_points | 16 header floats/ints | 7 x (u64 count + payload)
//   | u64 FNV-1a of everything before it.
namespace {
const char kBlobMagic[8] = {'T', 'M', 'B', '2', '0', '0', 'M', '\0'};
struct Fnv {
    uint64_t h = 1469598103934665603ull;
    void add(const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    }
};
struct BlobWriter {
    FILE* f;
    Fnv fnv;
    bool ok = true;
    void put(const void* p, size_t n) {
        if (n && fwrite(p, 1, n, f) != n) ok = false;
        fnv.add(p, n);
    }
    void vec(const std::vector<uint32_t>& v) {
        uint64_t n = v.size();
        put(&n, 8);
        put(v.data(), n * 4);
    }
};
struct BlobReader {
    FILE* f;
    Fnv fnv;
    bool ok = true;
    void get(void* p, size_t n) {
        if (n && fread(p, 1, n, f) != n) { ok = false; return; }
        fnv.add(p, n);
    }
    void vec(std::vector<uint32_t>& v, uint64_t limit) {
        uint64_t n = 0;
        get(&n, 8);
        if (!ok || n > limit) { ok = false; return; }
        v.resize(n);
        get(v.data(), n * 4);
    }
};
}  // namespace

int tm_hostmodel_save(const tm_hostmodel* m, uint32_t n_cloud_points, const char* path) {
    if (!m || !path) { g_host_err = "tm_hostmodel_save: null argument"; return TM_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { g_host_err = std::string("tm_hostmodel_save: cannot write '") + path + "'"; return TM_ERR_INVALID; }
    BlobWriter w{f};
    const uint32_t version = 1;
    w.put(kBlobMagic, 8);
    w.put(&version, 4);
    w.put(&n_cloud_points, 4);
    w.put(m->extents, sizeof(m->extents));
    w.put(m->to_voxel, sizeof(m->to_voxel));
    w.put(&m->resolution, 4); w.put(&m->diameter, 4);
    w.put(m->feat_min, 16); w.put(m->feat_max, 16);
    w.put(&m->distance_step_count, 4); w.put(&m->angle_step, 4);
    w.put(&m->n_entries, 8);
    w.vec(m->voxel); w.vec(m->subset); w.vec(m->entry_keys); w.vec(m->entry_pairs);
    w.vec(m->keys); w.vec(m->offsets); w.vec(m->pairs);
    const uint64_t sum = w.fnv.h;
    if (fwrite(&sum, 1, 8, f) != 8) w.ok = false;
    if (fclose(f) != 0) w.ok = false;
    if (!w.ok) { g_host_err = std::string("tm_hostmodel_save: write error on '") + path + "'"; return TM_ERR_INVALID; }
    return TM_OK;
}

int tm_hostmodel_load(const char* path, uint32_t n_cloud_points, tm_hostmodel** out) {
    if (!path || !out) { g_host_err = "tm_hostmodel_load: null argument"; return TM_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { g_host_err = std::string("tm_hostmodel_load: cannot open '") + path + "'"; return TM_ERR_INVALID; }
    tm_hostmodel* m = new tm_hostmodel();
    BlobReader r{f};
    auto bad = [&](const std::string& why) {
        fclose(f);
        delete m;
        g_host_err = "tm_hostmodel_load: '" + std::string(path) + "': " + why;
        return TM_ERR_INVALID;
    };
    char magic[8];
    uint32_t version = 0, n_pts = 0;
    r.get(magic, 8); r.get(&version, 4); r.get(&n_pts, 4);
    if (!r.ok || memcmp(magic, kBlobMagic, 8) != 0) return bad("not a model blob");
    if (version != 1) return bad("unsupported version " + std::to_string(version));
    if (n_pts != n_cloud_points) return bad("built for a cloud of " + std::to_string(n_pts) + " points, not " + std::to_string(n_cloud_points));
    r.get(m->extents, sizeof(m->extents));
    r.get(m->to_voxel, sizeof(m->to_voxel));
    r.get(&m->resolution, 4); r.get(&m->diameter, 4);
    r.get(m->feat_min, 16); r.get(m->feat_max, 16);
    r.get(&m->distance_step_count, 4); r.get(&m->angle_step, 4);
    r.get(&m->n_entries, 8);
    const uint64_t lim = 1ull << 31;
    r.vec(m->voxel, lim); r.vec(m->subset, lim); r.vec(m->entry_keys, 4 * lim); r.vec(m->entry_pairs, 2 * lim);
    r.vec(m->keys, lim); r.vec(m->offsets, lim); r.vec(m->pairs, lim);
    if (!r.ok) return bad("truncated");
    uint64_t sum = 0;
    if (fread(&sum, 1, 8, f) != 8 || sum != r.fnv.h) return bad("checksum mismatch");
    // structural checks: everything the device upload will index with
    bool sane = m->extents[0] > 0 && m->extents[1] > 0 && m->extents[2] > 0 &&
                m->voxel.size() == (size_t)m->extents[0] * m->extents[1] * m->extents[2] &&
                m->keys.size() % 4 == 0 && m->offsets.size() == m->keys.size() / 4 + 1 && m->pairs.size() % 2 == 0 &&
                (m->offsets.empty() || m->offsets.back() == m->pairs.size() / 2) &&
                m->entry_keys.size() == 4 * m->n_entries && m->entry_pairs.size() == 2 * m->n_entries;
    for (size_t i = 0; sane && i < m->voxel.size(); ++i) sane = m->voxel[i] < n_pts;
    for (size_t i = 0; sane && i < m->pairs.size(); ++i) sane = m->pairs[i] < n_pts;
    for (size_t i = 0; sane && i < m->subset.size(); ++i) sane = m->subset[i] < n_pts;
    for (size_t i = 0; sane && i + 1 < m->offsets.size(); ++i) sane = m->offsets[i] <= m->offsets[i + 1];
    if (!sane) return bad("inconsistent contents");
    fclose(f);
    *out = m;
    return TM_OK;
}

void tm_hostmodel_desc(const tm_hostmodel* m, tm_model_desc* d) {
    memset(d, 0, sizeof(*d));
    d->voxel = m->voxel.data();
    memcpy(d->extents, m->extents, sizeof(d->extents));
    memcpy(d->to_voxel, m->to_voxel, sizeof(d->to_voxel));
    d->resolution = m->resolution;
    d->diameter = m->diameter;
    d->keys = m->keys.data();
    d->offsets = m->offsets.data();
    d->pairs = m->pairs.data();
    d->n_keys = (uint32_t)(m->keys.size() / 4);
    memcpy(d->feat_min, m->feat_min, 16);
    memcpy(d->feat_max, m->feat_max, 16);
    d->distance_step_count = m->distance_step_count;
    d->angle_step = m->angle_step;
}
void tm_hostmodel_counts(const tm_hostmodel* m, uint64_t* n_subset, uint64_t* n_entries,
                         uint64_t* n_keys, uint64_t* n_kept) {
    if (n_subset) *n_subset = m->subset.size();
    if (n_entries) *n_entries = m->n_entries;
    if (n_keys) *n_keys = m->keys.size() / 4;
    if (n_kept) *n_kept = m->pairs.size() / 2;
}
const uint32_t* tm_hostmodel_subset(const tm_hostmodel* m) { return m->subset.data(); }
const uint32_t* tm_hostmodel_entry_keys(const tm_hostmodel* m) { return m->entry_keys.data(); }
const uint32_t* tm_hostmodel_entry_pairs(const tm_hostmodel* m) { return m->entry_pairs.data(); }

int tm_model_create(tm_ctx* ctx, const tm_cloud_view* cloud, const tm_hostmodel* hm, tm_model** out) {
    if (!hm) {
        g_host_err = "Cannot query uninitialized model";  // include/impl/model.hpp:171-173
        return TM_ERR_UNINITIALIZED;
    }
    tm_model_desc d;
    tm_hostmodel_desc(hm, &d);
    return tm_model_upload(ctx, cloud, &d, out);
}

}  // extern "C"
