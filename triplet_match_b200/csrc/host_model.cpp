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
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

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

uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint32_t murmur4(const uint32_t* key) {  // include/impl/discretize.hpp:10-45
    uint32_t h1 = 42u;
    for (int i = 0; i < 4; ++i) {
        uint32_t k1 = key[i];
        k1 *= 0xcc9e2d51u;
        k1 = rotl32(k1, 15);
        k1 *= 0x1b873593u;
        h1 ^= k1;
        h1 = rotl32(h1, 13);
        h1 = h1 * 5u + 0xe6546b64u;
    }
    h1 ^= 16u;
    h1 ^= h1 >> 16;
    h1 *= 0x85ebca6bu;
    h1 ^= h1 >> 13;
    h1 *= 0xc2b2ae35u;
    h1 ^= h1 >> 16;
    return h1;
}
uint32_t discretize_range(float value, float mn, float range, uint32_t steps) {
    float nval = (value - mn) / range;
    if (nval < 0.f) return 0;
    if (nval >= 1.f) return steps - 1;
    return static_cast<uint32_t>(nval * steps);
}
uint32_t discretize_step(float value, float step) { return static_cast<uint32_t>(value / step); }

struct key4 {
    uint32_t k[4];
    bool operator==(const key4& o) const { return !memcmp(k, o.k, 16); }
};
struct key4_hash {
    size_t operator()(const key4& k) const { return murmur4(k.k); }
};
typedef std::unordered_multimap<key4, std::pair<uint32_t, uint32_t>, key4_hash> hash_map_t;

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
    if (!c || !out || !c->pos || !c->nrm || !c->tgt || c->n == 0) {
        g_host_err = "tm_hostmodel_build: bad cloud";
        return TM_ERR_INVALID;
    }
    auto P = [&](uint32_t i) { const float* p = c->pos + (size_t)i * c->stride; return v3{p[0], p[1], p[2]}; };
    auto N = [&](uint32_t i) { const float* p = c->nrm + (size_t)i * c->stride; return v3{p[0], p[1], p[2]}; };
    auto T = [&](uint32_t i) { const float* p = c->tgt + (size_t)i * c->stride; return v3{p[0], p[1], p[2]}; };
    tm_hostmodel* m = new tm_hostmodel();
    m->distance_step_count = distance_step_count;
    m->angle_step = angle_step;
    std::vector<uint32_t> all;
    for (uint32_t i = 0; i < c->n; ++i) {  // model.hpp:24-30
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
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < m->extents[2]; ++k)
            for (int j = 0; j < m->extents[1]; ++j)
                for (int i = 0; i < m->extents[0]; ++i) {
                    v3 q = {((float)i - trans[0]) / scale[0], ((float)j - trans[1]) / scale[1],
                            ((float)k - trans[2]) / scale[2]};
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
    for (uint32_t i : m->subset)
        for (uint32_t j : m->subset) {
            float f[4];
            if (!pair_feature(i, j, f)) continue;
            for (int k = 0; k < 4; ++k) {
                fmn[k] = std::min(fmn[k], f[k]);
                fmx[k] = std::max(fmx[k], f[k]);
            }
        }
    // valid_bounds(bounds, ., ., 0, 1), feature.hpp:90-114
    float d0 = fmx[0] - fmn[0], d3 = fmx[3] - fmn[3];
    float nmn0 = fmn[0] + 0.0f * d0, nmx0 = fmn[0] + 1.f * d0;
    float nmn3 = fmn[3] + 0.0f * d3, nmx3 = fmn[3] + 1.f * d3;
    fmn[0] = nmn0; fmx[0] = nmx0; fmn[3] = nmn3; fmx[3] = nmx3;
    memcpy(m->feat_min, fmn, 16);
    memcpy(m->feat_max, fmx, 16);
    hash_map_t map;
    const uint32_t steps = static_cast<uint32_t>(distance_step_count);
    const float pi = static_cast<float>(M_PI);
    for (uint32_t i : m->subset)  // model.hpp:125-149
        for (uint32_t j : m->subset) {
            float f[4];
            if (!pair_feature(i, j, f)) continue;
            if (f[0] < fmn[0] || f[0] > fmx[0]) continue;  // valid(), feature.hpp:48-88
            if (!((f[1] >= 0.f && f[1] <= pi) && (f[2] >= 0.f && f[2] <= pi))) continue;
            key4 k;
            float diag0 = fmx[0] - fmn[0];
            k.k[0] = discretize_range(f[0], fmn[0], diag0, steps);
            k.k[1] = discretize_step(f[1], angle_step);
            k.k[2] = discretize_step(f[2], angle_step);
            k.k[3] = discretize_range(f[3], fmn[0], diag0, steps);
            map.insert({k, {i, j}});
            m->entry_keys.insert(m->entry_keys.end(), k.k, k.k + 4);
            m->entry_pairs.push_back(i);
            m->entry_pairs.push_back(j);
        }
    m->n_entries = map.size();
    // flatten: keys in lexicographic order (deterministic), values in equal_range order, capped
    std::vector<std::array<uint32_t, 4>> uk;
    for (auto it = map.begin(); it != map.end(); it = map.equal_range(it->first).second)
        uk.push_back({it->first.k[0], it->first.k[1], it->first.k[2], it->first.k[3]});
    std::sort(uk.begin(), uk.end());
    m->offsets.push_back(0);
    for (auto& k : uk) {
        key4 kk{{k[0], k[1], k[2], k[3]}};
        auto r = map.equal_range(kk);
        uint32_t cnt = 0;
        for (auto e = r.first; e != r.second; ++e) {
            if (cap && cnt >= cap) break;
            m->pairs.push_back(e->second.first);
            m->pairs.push_back(e->second.second);
            ++cnt;
        }
        m->keys.insert(m->keys.end(), k.begin(), k.end());
        m->offsets.push_back((uint32_t)(m->pairs.size() / 2));
    }
    *out = m;
    return TM_OK;
}

void tm_hostmodel_destroy(tm_hostmodel* m) { delete m; }

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
