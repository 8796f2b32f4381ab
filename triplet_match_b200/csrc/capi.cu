// capi.cu — implementation of include/tm_b200.h: resident model/scene state,
// the stage calls with host buffers, and the resident query pipeline.  Host
// logic only; every data-parallel step is one of the kernels in k_*.cu.
#include "../../include/tm_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "tm_kernels.cuh"

namespace tmk {
std::atomic<unsigned long long> g_launch_count{0};
}
using namespace tmk;

// ------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(TM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)
#define REQUIRE(cond, msg)                          \
    do {                                            \
        if (!(cond)) return fail(TM_ERR_INVALID, msg); \
    } while (0)
#define TRY(call)              \
    do {                       \
        int rc_ = (call);      \
        if (rc_) return rc_;   \
    } while (0)

// development knobs (environment), read once per process; none of them changes results
struct Knobs {
    bool occ = true;               // TM_OCC=0: never use the block-occupancy mask
    int fused_grid = -1;           // TM_FUSED_GRID=0/1: force the unfused / fused voxel grid
    int score_grid = 0;            // TM_SCORE_GRID=n: CTAs of the scoring kernel
    bool score_stats = false;      // TM_SCORE_STATS=1: cull / inlier statistics of the scoring kernel on stderr
    int scorer = 8;                // TM_SCORER=7: the fused count+score kernel everywhere (A/B against the count-only
                                   // packed-FP32 kernel + lazy score, which is the default where scores are not asked for)
};
static const Knobs& knobs() {
    static const Knobs k = [] {
        Knobs v;
        if (const char* e = getenv("TM_OCC")) v.occ = atoi(e) != 0;
        if (const char* e = getenv("TM_FUSED_GRID")) v.fused_grid = atoi(e) != 0 ? 1 : 0;
        if (const char* e = getenv("TM_SCORE_GRID")) v.score_grid = std::max(1, atoi(e));
        v.score_stats = getenv("TM_SCORE_STATS") != nullptr;
        if (const char* e = getenv("TM_SCORER")) v.scorer = atoi(e);
        return v;
    }();
    return k;
}

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return TM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(bytes, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess)
            return fail(TM_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(want) +
                                         "): " + cudaGetErrorString(e));
        cap = want;
        return TM_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
};

struct tm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    DevBuf flush;
    DevBuf scratch[13];  // stage-call scratch, grow-only
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    int score_bps[2][2] = {{0, 0}, {0, 0}};  // resident CTAs per SM of the scoring kernel [fused][with_score]
    int count_bps[2] = {0, 0};               // the same for the count-only kernel [fused]
};

struct OccMask {  // block-occupancy mask of one distance threshold (k_util.cu occupancy_kernel)
    float thres = -1.f;
    DevBuf bits;
    bool useful = false;  // enough empty blocks to pay for the extra look-up
};
struct tm_model {
    tm_ctx* ctx;
    DevBuf pos, nrm, tgt, voxel, vcell, vref, slots, hits;
    ModelDev dev;
    float centre[3];
    float half_diag;
    bool fused;
    OccMask occ[2];  // scoring threshold and the ICP one (2 x dist_thres); replaced round-robin
    int occ_next = 0;
};

struct tm_scene {
    tm_ctx* ctx;
    DevBuf pos, nrm, tgt, mask_tmp, seg_lo, seg_hi;
    CloudDev dev;
};

static int bind(tm_ctx* c) {
    CU(cudaSetDevice(c->device));
    return TM_OK;
}
// ModelDev for kernels that test against `thres`: the resident description plus, when it pays, the
// block-occupancy mask of that threshold (built on first use, cached per model).  TM_OCC=0 disables.
static int model_dev_for(tm_ctx* c, tm_model* m, float thres, ModelDev* out) {
    *out = m->dev;
    if (!knobs().occ || !(thres >= 0.f)) return TM_OK;
    OccMask* hit = nullptr;
    for (OccMask& o : m->occ)
        if (o.thres == thres) hit = &o;
    if (!hit) {
        hit = &m->occ[m->occ_next];
        m->occ_next ^= 1;
        const ModelDev& d = m->dev;
        const int ob = (1 << OCC_SHIFT) - 1;
        const int obx = (d.ex + ob) >> OCC_SHIFT, oby = (d.ey + ob) >> OCC_SHIFT, obz = (d.ez + ob) >> OCC_SHIFT;
        const size_t nb = (size_t)obx * oby * obz, words = (nb + 31) / 32;
        TRY(hit->bits.ensure(words * 4));
        CU(cudaMemsetAsync(hit->bits.p, 0, words * 4, c->stream));
        const double D = std::sqrt(1.0 / ((double)d.sx * d.sx) + 1.0 / ((double)d.sy * d.sy) + 1.0 / ((double)d.sz * d.sz));
        const double reach = ((double)thres + D * 1.001) * 1.0001 + 1e-30;
        launch_occupancy(c->stream, d.voxel, d.cloud.pos, d.ex, d.ey, d.ez, d.sx, d.sy, d.sz, d.tx, d.ty, d.tz,
                         (float)(reach * reach * 1.00001), obx, oby, hit->bits.as<uint32_t>());
        CU(cudaGetLastError());
        std::vector<uint32_t> hbits(words);
        CU(cudaMemcpyAsync(hbits.data(), hit->bits.p, words * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        size_t set = 0;
        for (uint32_t w : hbits) set += (size_t)__builtin_popcount(w);
        hit->thres = thres;
        hit->useful = set * 2 < nb;  // at least half of the blocks are empty
    }
    if (hit->useful) {
        out->occ = hit->bits.as<uint32_t>();
        out->obx = (m->dev.ex + (1 << OCC_SHIFT) - 1) >> OCC_SHIFT;
        out->oby = (m->dev.ey + (1 << OCC_SHIFT) - 1) >> OCC_SHIFT;
    }
    return TM_OK;
}
static int pinned_ensure(tm_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_cap) return TM_OK;
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->pinned_cap = 0;
    CU(cudaMallocHost(&c->pinned, bytes));
    c->pinned_cap = bytes;
    return TM_OK;
}

// `dist > thres` with dist = sqrtf(sq) (scene.hpp:464-465) <=> sq > S, where S is the
// largest float whose correctly rounded square root is <= thres.
static float sq_threshold(float thres) {
    if (!(thres >= 0.f)) return -1.f;
    float c = thres * thres;
    while (sqrtf(c) > thres) c = nextafterf(c, 0.f);
    for (;;) {
        float n = nextafterf(c, INFINITY);
        if (std::isinf(n) || sqrtf(n) > thres) break;
        c = n;
    }
    return c;
}

extern "C" {

const char* tm_last_error(void) { return g_err.c_str(); }
const char* tm_version(void) { return "triplet_match_b200 0.1 (sm_100a)"; }

int tm_ctx_create(int device, tm_ctx** out) {
    REQUIRE(out, "tm_ctx_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(TM_ERR_CUDA, std::string("no CUDA device: ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                     " (this library has no CPU fallback)");
    REQUIRE(device >= 0 && device < n, "tm_ctx_create: device out of range");
    tm_ctx* c = new tm_ctx();
    c->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        if (c->ev0) cudaEventDestroy(c->ev0);
        if (c->ev1) cudaEventDestroy(c->ev1);
        if (c->stream) cudaStreamDestroy(c->stream);
        delete c;
        return fail(TM_ERR_CUDA, std::string("tm_ctx_create: ") + cudaGetErrorString(e));
    }
    *out = c;
    return TM_OK;
}
void tm_ctx_destroy(tm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->flush.release();
    for (auto& s : c->scratch) s.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
}
int tm_ctx_sync(tm_ctx* c) {
    REQUIRE(c, "null ctx");
    TRY(bind(c));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}
void* tm_ctx_stream(tm_ctx* c) { return c ? (void*)c->stream : nullptr; }
int tm_ctx_sm_count(tm_ctx* c) { return c ? c->sm_count : 0; }
int tm_timer_start(tm_ctx* c) {
    REQUIRE(c, "null ctx");
    TRY(bind(c));
    CU(cudaEventRecord(c->ev0, c->stream));
    return TM_OK;
}
int tm_timer_stop(tm_ctx* c, float* ms) {
    REQUIRE(c && ms, "null arg");
    TRY(bind(c));
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return TM_OK;
}
int tm_ctx_flush_l2(tm_ctx* c) {
    REQUIRE(c, "null ctx");
    TRY(bind(c));
    const size_t bytes = 256ull << 20;
    TRY(c->flush.ensure(bytes));
    launch_flush(c->stream, c->flush.as<float4>(), bytes / sizeof(float4), 1.f);
    CU(cudaGetLastError());
    return TM_OK;
}
uint64_t tm_ctx_kernel_launches(tm_ctx*) { return g_launch_count.load(std::memory_order_relaxed); }

// ------------------------------------------------------------- cloud upload
static int upload_cloud(tm_ctx* c, const tm_cloud_view* v, const uint8_t* flags, int model_mode,
                        DevBuf& pos, DevBuf& nrm, DevBuf& tgt) {
    REQUIRE(v && v->pos && v->nrm && v->tgt, "cloud view has null arrays");
    REQUIRE(v->stride >= 3, "cloud stride must be >= 3 floats");
    const uint32_t n = v->n;
    TRY(pos.ensure(sizeof(float4) * (size_t)std::max(n, 1u)));
    TRY(nrm.ensure(sizeof(float4) * (size_t)std::max(n, 1u)));
    TRY(tgt.ensure(sizeof(float4) * (size_t)std::max(n, 1u)));
    if (!n) return TM_OK;
    // The three arrays may alias one AoS buffer (PointSurfel) or be separate packed
    // arrays; copy the covering byte range of each verbatim, then pack on the device.
    const size_t span = ((size_t)(n - 1) * v->stride + 3) * sizeof(float);
    DevBuf &rp = c->scratch[0], &rn = c->scratch[1], &rt = c->scratch[2], &rf = c->scratch[3];
    const float *dp, *dn, *dt;
    const float* lo = std::min(v->pos, std::min(v->nrm, v->tgt));
    const float* hi = std::max(v->pos, std::max(v->nrm, v->tgt));
    if (v->stride > 3 && (size_t)(hi - lo) < v->stride) {
        // interleaved: one copy
        size_t bytes = span + (size_t)(hi - lo) * sizeof(float);
        TRY(rp.ensure(bytes));
        CU(cudaMemcpyAsync(rp.p, lo, bytes, cudaMemcpyHostToDevice, c->stream));
        dp = rp.as<float>() + (v->pos - lo);
        dn = rp.as<float>() + (v->nrm - lo);
        dt = rp.as<float>() + (v->tgt - lo);
    } else {
        TRY(rp.ensure(span));
        TRY(rn.ensure(span));
        TRY(rt.ensure(span));
        CU(cudaMemcpyAsync(rp.p, v->pos, span, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(rn.p, v->nrm, span, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(rt.p, v->tgt, span, cudaMemcpyHostToDevice, c->stream));
        dp = rp.as<float>();
        dn = rn.as<float>();
        dt = rt.as<float>();
    }
    const uint8_t* dflags = nullptr;
    if (flags) {
        TRY(rf.ensure(n));
        CU(cudaMemcpyAsync(rf.p, flags, n, cudaMemcpyHostToDevice, c->stream));
        dflags = rf.as<uint8_t>();
    }
    launch_pack_cloud(c->stream, dp, dn, dt, v->stride, n, dflags, model_mode, pos.as<float4>(),
                      nrm.as<float4>(), tgt.as<float4>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));  // host buffers may be reused by the caller
    return TM_OK;
}

static int check_to_voxel(const float* t, float s[3], float tr[3]) {
    // column-major; must be diag(s) + translation with last row (0,0,0,1)
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            float v = t[col * 4 + row];
            bool diag = row == col, transl = col == 3 && row < 3;
            if (!diag && !transl && v != 0.f)
                return fail(TM_ERR_INVALID, "to_voxel must be diagonal + translation");
        }
    if (t[15] != 1.f) return fail(TM_ERR_INVALID, "to_voxel[3][3] must be 1");
    for (int k = 0; k < 3; ++k) {
        s[k] = t[k * 4 + k];
        tr[k] = t[12 + k];
    }
    return TM_OK;
}

int tm_model_upload(tm_ctx* c, const tm_cloud_view* cloud, const tm_model_desc* d, tm_model** out) {
    REQUIRE(c && cloud && d && out, "tm_model_upload: null argument");
    TRY(bind(c));
    REQUIRE(cloud->n > 0, "model cloud is empty");
    REQUIRE(d->voxel, "model not initialised: voxel grid missing");
    REQUIRE(d->extents[0] > 0 && d->extents[1] > 0 && d->extents[2] > 0, "bad extents");
    const size_t cells = (size_t)d->extents[0] * d->extents[1] * d->extents[2];
    REQUIRE(cells < (1ull << 31), "voxel grid too large for 32-bit linear index");
    REQUIRE(d->n_keys < (1u << 30), "too many hash keys");
    REQUIRE(d->n_keys == 0 || (d->keys && d->offsets && d->pairs), "hash table arrays missing");
    float s[3], tr[3];
    TRY(check_to_voxel(d->to_voxel, s, tr));
    tm_model* m = new tm_model();
    m->ctx = c;
    int rc = upload_cloud(c, cloud, nullptr, 1, m->pos, m->nrm, m->tgt);
    if (rc) {
        tm_model_destroy(m);
        return rc;
    }
    auto bail = [&](int code) {
        tm_model_destroy(m);
        return code;
    };
    if ((rc = m->voxel.ensure(cells * sizeof(uint32_t)))) return bail(rc);
    for (size_t i = 0; i < cells; ++i)
        if (d->voxel[i] >= cloud->n) return bail(fail(TM_ERR_INVALID, "voxel entry out of range"));
    auto cuda_bail = [&](cudaError_t e, const char* what) {
        return bail(fail(TM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)));
    };
    cudaError_t ce = cudaMemcpyAsync(m->voxel.p, d->voxel, cells * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (ce != cudaSuccess) return cuda_bail(ce, "voxel grid upload");
    // hash table: open addressing over the unique keys
    uint32_t cap = 16;
    while (cap < 2u * std::max(d->n_keys, 1u)) cap <<= 1;
    std::vector<HashSlot> slots(cap);
    memset(slots.data(), 0, sizeof(HashSlot) * cap);
    uint32_t n_hits = d->n_keys ? d->offsets[d->n_keys] : 0;
    for (uint32_t k = 0; k < d->n_keys; ++k) {
        const uint32_t* key = d->keys + 4 * (size_t)k;
        uint32_t cnt = d->offsets[k + 1] - d->offsets[k];
        if (!cnt) continue;
        uint32_t h = murmur4(key[0], key[1], key[2], key[3]) & (cap - 1);
        while (slots[h].count) h = (h + 1) & (cap - 1);
        for (int a = 0; a < 4; ++a) slots[h].k[a] = key[a];
        slots[h].begin = d->offsets[k];
        slots[h].count = cnt;
    }
    for (size_t i = 0; i < 2 * (size_t)n_hits; ++i)
        if (d->pairs[i] >= cloud->n) return bail(fail(TM_ERR_INVALID, "hash pair out of range"));
    if ((rc = m->slots.ensure(sizeof(HashSlot) * cap))) return bail(rc);
    if ((rc = m->hits.ensure(sizeof(uint2) * (size_t)std::max(n_hits, 1u)))) return bail(rc);
    ce = cudaMemcpyAsync(m->slots.p, slots.data(), sizeof(HashSlot) * cap, cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess && n_hits)
        ce = cudaMemcpyAsync(m->hits.p, d->pairs, sizeof(uint2) * (size_t)n_hits, cudaMemcpyHostToDevice, c->stream);
    if (ce != cudaSuccess) return cuda_bail(ce, "hash table upload");
    // fused grid (cell -> model point) when it stays L2-sized
    m->fused = cells * sizeof(float4) <= (96ull << 20);
    if (knobs().fused_grid >= 0) m->fused = knobs().fused_grid != 0;
    if (m->fused) {
        if ((rc = m->vcell.ensure(cells * sizeof(float4)))) return bail(rc);
        launch_fuse_grid(c->stream, m->voxel.as<uint32_t>(), cells, m->pos.as<float4>(), m->vcell.as<float4>());
    }
    if ((rc = m->vref.ensure((size_t)cloud->n * sizeof(float4)))) return bail(rc);
    launch_model_ref(c->stream, m->pos.as<float4>(), m->nrm.as<float4>(), m->tgt.as<float4>(), cloud->n,
                     m->vref.as<float4>());
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);  // `slots` (host) is read by the copy above
    if (ce != cudaSuccess) return cuda_bail(ce, "model upload");
    ModelDev& dv = m->dev;
    dv.cloud = CloudDev{m->pos.as<float4>(), m->nrm.as<float4>(), m->tgt.as<float4>(), cloud->n};
    dv.voxel = m->voxel.as<uint32_t>();
    dv.vcell = m->fused ? m->vcell.as<float4>() : nullptr;
    dv.mref = m->vref.as<float4>();
    dv.ex = d->extents[0];
    dv.ey = d->extents[1];
    dv.ez = d->extents[2];
    dv.exf = (float)dv.ex;
    dv.eyf = (float)dv.ey;
    dv.ezf = (float)dv.ez;
    dv.sx = s[0]; dv.sy = s[1]; dv.sz = s[2];
    dv.tx = tr[0]; dv.ty = tr[1]; dv.tz = tr[2];
    dv.slots = m->slots.as<HashSlot>();
    dv.slot_mask = cap - 1;
    dv.hits = m->hits.as<uint2>();
    dv.n_hits = n_hits;
    dv.fb_min0 = d->feat_min[0];
    dv.fb_max0 = d->feat_max[0];
    dv.dist_steps = (uint32_t)d->distance_step_count;  // float -> uint32 (feature.hpp:41)
    dv.angle_step = d->angle_step;
    dv.resolution = d->resolution;
    dv.diameter = d->diameter;
    // bbox centre / radius for the ICP fixed-point sums
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < cloud->n; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = cloud->pos[(size_t)i * cloud->stride + k];
            if (std::isfinite(v)) {
                lo[k] = std::min(lo[k], v);
                hi[k] = std::max(hi[k], v);
            }
        }
    double dd = 0;
    for (int k = 0; k < 3; ++k) {
        if (!(lo[k] <= hi[k])) lo[k] = hi[k] = 0.f;
        m->centre[k] = 0.5f * (lo[k] + hi[k]);
        dd += 0.25 * (double)(hi[k] - lo[k]) * (double)(hi[k] - lo[k]);
    }
    m->half_diag = (float)std::sqrt(dd);
    *out = m;
    return TM_OK;
}
void tm_model_destroy(tm_model* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    for (DevBuf* b : {&m->pos, &m->nrm, &m->tgt, &m->voxel, &m->vcell, &m->vref, &m->slots, &m->hits,
                      &m->occ[0].bits, &m->occ[1].bits})
        b->release();
    delete m;
}

int tm_voxel_fill(tm_ctx* c, const tm_cloud_view* cloud, const int32_t extents[3],
                  const float to_voxel[16], uint32_t* voxel_out) {
    REQUIRE(c && cloud && extents && to_voxel && voxel_out, "tm_voxel_fill: null argument");
    TRY(bind(c));
    REQUIRE(cloud->n > 0, "cloud is empty");
    float s[3], tr[3];
    TRY(check_to_voxel(to_voxel, s, tr));
    const size_t cells = (size_t)extents[0] * extents[1] * extents[2];
    REQUIRE(cells > 0 && cells < (1ull << 31), "bad extents");
    DevBuf pos, nrm, tgt, vox, blk;
    int rc = upload_cloud(c, cloud, nullptr, 1, pos, nrm, tgt);
    if (!rc) rc = vox.ensure(cells * sizeof(uint32_t));
    // pruned two-pass fill unless TM_VOXEL_FILL_BRUTE=1 (the brute-force kernel is kept as its cross-check)
    // read per call: the parity test fills the same grid both ways inside one process
    const char* brute_env = getenv("TM_VOXEL_FILL_BRUTE");
    const bool brute = brute_env && atoi(brute_env) != 0;
    if (!rc && !brute) rc = blk.ensure(voxel_fill_scratch_bytes(extents[0], extents[1], extents[2]));
    if (!rc) {
        launch_voxel_fill(c->stream, pos.as<float4>(), cloud->n, extents[0], extents[1], extents[2],
                          s[0], s[1], s[2], tr[0], tr[1], tr[2], vox.as<uint32_t>(),
                          brute ? nullptr : blk.as<float>());
        cudaError_t e = cudaMemcpyAsync(voxel_out, vox.p, cells * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(TM_ERR_CUDA, cudaGetErrorString(e));
    }
    pos.release(); nrm.release(); tgt.release(); vox.release(); blk.release();
    return rc;
}

// segment boxes of a resident scene (radius search, k-NN, ICP screening)
static int scene_seg_boxes(tm_ctx* c, tm_scene* s) {
    const uint32_t n_seg = (s->dev.n + BALL_SEG - 1) / BALL_SEG;
    if (!n_seg) return TM_OK;
    TRY(s->seg_lo.ensure((size_t)n_seg * 16));
    TRY(s->seg_hi.ensure((size_t)n_seg * 16));
    launch_seg_bbox(c->stream, s->dev.pos, s->dev.n, s->seg_lo.as<float4>(), s->seg_hi.as<float4>());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    s->dev.seg_lo = s->seg_lo.as<float4>();
    s->dev.seg_hi = s->seg_hi.as<float4>();
    return TM_OK;
}
static int scene_upload_impl(tm_ctx* c, const tm_cloud_view* cloud, const uint8_t* tangent_mask, bool sorted,
                             uint32_t* to_user, tm_scene** out) {
    REQUIRE(c && cloud && out, "tm_scene_upload: null argument");
    TRY(bind(c));
    tm_scene* s = new tm_scene();
    s->ctx = c;
    auto bail = [&](int rc) {
        tm_scene_destroy(s);
        return rc;
    };
    int rc = upload_cloud(c, cloud, tangent_mask, 0, s->pos, s->nrm, s->tgt);
    if (rc) return bail(rc);
    s->dev = CloudDev{s->pos.as<float4>(), s->nrm.as<float4>(), s->tgt.as<float4>(), cloud->n};
    // bounding boxes of the BALL_SEG-point segments: lets the radius search (a8) skip whole
    // segments; tight when the scene is in a space-filling-curve order
    if ((rc = scene_seg_boxes(c, s))) return bail(rc);
    const uint32_t n = cloud->n;
    if (sorted && n > 1) {
        // cloud bounding box from the segment boxes (NaN / empty segments carry inverted boxes)
        const uint32_t n_seg = (n + BALL_SEG - 1) / BALL_SEG;
        std::vector<float4> lo(n_seg), hi(n_seg);
        cudaError_t e = cudaMemcpyAsync(lo.data(), s->seg_lo.p, (size_t)n_seg * 16, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hi.data(), s->seg_hi.p, (size_t)n_seg * 16, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return bail(fail(TM_ERR_CUDA, cudaGetErrorString(e)));
        float blo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (uint32_t g = 0; g < n_seg; ++g) {
            if (!(lo[g].x <= hi[g].x)) continue;
            blo[0] = std::min(blo[0], lo[g].x); bhi[0] = std::max(bhi[0], hi[g].x);
            blo[1] = std::min(blo[1], lo[g].y); bhi[1] = std::max(bhi[1], hi[g].y);
            blo[2] = std::min(blo[2], lo[g].z); bhi[2] = std::max(bhi[2], hi[g].z);
        }
        float inv[3];
        for (int k = 0; k < 3; ++k) {
            const float d = bhi[k] - blo[k];
            inv[k] = d > 0.f && std::isfinite(d) ? 1.f / d : 0.f;
            if (!std::isfinite(blo[k])) blo[k] = 0.f;
        }
        const uint32_t nb = radix_blocks(n);
        DevBuf ka, va, kb, vb, hist, offs, p2, n2, t2;
        auto release = [&] { for (DevBuf* b : {&ka, &va, &kb, &vb, &hist, &offs, &p2, &n2, &t2}) b->release(); };
        rc = ka.ensure((size_t)n * 4);
        if (!rc) rc = va.ensure((size_t)n * 4);
        if (!rc) rc = kb.ensure((size_t)n * 4);
        if (!rc) rc = vb.ensure((size_t)n * 4);
        if (!rc) rc = hist.ensure(((size_t)256 * nb + 1) * 4);
        if (!rc) rc = offs.ensure(((size_t)256 * nb + 1) * 4);
        if (!rc) rc = p2.ensure((size_t)n * 16);
        if (!rc) rc = n2.ensure((size_t)n * 16);
        if (!rc) rc = t2.ensure((size_t)n * 16);
        if (rc) { release(); return bail(rc); }
        launch_morton_codes(c->stream, s->dev.pos, n, blo, inv, ka.as<uint32_t>(), va.as<uint32_t>());
        launch_radix_sort_pairs(c->stream, ka.as<uint32_t>(), va.as<uint32_t>(), kb.as<uint32_t>(), vb.as<uint32_t>(), n,
                                hist.as<uint32_t>(), offs.as<uint32_t>());
        launch_gather_cloud(c->stream, s->dev.pos, s->dev.nrm, s->dev.tgt, va.as<uint32_t>(), n, p2.as<float4>(),
                            n2.as<float4>(), t2.as<float4>());
        e = cudaGetLastError();
        if (e == cudaSuccess && to_user)
            e = cudaMemcpyAsync(to_user, va.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { release(); return bail(fail(TM_ERR_CUDA, cudaGetErrorString(e))); }
        std::swap(s->pos, p2); std::swap(s->nrm, n2); std::swap(s->tgt, t2);  // the sorted copies become the scene
        release();
        s->dev = CloudDev{s->pos.as<float4>(), s->nrm.as<float4>(), s->tgt.as<float4>(), n};
        if ((rc = scene_seg_boxes(c, s))) return bail(rc);
    } else if (to_user) {
        for (uint32_t i = 0; i < n; ++i) to_user[i] = i;
    }
    *out = s;
    return TM_OK;
}
int tm_scene_upload(tm_ctx* c, const tm_cloud_view* cloud, const uint8_t* tangent_mask,
                    tm_scene** out) {
    return scene_upload_impl(c, cloud, tangent_mask, false, nullptr, out);
}
int tm_scene_upload_sorted(tm_ctx* c, const tm_cloud_view* cloud, const uint8_t* tangent_mask,
                           uint32_t* to_user, tm_scene** out) {
    return scene_upload_impl(c, cloud, tangent_mask, true, to_user, out);
}
int tm_scene_set_mask(tm_scene* s, const uint8_t* mask) {
    REQUIRE(s, "null scene");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    const uint8_t* d = nullptr;
    if (mask && s->dev.n) {
        TRY(s->mask_tmp.ensure(s->dev.n));
        CU(cudaMemcpyAsync(s->mask_tmp.p, mask, s->dev.n, cudaMemcpyHostToDevice, c->stream));
        d = s->mask_tmp.as<uint8_t>();
    }
    launch_set_mask(c->stream, s->pos.as<float4>(), s->dev.n, d);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}
void tm_scene_destroy(tm_scene* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    for (DevBuf* b : {&s->pos, &s->nrm, &s->tgt, &s->mask_tmp, &s->seg_lo, &s->seg_hi}) b->release();
    delete s;
}

// ------------------------------------------------------------- stage calls
static void pair_window(const tm_model* m, float min_df, float max_df, float& lower, float& upper) {
    lower = m->dev.diameter * min_df;  // scene.hpp:117-120
    upper = m->dev.diameter * max_df;
    lower *= lower;
    upper *= upper;
}

int tm_features(tm_scene* s, tm_model* m, const uint32_t* pi, const uint32_t* pj, uint64_t n,
                float min_df, float max_df, float* feats, uint32_t* keys, uint8_t* valid) {
    REQUIRE(s && m, "null handle");
    REQUIRE(n == 0 || (pi && pj && keys && valid), "tm_features: null buffer");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n) return TM_OK;
    for (uint64_t i = 0; i < n; ++i)
        REQUIRE(pi[i] < s->dev.n && pj[i] < s->dev.n, "tm_features: scene index out of range");
    DevBuf &di = c->scratch[0], &dj = c->scratch[1], &df = c->scratch[2], &dk = c->scratch[3],
           &dv = c->scratch[4];
    TRY(di.ensure(n * 4)); TRY(dj.ensure(n * 4)); TRY(df.ensure(n * 16)); TRY(dk.ensure(n * 16));
    TRY(dv.ensure(n));
    CU(cudaMemcpyAsync(di.p, pi, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dj.p, pj, n * 4, cudaMemcpyHostToDevice, c->stream));
    float lower, upper;
    pair_window(m, min_df, max_df, lower, upper);
    launch_pair_features_probe(c->stream, s->dev, m->dev, nullptr, di.as<uint32_t>(),
                               dj.as<uint32_t>(), n, lower, upper, 0, df.as<float>(),
                               dk.as<uint4>(), dv.as<uint8_t>(), nullptr, nullptr, nullptr);
    CU(cudaGetLastError());
    if (feats) CU(cudaMemcpyAsync(feats, df.p, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(keys, dk.p, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(valid, dv.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_probe(tm_model* m, const uint32_t* keys, const uint8_t* valid, uint64_t n, uint32_t limit,
             uint64_t* offsets, uint32_t* hits, uint64_t hits_capacity) {
    REQUIRE(m && offsets, "tm_probe: null argument");
    REQUIRE(n == 0 || keys, "tm_probe: null keys");
    tm_ctx* c = m->ctx;
    TRY(bind(c));
    offsets[0] = 0;
    if (!n) return TM_OK;
    DevBuf &dk = c->scratch[0], &dv = c->scratch[1], &hb = c->scratch[2], &hc = c->scratch[3],
           &off = c->scratch[4], &out = c->scratch[5];
    TRY(dk.ensure(n * 16)); TRY(dv.ensure(n)); TRY(hb.ensure(n * 4)); TRY(hc.ensure(n * 4));
    TRY(off.ensure((n + 1) * 8));
    CU(cudaMemcpyAsync(dk.p, keys, n * 16, cudaMemcpyHostToDevice, c->stream));
    if (valid) CU(cudaMemcpyAsync(dv.p, valid, n, cudaMemcpyHostToDevice, c->stream));
    launch_probe(c->stream, m->dev, dk.as<uint4>(), valid ? dv.as<uint8_t>() : nullptr, n, limit,
                 hb.as<uint32_t>(), hc.as<uint32_t>());
    launch_exclusive_scan_u64(c->stream, hc.as<uint32_t>(), off.as<unsigned long long>(), n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(offsets, off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!hits) return TM_OK;
    uint64_t total = offsets[n];
    if (total > hits_capacity) return fail(TM_ERR_CAPACITY, "tm_probe: hits buffer too small");
    if (!total) return TM_OK;
    TRY(out.ensure(total * 8));
    launch_gather_hits(c->stream, m->dev, hb.as<uint32_t>(), off.as<unsigned long long>(), n,
                       out.as<uint2>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hits, out.p, total * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_hypotheses(tm_scene* s, tm_model* m, const uint32_t* pi, const uint32_t* pj, uint64_t n,
                  const uint64_t* offsets, const uint32_t* hits, int force_up, float* T16s,
                  uint8_t* hyp_valid) {
    REQUIRE(s && m, "null handle");
    REQUIRE(n == 0 || (pi && pj && offsets), "tm_hypotheses: null buffer");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n) return TM_OK;
    uint64_t total = offsets[n];
    if (!total) return TM_OK;
    REQUIRE(hits && T16s && hyp_valid, "tm_hypotheses: null output");
    for (uint64_t i = 0; i < 2 * total; ++i)
        REQUIRE(hits[i] < m->dev.cloud.n, "tm_hypotheses: model index out of range");
    for (uint64_t i = 0; i < n; ++i)
        REQUIRE(pi[i] < s->dev.n && pj[i] < s->dev.n, "tm_hypotheses: scene index out of range");
    DevBuf &di = c->scratch[0], &dj = c->scratch[1], &off = c->scratch[2], &dh = c->scratch[3],
           &dT = c->scratch[4], &dv = c->scratch[5], &d16 = c->scratch[6], &sh = c->scratch[7];
    TRY(di.ensure(n * 4)); TRY(dj.ensure(n * 4)); TRY(off.ensure((n + 1) * 8));
    TRY(dh.ensure(total * 8)); TRY(dT.ensure(total * 48)); TRY(dv.ensure(total));
    TRY(d16.ensure(total * 64)); TRY(sh.ensure(24));
    unsigned long long shard[3] = {0ull, total, total};
    CU(cudaMemcpyAsync(di.p, pi, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dj.p, pj, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dh.p, hits, total * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(sh.p, shard, 24, cudaMemcpyHostToDevice, c->stream));
    launch_hypotheses(c->stream, s->dev, m->dev, nullptr, di.as<uint32_t>(), dj.as<uint32_t>(), n,
                      off.as<unsigned long long>(), nullptr, dh.as<uint2>(), force_up,
                      sh.as<unsigned long long>(), dT.as<float4>(), dv.as<uint8_t>(), nullptr);
    launch_colmajor_from_rows(c->stream, dT.as<float4>(), total, d16.as<float>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(T16s, d16.p, total * 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hyp_valid, dv.p, total, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

// device-side ball subsets into (row offsets u64, indices); returns total via host sync
// counts: [centre][segment] u32 (becomes in-row offsets), row_tot: per-centre totals,
// row_off: CSR offsets.  active_ranges (device, n_centres + 1, or null): centre c is searched
// only when active_ranges[c + 1] > active_ranges[c]; skipped centres get an empty row.
static int ball_subsets_dev(tm_ctx* c, const CloudDev& scene, const uint32_t* d_centres,
                            uint32_t n_centres, const uint32_t* active_ranges, float radius, DevBuf& counts,
                            DevBuf& row_tot, DevBuf& row_off, DevBuf* indices, uint64_t* total_out) {
    const uint32_t n_seg = (scene.n + BALL_SEG - 1) / BALL_SEG;
    const size_t nc = (size_t)n_centres * n_seg;
    TRY(counts.ensure(std::max<size_t>(nc, 1) * 4));
    TRY(row_tot.ensure(((size_t)n_centres + 1) * 4));
    TRY(row_off.ensure(((size_t)n_centres + 1) * 8));
    float r2 = radius * radius;
    launch_ball_count(c->stream, scene, d_centres, n_centres, active_ranges, r2, n_seg, counts.as<uint32_t>());
    launch_ball_seg_scan(c->stream, counts.as<uint32_t>(), n_centres, n_seg, row_tot.as<uint32_t>());
    launch_exclusive_scan_u64(c->stream, row_tot.as<uint32_t>(), row_off.as<unsigned long long>(), n_centres);
    CU(cudaGetLastError());
    if (total_out) {
        unsigned long long t = 0;
        CU(cudaMemcpyAsync(&t, row_off.as<unsigned long long>() + n_centres, 8, cudaMemcpyDeviceToHost,
                           c->stream));
        CU(cudaStreamSynchronize(c->stream));
        *total_out = t;
        if (indices) TRY(indices->ensure(std::max<uint64_t>(t, 1) * 4));
    }
    if (indices && indices->p) {
        launch_ball_fill(c->stream, scene, d_centres, n_centres, active_ranges, r2, n_seg, counts.as<uint32_t>(),
                         row_off.as<unsigned long long>(), indices->as<int32_t>());
        CU(cudaGetLastError());
    }
    return TM_OK;
}

int tm_ball_subsets(tm_scene* s, const uint32_t* centres, uint32_t n_centres, float radius,
                    uint64_t* offsets, int32_t* indices, uint64_t capacity) {
    REQUIRE(s && offsets, "tm_ball_subsets: null argument");
    REQUIRE(n_centres == 0 || centres, "tm_ball_subsets: null centres");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    offsets[0] = 0;
    if (!n_centres) return TM_OK;
    for (uint32_t i = 0; i < n_centres; ++i)
        REQUIRE(centres[i] < s->dev.n, "tm_ball_subsets: centre out of range");
    DevBuf &dc = c->scratch[0], &cnt = c->scratch[1], &so = c->scratch[2], &ro = c->scratch[3],
           &idx = c->scratch[4];
    TRY(dc.ensure((size_t)n_centres * 4));
    CU(cudaMemcpyAsync(dc.p, centres, (size_t)n_centres * 4, cudaMemcpyHostToDevice, c->stream));
    uint64_t total = 0;
    TRY(ball_subsets_dev(c, s->dev, dc.as<uint32_t>(), n_centres, nullptr, radius, cnt, so, ro,
                         indices ? &idx : nullptr, &total));
    CU(cudaMemcpyAsync(offsets, ro.p, ((size_t)n_centres + 1) * 8, cudaMemcpyDeviceToHost,
                       c->stream));
    if (indices) {
        if (total > capacity) {
            cudaStreamSynchronize(c->stream);
            return fail(TM_ERR_CAPACITY, "tm_ball_subsets: indices buffer too small");
        }
        if (total)
            CU(cudaMemcpyAsync(indices, idx.p, total * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

// score hypotheses whose rows are resident in d_T; groups = subset rows.
// Full mode: work list + persistent register-tiled kernel.
static int score_full_dev(tm_ctx* c, const CloudDev& scene, const tm_model* m, const float4* d_T,
                          const int32_t* d_sub_idx, const unsigned long long* d_sub_off,
                          const uint32_t* d_g_hyp, uint32_t n_groups, uint32_t items_capacity,
                          DevBuf& n_items_g, DevBuf& item_off, DevBuf& items, DevBuf& ctrl,
                          float thres, float sq_thres, uint32_t* d_counts, unsigned long long* d_scores,
                          bool with_score) {
    // ctrl: [0] work counter (u32) [1] pad, [2..3] n_tests (u64)
    TRY(n_items_g.ensure(std::max<size_t>(n_groups, 1) * 4));
    TRY(item_off.ensure(((size_t)n_groups + 1) * 4));
    TRY(items.ensure(std::max<size_t>(items_capacity, 1) * sizeof(WorkItem)));
    CU(cudaMemsetAsync(ctrl.p, 0, 16, c->stream));
    launch_work_count(c->stream, d_sub_off, d_g_hyp, n_groups, n_items_g.as<uint32_t>(),
                      (unsigned long long*)(ctrl.as<uint32_t>() + 2));
    launch_exclusive_scan_u32(c->stream, n_items_g.as<uint32_t>(), item_off.as<uint32_t>(), n_groups);
    launch_work_fill(c->stream, d_sub_off, d_g_hyp, n_groups, item_off.as<uint32_t>(),
                     items.as<WorkItem>());
    ScoreArgs a;
    a.scene = scene;
    TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &a.model));
    a.sub_idx = d_sub_idx;
    a.items = items.as<WorkItem>();
    a.n_items = item_off.as<uint32_t>() + n_groups;
    a.work_counter = ctrl.as<uint32_t>();
    a.T = d_T;
    a.counts = d_counts;
    a.scores = d_scores;
    a.sq_thres = sq_thres;
    a.stats = nullptr;
    if (!with_score && knobs().scorer >= 8) {
        int& b = c->count_bps[m->fused ? 1 : 0];
        if (!b) b = score_count_x2_max_blocks_per_sm(m->fused);
        const int grid = knobs().score_grid ? knobs().score_grid : c->sm_count * b;
        launch_score_count_x2(c->stream, a, grid, m->fused);
    } else {
        int& b = c->score_bps[m->fused ? 1 : 0][with_score ? 1 : 0];
        if (!b) b = score_full_max_blocks_per_sm(m->fused, with_score);
        const int grid = knobs().score_grid ? knobs().score_grid : c->sm_count * b;
        launch_score_full(c->stream, a, grid, m->fused, with_score);
    }
    CU(cudaGetLastError());
    return TM_OK;
}

int tm_score(tm_scene* s, tm_model* m, const float* T16s, uint64_t n_hyp, const uint32_t* hyp_sub,
             const uint64_t* sub_offsets, const int32_t* sub_indices, uint32_t n_sub,
             float dist_thres, float accept_prob, int early_out, uint32_t* counts, double* scores,
             uint8_t* dropped) {
    REQUIRE(s && m, "null handle");
    REQUIRE(n_hyp == 0 || (T16s && counts), "tm_score: null buffer");
    REQUIRE(n_hyp < (1ull << 31), "tm_score: too many hypotheses for one call");
    REQUIRE(early_out >= 0 && early_out <= 2, "tm_score: early_out must be 0, 1 or 2");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n_hyp) return TM_OK;
    const bool all_scene = hyp_sub == nullptr;
    if (!all_scene) REQUIRE(sub_offsets && (sub_indices || sub_offsets[n_sub] == 0) && n_sub > 0,
                            "tm_score: subset CSR missing");
    // group hypotheses by subset row (stable counting sort on the host)
    const uint32_t n_groups = all_scene ? 1u : n_sub;
    std::vector<uint32_t> g_hyp(n_groups + 1, 0), perm(n_hyp);
    if (all_scene) {
        g_hyp[1] = (uint32_t)n_hyp;
        for (uint64_t h = 0; h < n_hyp; ++h) perm[h] = (uint32_t)h;
    } else {
        for (uint64_t h = 0; h < n_hyp; ++h) {
            REQUIRE(hyp_sub[h] < n_sub, "tm_score: hyp_sub out of range");
            ++g_hyp[hyp_sub[h] + 1];
        }
        for (uint32_t g = 0; g < n_groups; ++g) g_hyp[g + 1] += g_hyp[g];
        std::vector<uint32_t> cur(g_hyp.begin(), g_hyp.end() - 1);
        for (uint64_t h = 0; h < n_hyp; ++h) perm[cur[hyp_sub[h]]++] = (uint32_t)h;
        uint64_t tot = sub_offsets[n_sub];
        for (uint64_t i = 0; i < tot; ++i)
            REQUIRE(sub_indices[i] >= 0 && (uint32_t)sub_indices[i] < s->dev.n,
                    "tm_score: subset index out of range");
    }
    std::vector<float> Tp(16 * n_hyp);
    for (uint64_t l = 0; l < n_hyp; ++l) memcpy(&Tp[16 * l], T16s + 16 * (size_t)perm[l], 64);
    std::vector<unsigned long long> soff(n_groups + 1);
    uint64_t items_cap = 0;
    if (all_scene) {
        soff[0] = 0;
        soff[1] = s->dev.n;
    } else {
        for (uint32_t g = 0; g <= n_groups; ++g) soff[g] = sub_offsets[g];
    }
    for (uint32_t g = 0; g < n_groups; ++g) {
        uint64_t np = soff[g + 1] - soff[g], nh = g_hyp[g + 1] - g_hyp[g];
        items_cap += ((np + SCORE_TILE - 1) / SCORE_TILE) * ((nh + SCORE_HCHUNK - 1) / SCORE_HCHUNK);
    }
    REQUIRE(items_cap < (1ull << 31), "tm_score: too many work items");
    DevBuf &d16 = c->scratch[0], &dT = c->scratch[1], &dso = c->scratch[2], &dsi = c->scratch[3],
           &dgh = c->scratch[4], &dcnt = c->scratch[5], &dsc = c->scratch[6], &w0 = c->scratch[7],
           &w1 = c->scratch[8], &w2 = c->scratch[9], &ctrl = c->scratch[10], &ddrop = c->scratch[11];
    TRY(d16.ensure(n_hyp * 64)); TRY(dT.ensure(n_hyp * 48));
    TRY(dso.ensure((n_groups + 1) * 8)); TRY(dgh.ensure((n_groups + 1) * 4));
    TRY(dcnt.ensure(n_hyp * 4)); TRY(dsc.ensure(n_hyp * 8)); TRY(ctrl.ensure(64));
    CU(cudaMemcpyAsync(d16.p, Tp.data(), n_hyp * 64, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dso.p, soff.data(), (n_groups + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dgh.p, g_hyp.data(), (n_groups + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    const int32_t* d_idx = nullptr;
    if (!all_scene && sub_offsets[n_sub]) {
        TRY(dsi.ensure(sub_offsets[n_sub] * 4));
        CU(cudaMemcpyAsync(dsi.p, sub_indices, sub_offsets[n_sub] * 4, cudaMemcpyHostToDevice,
                           c->stream));
        d_idx = dsi.as<int32_t>();
    }
    launch_rows_from_colmajor(c->stream, d16.as<float>(), n_hyp, dT.as<float4>());
    CU(cudaMemsetAsync(dcnt.p, 0, n_hyp * 4, c->stream));
    CU(cudaMemsetAsync(dsc.p, 0, n_hyp * 8, c->stream));
    const float thres = dist_thres * m->dev.resolution;  // scene.hpp:413
    const float sqt = sq_threshold(thres);
    std::vector<uint8_t> drop_l(n_hyp, 0);
    if (!early_out) {
        TRY(score_full_dev(c, s->dev, m, dT.as<float4>(), d_idx, dso.as<unsigned long long>(),
                           dgh.as<uint32_t>(), n_groups, (uint32_t)items_cap, w0, w1, w2, ctrl, thres, sqt,
                           dcnt.as<uint32_t>(), dsc.as<unsigned long long>(), scores != nullptr));
    } else {
        TRY(w0.ensure(n_hyp * 4)); TRY(ddrop.ensure(n_hyp));
        launch_group_of_hyp(c->stream, dgh.as<uint32_t>(), n_groups, w0.as<uint32_t>());
        // boxes of every 32 subset positions: lets the walker skip steps that cannot reach the grid
        uint64_t max_sub = 0;
        for (uint32_t g = 0; g < n_groups; ++g) max_sub = std::max<uint64_t>(max_sub, soff[g + 1] - soff[g]);
        EarlyArgs a;
        if (early_out == 2) {
            // evenly sampling walk: the rows rewritten in walk order; consecutive positions are far apart,
            // so there are no useful step boxes
            DevBuf& walk = c->scratch[12];
            TRY(walk.ensure(std::max<uint64_t>(soff[n_groups], 1) * 4));
            launch_walk_order_rows(c->stream, d_idx, dso.as<unsigned long long>(), n_groups, (uint32_t)max_sub,
                                   walk.as<int32_t>());
            d_idx = walk.as<int32_t>();
        } else {
            const size_t n_tiles = (size_t)(soff[n_groups] / 32) + n_groups + 2;
            TRY(w1.ensure(n_tiles * 16)); TRY(w2.ensure(n_tiles * 16));
            launch_subset_tile_boxes(c->stream, s->dev, d_idx, dso.as<unsigned long long>(), n_groups,
                                     (uint32_t)max_sub, w1.as<float4>(), w2.as<float4>());
            a.tile_lo = w1.as<float4>();
            a.tile_hi = w2.as<float4>();
        }
        a.scene = s->dev;
        TRY(model_dev_for(c, m, thres, &a.model));
        a.sub_idx = d_idx;
        a.sub_off = dso.as<unsigned long long>();
        a.g_of_hyp = w0.as<uint32_t>();
        a.T = dT.as<float4>();
        a.n_hyp = (uint32_t)n_hyp;
        a.n_hyp_dev = nullptr;
        a.n_tests = nullptr;
        a.sq_thres = sqt;
        a.accept_prob = accept_prob;
        a.early_out = 1;
        a.counts = dcnt.as<uint32_t>();
        a.scores = dsc.as<unsigned long long>();
        a.dropped = ddrop.as<uint8_t>();
        a.tested = nullptr;
        launch_score_early_drop(c->stream, a, m->fused);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(drop_l.data(), ddrop.p, n_hyp, cudaMemcpyDeviceToHost, c->stream));
    }
    std::vector<uint32_t> cnt_l(n_hyp);
    std::vector<unsigned long long> sc_l(n_hyp);
    CU(cudaMemcpyAsync(cnt_l.data(), dcnt.p, n_hyp * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(sc_l.data(), dsc.p, n_hyp * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const double mn = (double)m->dev.cloud.n;
    for (uint64_t l = 0; l < n_hyp; ++l) {
        uint32_t h = perm[l];
        counts[h] = cnt_l[l];
        if (scores) {
            double v = (double)sc_l[l] / SCORE_SCALE;
            scores[h] = drop_l[l] ? v : v / mn;  // un-normalised on drop (scene.hpp:502 vs 509)
        }
        if (dropped) dropped[h] = drop_l[l];
    }
    return TM_OK;
}

uint32_t tm_walk_stride(uint32_t n) { return walk_stride(n); }

int tm_correspondences(tm_scene* s, tm_model* m, const float* T16, float dist_thres,
                       uint32_t* scene_corrs, uint32_t* model_corrs, uint32_t* n_corr,
                       double* score) {
    REQUIRE(s && m && T16 && n_corr, "tm_correspondences: null argument");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    *n_corr = 0;
    if (score) *score = 0.0;
    if (!s->dev.n) return TM_OK;
    Rows R;
    R.r0 = make_float4(T16[0], T16[4], T16[8], T16[12]);
    R.r1 = make_float4(T16[1], T16[5], T16[9], T16[13]);
    R.r2 = make_float4(T16[2], T16[6], T16[10], T16[14]);
    const uint32_t n_seg = (s->dev.n + CORR_SEG - 1) / CORR_SEG;
    DevBuf &cnt = c->scratch[0], &off = c->scratch[1], &sc = c->scratch[2], &mc = c->scratch[3],
           &acc = c->scratch[4];
    TRY(cnt.ensure(n_seg * 4)); TRY(off.ensure((n_seg + 1) * 4)); TRY(acc.ensure(8));
    TRY(sc.ensure((size_t)s->dev.n * 4)); TRY(mc.ensure((size_t)s->dev.n * 4));
    const float sqt = sq_threshold(dist_thres * m->dev.resolution);
    CU(cudaMemsetAsync(acc.p, 0, 8, c->stream));
    launch_corr_count(c->stream, s->dev, m->dev, R, sqt, n_seg, cnt.as<uint32_t>(),
                      acc.as<unsigned long long>(), m->fused);
    launch_exclusive_scan_u32(c->stream, cnt.as<uint32_t>(), off.as<uint32_t>(), n_seg);
    launch_corr_fill(c->stream, s->dev, m->dev, R, sqt, n_seg, off.as<uint32_t>(),
                     sc.as<uint32_t>(), mc.as<uint32_t>(), m->fused);
    CU(cudaGetLastError());
    uint32_t total = 0;
    unsigned long long fx = 0;
    CU(cudaMemcpyAsync(&total, off.as<uint32_t>() + n_seg, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&fx, acc.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (total && scene_corrs)
        CU(cudaMemcpyAsync(scene_corrs, sc.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
    if (total && model_corrs)
        CU(cudaMemcpyAsync(model_corrs, mc.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *n_corr = total;
    if (score) *score = (double)fx / SCORE_SCALE / (double)m->dev.cloud.n;
    return TM_OK;
}

// --------------------------------------------------------------------- ICP
struct IcpBufs {
    DevBuf Tcur, Tbest, sums_cur, sums_best, iters, active;
    void release() {
        for (DevBuf* b : {&Tcur, &Tbest, &sums_cur, &sums_best, &iters, &active}) b->release();
    }
    int ensure(uint32_t k) {
        size_t kk = std::max(k, 1u);
        TRY(Tcur.ensure(kk * 48)); TRY(Tbest.ensure(kk * 48));
        TRY(sums_cur.ensure(kk * ICP_NSUM * 8)); TRY(sums_best.ensure(kk * ICP_NSUM * 8));
        TRY(iters.ensure(kk * 4)); TRY(active.ensure(kk * 4));
        return TM_OK;
    }
    IcpState state() {
        return IcpState{Tcur.as<float4>(), Tbest.as<float4>(), sums_cur.as<long long>(),
                        sums_best.as<long long>(), iters.as<uint32_t>(), active.as<uint32_t>()};
    }
};
static double icp_fix_scale(const tm_model* m, uint32_t n_scene, float thres) {
    // |s'|,|m'| <= r = half bbox diagonal + thres; n * r^2 * 2^bits < 2^62
    double r = (double)m->half_diag + (double)thres + 1e-6;
    double bound = std::max(1.0, (double)std::max(n_scene, 1u) * std::max(r * r, r));
    int bits = (int)std::floor(62.0 - std::log2(bound));
    bits = std::max(8, std::min(40, bits));
    return std::ldexp(1.0, bits);
}
struct tm_comm;
static int comm_allreduce_sum_i64(tm_comm* cm, void* buf, size_t count, cudaStream_t st);  // NCCL section
// how the scene points of one ICP pass are split: this process accumulates [pt_begin, pt_end)
// (as `emulate` consecutive sub-ranges when emulate > 1) and, with a communicator, the 64-bit
// fixed-point sums are all-reduced — integer sums, so any split gives the same bits.
struct IcpSplit {
    uint32_t pt_begin = 0, pt_end = 0;
    uint64_t n_total = 0;  // scene points over all ranks (fixes the fixed-point scale)
    tm_comm* comm = nullptr;
    uint32_t emulate = 1;
};
// enqueue the ICP loop for k transforms already in b.Tcur with b.active set
static int icp_enqueue(tm_ctx* c, const CloudDev& scene, const tm_model* m, IcpBufs& b, uint32_t k,
                       uint32_t max_iterations, float dist_thres, const IcpSplit* split = nullptr) {
    const float thres = (2 * dist_thres) * m->dev.resolution;  // scene.hpp:373 + :413
    const float sqt = sq_threshold(thres);
    IcpSplit sp;
    if (split) sp = *split;
    else { sp.pt_end = scene.n; sp.n_total = scene.n; }
    const double fs = icp_fix_scale(m, (uint32_t)std::min<uint64_t>(sp.n_total, 0xffffffffull), thres);
    ModelDev mdev;
    TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &mdev));
    CU(cudaMemsetAsync(b.sums_cur.p, 0, (size_t)k * ICP_NSUM * 8, c->stream));
    CU(cudaMemsetAsync(b.sums_best.p, 0, (size_t)k * ICP_NSUM * 8, c->stream));
    CU(cudaMemsetAsync(b.iters.p, 0, (size_t)k * 4, c->stream));
    const int grid = c->sm_count * 4;
    IcpState st = b.state();
    const uint32_t parts = std::max(1u, sp.emulate);
    const uint64_t span = sp.pt_end - sp.pt_begin;
    for (uint32_t it = 0; it <= max_iterations; ++it) {
        for (uint32_t w = 0; w < parts; ++w) {
            const uint32_t b0 = sp.pt_begin + (uint32_t)(span * w / parts);
            const uint32_t b1 = sp.pt_begin + (uint32_t)(span * (w + 1) / parts);
            if (b1 > b0)
                launch_icp_accumulate(c->stream, scene, mdev, st.Tcur, st.active, k, b0, b1, sqt,
                                      m->centre[0], m->centre[1], m->centre[2], fs, st.sums_cur, grid,
                                      m->fused);
        }
        if (sp.comm) TRY(comm_allreduce_sum_i64(sp.comm, st.sums_cur, (size_t)k * ICP_NSUM, c->stream));
        launch_icp_step(c->stream, st, k, it == 0 ? 1 : 0, max_iterations, 1.0 / fs, m->centre[0],
                        m->centre[1], m->centre[2]);
    }
    CU(cudaGetLastError());
    return TM_OK;
}

static int icp_run(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations,
                   float dist_thres, float* T16s_out, uint32_t* counts, double* scores, uint32_t* iters,
                   const IcpSplit* split) {
    REQUIRE(s && m, "null handle");
    REQUIRE(n == 0 || (T16s && T16s_out && counts), "tm_icp: null buffer");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n) return TM_OK;
    if (max_iterations == 0 && !split) {  // scene.hpp:371: the match is returned unchanged
        memcpy(T16s_out, T16s, (size_t)n * 64);
        if (iters) memset(iters, 0, (size_t)n * 4);
        return tm_score(s, m, T16s, n, nullptr, nullptr, nullptr, 0, dist_thres, 0.f, 0, counts,
                        scores, nullptr);
    }
    IcpBufs b;
    int rc = b.ensure(n);
    DevBuf d16;
    if (!rc) rc = d16.ensure((size_t)n * 64);
    auto done = [&](int code) {
        b.release();
        d16.release();
        return code;
    };
    if (rc) return done(rc);
    std::vector<uint32_t> ones(n, 1u);
    cudaError_t e = cudaMemcpyAsync(d16.p, T16s, (size_t)n * 64, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(b.active.p, ones.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return done(fail(TM_ERR_CUDA, cudaGetErrorString(e)));
    launch_rows_from_colmajor(c->stream, d16.as<float>(), n, b.Tcur.as<float4>());
    if ((rc = icp_enqueue(c, s->dev, m, b, n, max_iterations, dist_thres, split))) return done(rc);
    launch_colmajor_from_rows(c->stream, b.Tbest.as<float4>(), n, d16.as<float>());
    std::vector<long long> sums((size_t)n * ICP_NSUM);
    e = cudaMemcpyAsync(T16s_out, d16.p, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(sums.data(), b.sums_best.p, sums.size() * 8, cudaMemcpyDeviceToHost,
                            c->stream);
    if (e == cudaSuccess && iters)
        e = cudaMemcpyAsync(iters, b.iters.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return done(fail(TM_ERR_CUDA, cudaGetErrorString(e)));
    for (uint32_t h = 0; h < n; ++h) {
        counts[h] = (uint32_t)sums[(size_t)h * ICP_NSUM];
        if (scores)
            scores[h] = (double)sums[(size_t)h * ICP_NSUM + 16] / SCORE_SCALE / (double)m->dev.cloud.n;
    }
    return done(TM_OK);
}

int tm_icp(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations,
           float dist_thres, float* T16s_out, uint32_t* counts, double* scores, uint32_t* iters) {
    return icp_run(s, m, T16s, n, max_iterations, dist_thres, T16s_out, counts, scores, iters, nullptr);
}

int tm_traits_project(tm_ctx* c, int kind, const float g2l[16], float radius, float threshold,
                      const float* xyz, uint64_t n, float* uvw, uint8_t* ok) {
    REQUIRE(c && g2l, "tm_traits_project: null argument");
    REQUIRE(kind >= 0 && kind <= 3, "tm_traits_project: unknown kind");
    REQUIRE(n == 0 || (xyz && uvw && ok), "tm_traits_project: null buffer");
    TRY(bind(c));
    if (!n) return TM_OK;
    DevBuf &in = c->scratch[0], &out = c->scratch[1], &dok = c->scratch[2];
    TRY(in.ensure(n * 12)); TRY(out.ensure(n * 12)); TRY(dok.ensure(n));
    CU(cudaMemcpyAsync(in.p, xyz, n * 12, cudaMemcpyHostToDevice, c->stream));
    float4 r0 = make_float4(g2l[0], g2l[4], g2l[8], g2l[12]);
    float4 r1 = make_float4(g2l[1], g2l[5], g2l[9], g2l[13]);
    float4 r2 = make_float4(g2l[2], g2l[6], g2l[10], g2l[14]);
    launch_traits_project(c->stream, kind, r0, r1, r2, radius, threshold, in.as<float>(), n,
                          out.as<float>(), dok.as<uint8_t>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(uvw, out.p, n * 12, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ok, dok.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_ctx_measure_l2_gather(tm_ctx* c, uint64_t working_set_bytes, double* gb_per_s) {
    REQUIRE(c && gb_per_s, "tm_ctx_measure_l2_gather: null argument");
    REQUIRE(working_set_bytes >= (1u << 20), "tm_ctx_measure_l2_gather: working set too small");
    TRY(bind(c));
    uint64_t cells = 1;
    while (cells * 2 * 16 <= working_set_bytes && cells < (1ull << 31)) cells *= 2;
    DevBuf &buf = c->scratch[0], &out = c->scratch[1];
    TRY(buf.ensure(cells * 16)); TRY(out.ensure(16));
    CU(cudaMemsetAsync(buf.p, 0, cells * 16, c->stream));
    const int grid = c->sm_count * 8;
    const uint32_t iters = 256;
    launch_l2_gather(c->stream, buf.as<float4>(), (uint32_t)(cells - 1), iters, out.as<float>(), grid);  // warm-up: fills L2
    CU(cudaGetLastError());
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, c->stream));
    for (int r = 0; r < 3; ++r)
        launch_l2_gather(c->stream, buf.as<float4>(), (uint32_t)(cells - 1), iters, out.as<float>(), grid);
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double loads = 3.0 * (double)grid * 256.0 * iters * 8.0;
    *gb_per_s = ms > 0.f ? loads * 32.0 / (ms * 1e-3) / 1e9 : 0.0;
    return TM_OK;
}

// ------------------------------------------------------------ model::init pair enumeration
int tm_model_pair_bounds(tm_ctx* c, const float* pos3, const float* tgt3, uint32_t T, float lower, float upper,
                         float feat_min[3], float feat_max[3], uint64_t* n_pass) {
    REQUIRE(c && feat_min && feat_max, "tm_model_pair_bounds: null argument");
    REQUIRE(T == 0 || (pos3 && tgt3), "tm_model_pair_bounds: null buffer");
    TRY(bind(c));
    for (int k = 0; k < 3; ++k) {
        feat_min[k] = std::numeric_limits<float>::max();
        feat_max[k] = std::numeric_limits<float>::lowest();
    }
    if (n_pass) *n_pass = 0;
    if (!T) return TM_OK;
    DevBuf &dp = c->scratch[0], &dt = c->scratch[1], &db = c->scratch[2];
    TRY(dp.ensure((size_t)T * 12)); TRY(dt.ensure((size_t)T * 12)); TRY(db.ensure(64));
    CU(cudaMemcpyAsync(dp.p, pos3, (size_t)T * 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dt.p, tgt3, (size_t)T * 12, cudaMemcpyHostToDevice, c->stream));
    const uint32_t init[8] = {0x7f7fffffu, 0x7f7fffffu, 0x7f7fffffu, 0u, 0u, 0u, 0u, 0u};  // min: FLT_MAX bits, max: +0
    CU(cudaMemcpyAsync(db.p, init, 32, cudaMemcpyHostToDevice, c->stream));
    launch_model_pair_bounds(c->stream, dp.as<float>(), dt.as<float>(), T, lower, upper, db.as<uint32_t>(),
                             reinterpret_cast<unsigned long long*>(db.as<uint32_t>() + 6), c->sm_count * 8);
    CU(cudaGetLastError());
    uint32_t out[8];
    CU(cudaMemcpyAsync(out, db.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long cnt;
    memcpy(&cnt, &out[6], 8);
    if (n_pass) *n_pass = cnt;
    if (cnt) {
        memcpy(feat_min, &out[0], 12);
        memcpy(feat_max, &out[3], 12);
    }
    return TM_OK;
}
int tm_model_pair_keys(tm_ctx* c, const float* pos3, const float* tgt3, uint32_t T, float lower, float upper,
                       float fmn0, float fmx0, uint32_t steps, float angle_step, uint64_t* keys) {
    REQUIRE(c, "tm_model_pair_keys: null context");
    REQUIRE(T == 0 || (pos3 && tgt3 && keys), "tm_model_pair_keys: null buffer");
    REQUIRE((uint64_t)T * T <= (1ull << 28), "tm_model_pair_keys: too many pairs for one call");
    REQUIRE(steps < (1u << 24), "tm_model_pair_keys: distance_step_count too large for the packed key");
    TRY(bind(c));
    if (!T) return TM_OK;
    const size_t n = (size_t)T * T;
    DevBuf &dp = c->scratch[0], &dt = c->scratch[1], &dk = c->scratch[2];
    TRY(dp.ensure((size_t)T * 12)); TRY(dt.ensure((size_t)T * 12)); TRY(dk.ensure(n * 8));
    CU(cudaMemcpyAsync(dp.p, pos3, (size_t)T * 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dt.p, tgt3, (size_t)T * 12, cudaMemcpyHostToDevice, c->stream));
    launch_model_pair_keys(c->stream, dp.as<float>(), dt.as<float>(), T, lower, upper, fmn0, fmx0, steps, angle_step,
                           dk.as<unsigned long long>(), c->sm_count * 8);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(keys, dk.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

// ------------------------------------------------------------ pre-processing (k-NN, curvature)
static int knn_dev(tm_ctx* c, tm_scene* s, const uint32_t* d_query, uint32_t n_query, uint32_t k, DevBuf& idx,
                   DevBuf* d2) {
    TRY(idx.ensure((size_t)std::max(n_query, 1u) * k * 4));
    if (d2) TRY(d2->ensure((size_t)std::max(n_query, 1u) * k * 4));
    launch_knn(c->stream, s->dev, d_query, n_query, k, idx.as<int32_t>(), d2 ? d2->as<float>() : nullptr);
    CU(cudaGetLastError());
    return TM_OK;
}
int tm_scene_knn(tm_scene* s, const uint32_t* query_idx, uint32_t n_query, uint32_t k, int32_t* out_idx,
                 float* out_d2) {
    REQUIRE(s, "null scene");
    REQUIRE(k >= 1 && k <= 32, "tm_scene_knn: k must be in [1, 32]");
    REQUIRE(n_query == 0 || (query_idx && out_idx), "tm_scene_knn: null buffer");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n_query) return TM_OK;
    for (uint32_t i = 0; i < n_query; ++i) REQUIRE(query_idx[i] < s->dev.n, "tm_scene_knn: query index out of range");
    DevBuf &dq = c->scratch[0], &di = c->scratch[1], &dd = c->scratch[2];
    TRY(dq.ensure((size_t)n_query * 4));
    CU(cudaMemcpyAsync(dq.p, query_idx, (size_t)n_query * 4, cudaMemcpyHostToDevice, c->stream));
    TRY(knn_dev(c, s, dq.as<uint32_t>(), n_query, k, di, out_d2 ? &dd : nullptr));
    CU(cudaMemcpyAsync(out_idx, di.p, (size_t)n_query * k * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_d2) CU(cudaMemcpyAsync(out_d2, dd.p, (size_t)n_query * k * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}
int tm_scene_curvature(tm_scene* s, const uint32_t* query_idx, uint32_t n_query, uint32_t k, float* pc_min,
                       float* pc_max, float* cov9) {
    REQUIRE(s, "null scene");
    REQUIRE(k >= 1 && k <= 32, "tm_scene_curvature: k must be in [1, 32]");
    REQUIRE(n_query == 0 || (query_idx && pc_min && pc_max), "tm_scene_curvature: null buffer");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n_query) return TM_OK;
    for (uint32_t i = 0; i < n_query; ++i)
        REQUIRE(query_idx[i] < s->dev.n, "tm_scene_curvature: query index out of range");
    DevBuf &dq = c->scratch[0], &di = c->scratch[1], &dmn = c->scratch[2], &dmx = c->scratch[3], &dcov = c->scratch[4];
    TRY(dq.ensure((size_t)n_query * 4)); TRY(dmn.ensure((size_t)n_query * 4)); TRY(dmx.ensure((size_t)n_query * 4));
    if (cov9) TRY(dcov.ensure((size_t)n_query * 36));
    CU(cudaMemcpyAsync(dq.p, query_idx, (size_t)n_query * 4, cudaMemcpyHostToDevice, c->stream));
    TRY(knn_dev(c, s, dq.as<uint32_t>(), n_query, k, di, nullptr));
    launch_curvature(c->stream, s->dev, dq.as<uint32_t>(), n_query, k, di.as<int32_t>(), dmn.as<float>(),
                     dmx.as<float>(), cov9 ? dcov.as<float>() : nullptr);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pc_min, dmn.p, (size_t)n_query * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(pc_max, dmx.p, (size_t)n_query * 4, cudaMemcpyDeviceToHost, c->stream));
    if (cov9) CU(cudaMemcpyAsync(cov9, dcov.p, (size_t)n_query * 36, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}
int tm_scene_tangent_mask(tm_scene* s, uint32_t k, float ratio, uint8_t* mask_out, int apply,
                          uint32_t* n_tangent) {
    REQUIRE(s, "null scene");
    REQUIRE(k >= 1 && k <= 32, "tm_scene_tangent_mask: k must be in [1, 32]");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    const uint32_t n = s->dev.n;
    if (n_tangent) *n_tangent = 0;
    if (!n) return TM_OK;
    DevBuf &fl = c->scratch[0], &off = c->scratch[1], &cand = c->scratch[2], &di = c->scratch[3],
           &dmn = c->scratch[4], &dmx = c->scratch[5], &dmask = c->scratch[6];
    TRY(fl.ensure((size_t)n * 4)); TRY(off.ensure(((size_t)n + 1) * 4)); TRY(dmask.ensure(n));
    launch_tangent_candidates(c->stream, s->dev.tgt, n, fl.as<uint32_t>());
    launch_exclusive_scan_u32(c->stream, fl.as<uint32_t>(), off.as<uint32_t>(), n);
    uint32_t n_cand = 0;
    CU(cudaMemcpyAsync(&n_cand, off.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemsetAsync(dmask.p, 0, n, c->stream));
    TRY(cand.ensure((size_t)std::max(n_cand, 1u) * 4));
    TRY(dmn.ensure((size_t)std::max(n_cand, 1u) * 4)); TRY(dmx.ensure((size_t)std::max(n_cand, 1u) * 4));
    if (n_cand) {
        launch_compact(c->stream, fl.as<uint32_t>(), off.as<uint32_t>(), n, cand.as<uint32_t>());
        TRY(knn_dev(c, s, cand.as<uint32_t>(), n_cand, k, di, nullptr));
        launch_curvature(c->stream, s->dev, cand.as<uint32_t>(), n_cand, k, di.as<int32_t>(), dmn.as<float>(),
                         dmx.as<float>(), nullptr);
    }
    launch_tangent_mask(c->stream, s->pos.as<float4>(), n, cand.as<uint32_t>(), n_cand, dmn.as<float>(),
                        dmx.as<float>(), ratio, dmask.as<uint8_t>(), apply);
    CU(cudaGetLastError());
    std::vector<uint8_t> tmp;
    uint8_t* dst = mask_out;
    if (!dst && n_tangent) {
        tmp.resize(n);
        dst = tmp.data();
    }
    if (dst) CU(cudaMemcpyAsync(dst, dmask.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n_tangent && dst) {
        uint32_t cnt = 0;
        for (uint32_t i = 0; i < n; ++i) cnt += dst[i];
        *n_tangent = cnt;
    }
    return TM_OK;
}

// ------------------------------------------------------------ opencl/icp.cl path (a15)
int tm_uvicp_projection(tm_ctx* c, int projector, const float* pnts4, int32_t n, const float* image4,
                        const int32_t img_size[2], const int32_t img_margin[2], const float mat_align[16],
                        const float mat_uvw[16], const float mat_proj[16], const float mat_norm[16],
                        float max_corr_dist, float* out_positions4, int32_t* model_indices,
                        int32_t* scene_indices, uint32_t* n_corr) {
    REQUIRE(c && img_size && img_margin && mat_align && mat_uvw && mat_proj && mat_norm,
            "tm_uvicp_projection: null argument");
    REQUIRE(projector == 0 || projector == 1, "tm_uvicp_projection: unknown projector");
    REQUIRE(n >= 0 && img_size[0] > 0 && img_size[1] > 0, "tm_uvicp_projection: bad sizes");
    REQUIRE(n == 0 || (pnts4 && image4 && out_positions4 && model_indices && scene_indices),
            "tm_uvicp_projection: null buffer");
    TRY(bind(c));
    if (n_corr) *n_corr = 0;
    if (!n) return TM_OK;
    const size_t n_img = (size_t)img_size[0] * (size_t)img_size[1];
    DevBuf &dp = c->scratch[0], &di = c->scratch[1], &dop = c->scratch[2], &dmi = c->scratch[3],
           &dsi = c->scratch[4], &dn = c->scratch[5];
    TRY(dp.ensure((size_t)n * 16)); TRY(di.ensure(n_img * 16)); TRY(dop.ensure((size_t)n * 16));
    TRY(dmi.ensure((size_t)n * 4)); TRY(dsi.ensure((size_t)n * 4)); TRY(dn.ensure(4));
    CU(cudaMemcpyAsync(dp.p, pnts4, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(di.p, image4, n_img * 16, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(dn.p, 0, 4, c->stream));
    launch_uvicp_projection(c->stream, projector, dp.as<float4>(), n, di.as<float4>(), img_size, img_margin,
                            mat_align, mat_uvw, mat_proj, mat_norm, max_corr_dist, dop.as<float4>(),
                            dmi.as<int>(), dsi.as<int>(), dn.as<unsigned int>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_positions4, dop.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(model_indices, dmi.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(scene_indices, dsi.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    uint32_t nc = 0;
    CU(cudaMemcpyAsync(&nc, dn.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n_corr) *n_corr = nc;
    return TM_OK;
}

int tm_uvicp_correlation(tm_ctx* c, const float* scene4, uint32_t n_scene, const float* model4,
                         uint32_t n_model, const int32_t* indices_scene, const int32_t* indices_model, int32_t n,
                         const float centroid_scene[4], const float centroid_model[4], float* records16,
                         double cov9[9]) {
    REQUIRE(c && centroid_scene && centroid_model && cov9, "tm_uvicp_correlation: null argument");
    REQUIRE(n >= 0, "tm_uvicp_correlation: negative n");
    REQUIRE(n == 0 || (scene4 && model4 && indices_scene && indices_model), "tm_uvicp_correlation: null buffer");
    for (int32_t k = 0; k < n; ++k)
        REQUIRE(indices_scene[k] >= 0 && (uint32_t)indices_scene[k] < n_scene && indices_model[k] >= 0 &&
                    (uint32_t)indices_model[k] < n_model,
                "tm_uvicp_correlation: index out of range");
    TRY(bind(c));
    DevBuf &ds = c->scratch[0], &dm = c->scratch[1], &dis = c->scratch[2], &dim = c->scratch[3],
           &drec = c->scratch[4], &dpart = c->scratch[5], &dcov = c->scratch[6];
    const int blocks = uvicp_correlation_blocks(n);
    TRY(ds.ensure((size_t)n_scene * 16 + 16)); TRY(dm.ensure((size_t)n_model * 16 + 16));
    TRY(dis.ensure((size_t)n * 4 + 4)); TRY(dim.ensure((size_t)n * 4 + 4));
    TRY(dpart.ensure((size_t)blocks * 72 + 72)); TRY(dcov.ensure(72));
    if (records16) TRY(drec.ensure((size_t)n * 64 + 64));
    if (n) {
        CU(cudaMemcpyAsync(ds.p, scene4, (size_t)n_scene * 16, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(dm.p, model4, (size_t)n_model * 16, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(dis.p, indices_scene, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(dim.p, indices_model, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    }
    launch_uvicp_correlation(c->stream, ds.as<float4>(), dm.as<float4>(), dis.as<int>(), dim.as<int>(), n,
                             centroid_scene, centroid_model, records16 ? drec.as<float>() : nullptr,
                             dpart.as<double>(), dcov.as<double>());
    CU(cudaGetLastError());
    if (records16 && n) CU(cudaMemcpyAsync(records16, drec.p, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(cov9, dcov.p, 72, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

// ------------------------------------------------------------ resident query
struct QueryOut {  // one contiguous device block, read back in one copy
    unsigned long long shard[3];  // h_begin, h_end, H
    unsigned long long best;
    unsigned long long n_tests;
    unsigned long long n_valid;
    double best_score;
    float best_T16[16];
    uint32_t n_local;
    uint32_t err;
    uint32_t work_counter;
    uint32_t pad;
    unsigned long long best_acc[2];  // lazy score of the selected pose: fixed-point sum, inlier count
};

struct tm_query {
    tm_scene* s;
    tm_model* m;
    tm_query_params p;
    uint32_t rank = 0, world = 1;
    uint32_t n_outer = 0;
    uint64_t n_pairs = 0;
    uint64_t cap_hyp = 0;
    uint32_t items_cap = 0;
    uint64_t sub_total = 0;
    DevBuf outer, pair_outer, pair_j, outer_pair_off;
    DevBuf ball_counts, ball_seg_off, sub_off, sub_idx, sub_idx_walk;
    DevBuf valid, hit_begin, hit_count, hyp_off;
    DevBuf g_hyp, g_of_hyp, T, hyp_valid, hyp_pair, counts, scores, dropped;
    DevBuf n_items_g, item_off, items, ctrl;
    DevBuf out;  // QueryOut
    DevBuf topk_ids, topk_keys, icp_T16, stats, tile_lo, tile_hi;
    uint32_t max_sub = 0;
    IcpBufs icp;
    QueryOut host_out;
    bool ran = false;
    bool lazy = false;          // last run used the count-only scorer: scores[] is filled on demand
    bool scores_valid = false;  // scores[] holds every hypothesis' score
    float run_thres = 0.f, run_sqt = 0.f;
    cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;  // around the scoring kernel
};

static int count_valid_pairs_dev(tm_query* q);

int tm_query_create(tm_scene* s, tm_model* m, const tm_query_params* p, tm_query** out) {
    REQUIRE(s && m && p && out, "tm_query_create: null argument");
    REQUIRE(s->ctx == m->ctx, "scene and model live in different contexts");
    REQUIRE(p->icp_top_k <= 4096, "icp_top_k too large");
    REQUIRE(p->early_out >= 0 && p->early_out <= 2, "early_out must be 0, 1 or 2");
    tm_query* q = new tm_query();
    q->s = s;
    q->m = m;
    q->p = *p;
    memset(&q->host_out, 0, sizeof(QueryOut));
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e == cudaSuccess) e = cudaEventCreate(&q->ev_s0);
    if (e == cudaSuccess) e = cudaEventCreate(&q->ev_s1);
    if (e != cudaSuccess) {
        tm_query_destroy(q);
        return fail(TM_ERR_CUDA, std::string("tm_query_create: ") + cudaGetErrorString(e));
    }
    *out = q;
    return TM_OK;
}
void tm_query_destroy(tm_query* q) {
    if (!q) return;
    cudaSetDevice(q->s->ctx->device);
    if (q->ev_s0) cudaEventDestroy(q->ev_s0);
    if (q->ev_s1) cudaEventDestroy(q->ev_s1);
    for (DevBuf* b :
         {&q->outer, &q->pair_outer, &q->pair_j, &q->outer_pair_off, &q->ball_counts,
          &q->ball_seg_off, &q->sub_off, &q->sub_idx, &q->sub_idx_walk, &q->valid, &q->hit_begin, &q->hit_count,
          &q->hyp_off, &q->g_hyp, &q->g_of_hyp, &q->T, &q->hyp_valid, &q->hyp_pair, &q->counts,
          &q->scores, &q->dropped, &q->n_items_g, &q->item_off, &q->items, &q->ctrl, &q->out,
          &q->topk_ids, &q->topk_keys, &q->icp_T16, &q->stats, &q->tile_lo, &q->tile_hi})
        b->release();
    q->icp.release();
    delete q;
}

int tm_query_set_shard(tm_query* q, uint32_t rank, uint32_t world) {
    REQUIRE(q && world > 0 && rank < world, "tm_query_set_shard: bad rank/world");
    q->rank = rank;
    q->world = world;
    return TM_OK;
}

int tm_query_set_pairs(tm_query* q, const uint32_t* outer, uint32_t n_outer,
                       const uint32_t* pair_outer, const uint32_t* pair_j, uint64_t n_pairs) {
    REQUIRE(q, "null query");
    REQUIRE(n_outer == 0 || outer, "null outer");
    REQUIRE(n_pairs == 0 || (pair_outer && pair_j), "null pairs");
    REQUIRE(n_pairs < (1ull << 31), "too many pairs");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    const uint32_t ns = q->s->dev.n;
    for (uint32_t o = 0; o < n_outer; ++o) REQUIRE(outer[o] < ns, "outer index out of range");
    std::vector<uint32_t> opo(n_outer + 1, 0);
    for (uint64_t k = 0; k < n_pairs; ++k) {
        REQUIRE(pair_outer[k] < n_outer && pair_j[k] < ns, "pair index out of range");
        REQUIRE(k == 0 || pair_outer[k] >= pair_outer[k - 1], "pairs must be sorted by outer");
        ++opo[pair_outer[k] + 1];
    }
    for (uint32_t o = 0; o < n_outer; ++o) opo[o + 1] += opo[o];
    q->n_outer = n_outer;
    q->n_pairs = n_pairs;
    TRY(q->outer.ensure(std::max(n_outer, 1u) * 4ull));
    TRY(q->pair_outer.ensure(std::max<uint64_t>(n_pairs, 1) * 4));
    TRY(q->pair_j.ensure(std::max<uint64_t>(n_pairs, 1) * 4));
    TRY(q->outer_pair_off.ensure((n_outer + 1) * 4ull));
    if (n_outer) CU(cudaMemcpyAsync(q->outer.p, outer, n_outer * 4ull, cudaMemcpyHostToDevice, c->stream));
    if (n_pairs) {
        CU(cudaMemcpyAsync(q->pair_outer.p, pair_outer, n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(q->pair_j.p, pair_j, n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaMemcpyAsync(q->outer_pair_off.p, opo.data(), (n_outer + 1) * 4ull, cudaMemcpyHostToDevice,
                       c->stream));
    // capacities: hypotheses, subset indices (one sizing pass), work items
    uint64_t limit = q->p.query_limit ? q->p.query_limit : 200;
    uint64_t cap = q->p.max_hypotheses ? q->p.max_hypotheses
                                       : std::min<uint64_t>(n_pairs * limit, 1ull << 24);
    if (q->p.hyp_limit) cap = std::min<uint64_t>(cap, q->p.hyp_limit);
    cap = std::max<uint64_t>(cap, 1);
    q->cap_hyp = cap;
    TRY(q->T.ensure(cap * 48)); TRY(q->hyp_valid.ensure(cap)); TRY(q->hyp_pair.ensure(cap * 4));
    TRY(q->counts.ensure(cap * 4)); TRY(q->scores.ensure(cap * 8)); TRY(q->dropped.ensure(cap));
    TRY(q->g_of_hyp.ensure(cap * 4));
    TRY(q->valid.ensure(std::max<uint64_t>(n_pairs, 1))); TRY(q->hit_begin.ensure(std::max<uint64_t>(n_pairs, 1) * 4));
    TRY(q->hit_count.ensure(std::max<uint64_t>(n_pairs, 1) * 4)); TRY(q->hyp_off.ensure((n_pairs + 1) * 8));
    TRY(q->g_hyp.ensure((n_outer + 1) * 4ull));
    TRY(q->out.ensure(sizeof(QueryOut))); TRY(q->ctrl.ensure(64));
    uint64_t total = 0;
    std::vector<unsigned long long> so(n_outer + 1, 0);
    if (n_outer) {
        // sizing pass over ALL outer samples (a rank's shard is only known per run)
        TRY(ball_subsets_dev(c, q->s->dev, q->outer.as<uint32_t>(), n_outer, nullptr, q->m->dev.diameter,
                             q->ball_counts, q->ball_seg_off, q->sub_off, &q->sub_idx, &total));
        CU(cudaMemcpyAsync(so.data(), q->sub_off.p, (n_outer + 1) * 8ull, cudaMemcpyDeviceToHost,
                           c->stream));
    } else {
        TRY(q->sub_off.ensure(8));
        CU(cudaMemsetAsync(q->sub_off.p, 0, 8, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    q->sub_total = total;
    uint64_t items = 0;
    q->max_sub = 0;
    for (uint32_t o = 0; o < n_outer; ++o) {
        uint64_t np = so[o + 1] - so[o];
        q->max_sub = (uint32_t)std::max<uint64_t>(q->max_sub, np);
        uint64_t nh = std::min<uint64_t>((uint64_t)(opo[o + 1] - opo[o]) * limit, cap);
        items += ((np + SCORE_TILE - 1) / SCORE_TILE) * ((nh + SCORE_HCHUNK - 1) / SCORE_HCHUNK + 1);
    }
    REQUIRE(items < (1ull << 31), "too many work items");
    q->items_cap = (uint32_t)std::max<uint64_t>(items, 1);
    TRY(q->n_items_g.ensure(std::max(n_outer, 1u) * 4ull));
    TRY(q->item_off.ensure((n_outer + 1) * 4ull));
    TRY(q->items.ensure((size_t)q->items_cap * sizeof(WorkItem)));
    if (q->p.early_out == 1) {
        const size_t n_tiles = (size_t)(total / 32) + n_outer + 2;
        TRY(q->tile_lo.ensure(n_tiles * 16)); TRY(q->tile_hi.ensure(n_tiles * 16));
    } else if (q->p.early_out == 2) {
        TRY(q->sub_idx_walk.ensure(std::max<uint64_t>(total, 1) * 4));
    }
    if (q->p.icp_top_k) {
        TRY(q->icp.ensure(q->p.icp_top_k));
        TRY(q->topk_ids.ensure(q->p.icp_top_k * 4ull));
        TRY(q->topk_keys.ensure(topk_scratch_bytes(q->cap_hyp, q->p.icp_top_k)));
        TRY(q->icp_T16.ensure(q->p.icp_top_k * 64ull));
    }
    q->ran = false;
    return TM_OK;
}

// export the pose and score of the hypothesis named by out->best (if this shard owns it).  After the count-only
// scorer the score of that one pose is summed here (score_best_kernel); scores[] stays empty until asked for.
static int finalize_best(tm_query* q) {
    tm_ctx* c = q->s->ctx;
    QueryOut* out = q->out.as<QueryOut>();
    const unsigned long long* lazy_acc = nullptr;
    if (q->lazy && !q->scores_valid) {
        ModelDev md;
        TRY(model_dev_for(c, q->m, q->run_thres, &md));
        CU(cudaMemsetAsync(out->best_acc, 0, 16, c->stream));
        launch_score_best(c->stream, q->s->dev, md, q->sub_idx.as<int32_t>(), q->sub_off.as<unsigned long long>(),
                          q->g_hyp.as<uint32_t>(), q->n_outer, q->T.as<float4>(), &out->best, out->shard, q->run_sqt,
                          out->best_acc, q->m->fused);
        lazy_acc = out->best_acc;
    }
    launch_finalize_best(c->stream, &out->best, out->shard, q->T.as<float4>(), q->scores.as<unsigned long long>(),
                         lazy_acc, q->m->dev.cloud.n, out->best_T16, &out->best_score);
    return TM_OK;
}

// scores[] of every hypothesis on request (tm_query_download): re-run the scoring pass with the fused
// count+score kernel over the resident work list.  Counts go to a scratch array and must come out the same.
static int ensure_scores(tm_query* q) {
    if (q->scores_valid || !q->lazy || !q->n_outer) return TM_OK;
    tm_ctx* c = q->s->ctx;
    QueryOut* out = q->out.as<QueryOut>();
    DevBuf& cnt2 = c->scratch[5];
    TRY(cnt2.ensure(q->cap_hyp * 4));
    CU(cudaMemsetAsync(cnt2.p, 0, q->cap_hyp * 4, c->stream));
    CU(cudaMemsetAsync(q->scores.p, 0, q->cap_hyp * 8, c->stream));
    CU(cudaMemsetAsync(&out->work_counter, 0, 4, c->stream));
    ScoreArgs a;
    a.scene = q->s->dev;
    TRY(model_dev_for(c, q->m, q->run_thres, &a.model));
    a.sub_idx = q->sub_idx.as<int32_t>();
    a.items = q->items.as<WorkItem>();
    a.n_items = q->item_off.as<uint32_t>() + q->n_outer;
    a.work_counter = &out->work_counter;
    a.T = q->T.as<float4>();
    a.counts = cnt2.as<uint32_t>();
    a.scores = q->scores.as<unsigned long long>();
    a.sq_thres = q->run_sqt;
    a.stats = nullptr;
    int& b = c->score_bps[q->m->fused ? 1 : 0][1];
    if (!b) b = score_full_max_blocks_per_sm(q->m->fused, true);
    launch_score_full(c->stream, a, c->sm_count * b, q->m->fused, true);
    CU(cudaGetLastError());
    q->scores_valid = true;
    return TM_OK;
}

int tm_query_run(tm_query* q) {
    REQUIRE(q, "null query");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    const tm_model* m = q->m;
    const CloudDev& sc = q->s->dev;
    QueryOut* out = q->out.as<QueryOut>();
    CU(cudaMemsetAsync(out, 0, sizeof(QueryOut), c->stream));
    CU(cudaMemsetAsync(q->counts.p, 0, q->cap_hyp * 4, c->stream));
    CU(cudaMemsetAsync(q->scores.p, 0, q->cap_hyp * 8, c->stream));
    // (a1-a5) pair filter, feature, key, probe
    float lower, upper;
    pair_window(m, q->p.min_diameter_factor, q->p.max_diameter_factor, lower, upper);
    const uint32_t limit = q->p.query_limit ? q->p.query_limit : 200;
    launch_pair_features_probe(c->stream, sc, m->dev, q->outer.as<uint32_t>(),
                               q->pair_outer.as<uint32_t>(), q->pair_j.as<uint32_t>(), q->n_pairs,
                               lower, upper, limit, nullptr, nullptr, q->valid.as<uint8_t>(),
                               q->hit_begin.as<uint32_t>(), q->hit_count.as<uint32_t>(),
                               &out->n_valid);
    if (q->n_pairs == 0) CU(cudaMemsetAsync(q->hyp_off.p, 0, 8, c->stream));
    else
        launch_exclusive_scan_u64(c->stream, q->hit_count.as<uint32_t>(),
                                  q->hyp_off.as<unsigned long long>(), q->n_pairs);
    launch_shard_range(c->stream, q->hyp_off.as<unsigned long long>(), q->n_pairs, q->p.hyp_limit,
                       q->rank, q->world, q->cap_hyp, out->shard, &out->n_local, &out->err);
    launch_group_hyp_ranges(c->stream, q->hyp_off.as<unsigned long long>(),
                            q->outer_pair_off.as<uint32_t>(), q->n_outer, out->shard,
                            q->g_hyp.as<uint32_t>());
    // (a8) radius subsets, only of the outer samples that own hypotheses of this rank's shard
    // (g_hyp); the others get empty rows, so N ranks do not repeat each other's searches
    if (q->n_outer) {
        const float r2 = m->dev.diameter * m->dev.diameter;
        const uint32_t n_seg = (sc.n + BALL_SEG - 1) / BALL_SEG;
        launch_ball_count(c->stream, sc, q->outer.as<uint32_t>(), q->n_outer, q->g_hyp.as<uint32_t>(), r2, n_seg,
                          q->ball_counts.as<uint32_t>());
        launch_ball_seg_scan(c->stream, q->ball_counts.as<uint32_t>(), q->n_outer, n_seg,
                             q->ball_seg_off.as<uint32_t>());
        launch_exclusive_scan_u64(c->stream, q->ball_seg_off.as<uint32_t>(), q->sub_off.as<unsigned long long>(),
                                  q->n_outer);
        launch_ball_fill(c->stream, sc, q->outer.as<uint32_t>(), q->n_outer, q->g_hyp.as<uint32_t>(), r2, n_seg,
                         q->ball_counts.as<uint32_t>(), q->sub_off.as<unsigned long long>(),
                         q->sub_idx.as<int32_t>());
    }
    // (a6, a7) hypotheses
    launch_hypotheses(c->stream, sc, m->dev, q->outer.as<uint32_t>(), q->pair_outer.as<uint32_t>(),
                      q->pair_j.as<uint32_t>(), q->n_pairs, q->hyp_off.as<unsigned long long>(),
                      q->hit_begin.as<uint32_t>(), m->dev.hits, q->p.force_up, out->shard,
                      q->T.as<float4>(), q->hyp_valid.as<uint8_t>(), q->hyp_pair.as<uint32_t>());
    // (a10) scoring
    const float thres = q->p.dist_thres * m->dev.resolution;
    const float sqt = sq_threshold(thres);
    q->lazy = false;
    q->run_thres = thres;
    q->run_sqt = sqt;
    if (q->n_outer) {
        if (!q->p.early_out) {
            launch_work_count(c->stream, q->sub_off.as<unsigned long long>(),
                              q->g_hyp.as<uint32_t>(), q->n_outer, q->n_items_g.as<uint32_t>(),
                              &out->n_tests);
            launch_exclusive_scan_u32(c->stream, q->n_items_g.as<uint32_t>(),
                                      q->item_off.as<uint32_t>(), q->n_outer);
            launch_work_fill(c->stream, q->sub_off.as<unsigned long long>(), q->g_hyp.as<uint32_t>(),
                             q->n_outer, q->item_off.as<uint32_t>(), q->items.as<WorkItem>());
            ScoreArgs a;
            a.scene = sc;
            TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &a.model));
            a.sub_idx = q->sub_idx.as<int32_t>();
            a.items = q->items.as<WorkItem>();
            a.n_items = q->item_off.as<uint32_t>() + q->n_outer;
            a.work_counter = &out->work_counter;
            a.T = q->T.as<float4>();
            a.counts = q->counts.as<uint32_t>();
            a.scores = q->scores.as<unsigned long long>();
            a.sq_thres = sqt;
            a.stats = nullptr;
            if (knobs().score_stats) {
                TRY(q->stats.ensure(64));
                CU(cudaMemsetAsync(q->stats.p, 0, 64, c->stream));
                a.stats = q->stats.as<unsigned long long>();
            }
            q->lazy = knobs().scorer >= 8;
            if (q->lazy) {
                int& b = c->count_bps[m->fused ? 1 : 0];
                if (!b) b = score_count_x2_max_blocks_per_sm(m->fused);
                const int grid = knobs().score_grid ? knobs().score_grid : c->sm_count * b;
                CU(cudaEventRecord(q->ev_s0, c->stream));
                launch_score_count_x2(c->stream, a, grid, m->fused);
                CU(cudaEventRecord(q->ev_s1, c->stream));
            } else {
                int& b = c->score_bps[m->fused ? 1 : 0][1];
                if (!b) b = score_full_max_blocks_per_sm(m->fused, true);
                const int grid = knobs().score_grid ? knobs().score_grid : c->sm_count * b;
                CU(cudaEventRecord(q->ev_s0, c->stream));
                launch_score_full(c->stream, a, grid, m->fused, true);
                CU(cudaEventRecord(q->ev_s1, c->stream));
            }
        } else {
            launch_group_of_hyp(c->stream, q->g_hyp.as<uint32_t>(), q->n_outer,
                                q->g_of_hyp.as<uint32_t>());
            EarlyArgs a;
            a.sub_idx = q->sub_idx.as<int32_t>();
            if (q->p.early_out == 2) {  // evenly sampling walk order (see tm_score)
                launch_walk_order_rows(c->stream, q->sub_idx.as<int32_t>(), q->sub_off.as<unsigned long long>(),
                                       q->n_outer, q->max_sub, q->sub_idx_walk.as<int32_t>());
                a.sub_idx = q->sub_idx_walk.as<int32_t>();
            } else {
                launch_subset_tile_boxes(c->stream, sc, q->sub_idx.as<int32_t>(), q->sub_off.as<unsigned long long>(),
                                         q->n_outer, q->max_sub, q->tile_lo.as<float4>(), q->tile_hi.as<float4>());
                a.tile_lo = q->tile_lo.as<float4>();
                a.tile_hi = q->tile_hi.as<float4>();
            }
            a.scene = sc;
            TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &a.model));
            a.sub_off = q->sub_off.as<unsigned long long>();
            a.g_of_hyp = q->g_of_hyp.as<uint32_t>();
            a.T = q->T.as<float4>();
            a.n_hyp = (uint32_t)q->cap_hyp;  // grid bound; the kernel clips to n_local
            a.n_hyp_dev = &out->n_local;
            a.n_tests = &out->n_tests;
            a.sq_thres = sqt;
            a.accept_prob = q->p.accept_prob;
            a.early_out = 1;
            a.counts = q->counts.as<uint32_t>();
            a.scores = q->scores.as<unsigned long long>();
            a.dropped = q->dropped.as<uint8_t>();
            a.tested = nullptr;
            CU(cudaEventRecord(q->ev_s0, c->stream));
            launch_score_early_drop(c->stream, a, m->fused);
            CU(cudaEventRecord(q->ev_s1, c->stream));
        }
    }
    launch_argmax(c->stream, q->counts.as<uint32_t>(), q->hyp_valid.as<uint8_t>(), &out->n_local,
                  out->shard, &out->best, c->sm_count * 2);
    // (a12) ICP of the local top-k
    if (q->p.icp_top_k && q->p.max_icp_iterations) {
        // a hypothesis the early drop gave up on never becomes a candidate (scene.hpp:330: a dropped
        // project_ returns fewer correspondences than the acceptance bound), whatever its partial count
        launch_select_topk(c->stream, q->counts.as<uint32_t>(), q->hyp_valid.as<uint8_t>(),
                           q->p.early_out ? q->dropped.as<uint8_t>() : nullptr,
                           &out->n_local, q->cap_hyp, q->p.icp_top_k, q->topk_ids.as<uint32_t>(),
                           q->topk_keys.as<unsigned long long>());
        launch_gather_rows(c->stream, q->T.as<float4>(), q->topk_ids.as<uint32_t>(), q->p.icp_top_k,
                           q->icp.Tcur.as<float4>(), q->icp.active.as<uint32_t>());
        TRY(icp_enqueue(c, sc, m, q->icp, q->p.icp_top_k, q->p.max_icp_iterations, q->p.dist_thres));
    }
    TRY(finalize_best(q));
    CU(cudaGetLastError());
    q->ran = true;
    q->scores_valid = !q->lazy;
    return TM_OK;
}

int tm_query_result_get(tm_query* q, tm_query_result* r) {
    REQUIRE(q && r, "null argument");
    REQUIRE(q->ran, "tm_query_result_get before tm_query_run");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    TRY(pinned_ensure(c, sizeof(QueryOut) + 16));
    CU(cudaMemcpyAsync(c->pinned, q->out.p, sizeof(QueryOut), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(&q->host_out, c->pinned, sizeof(QueryOut));
    const QueryOut& o = q->host_out;
    if (q->stats.p && knobs().score_stats) {
        unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        CU(cudaMemcpy(st, q->stats.p, 64, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[tm stats] (warp-tile,hyp) pairs %llu  survive cull %llu (%.1f%%)  with inliers %llu (%.1f%%)"
                "  all-inlier tiles %llu  >=90%% %llu  inliers %llu\n",
                st[0], st[1], st[0] ? 100.0 * st[1] / st[0] : 0.0, st[2], st[0] ? 100.0 * st[2] / st[0] : 0.0, st[3], st[4],
                st[5]);
    }
    if (o.err) return fail(TM_ERR_CAPACITY, "query: hypothesis capacity exceeded (max_hypotheses)");
    memset(r, 0, sizeof(*r));
    r->n_pairs_valid = o.n_valid;
    r->n_hypotheses = o.shard[2];
    r->n_scored = o.n_local;
    r->n_tests = o.n_tests;
    r->best_key = o.best;
    if (o.best) {
        r->best_inliers = (uint32_t)(o.best >> 32);
        r->best_hypothesis = 0xFFFFFFFFu - (uint32_t)(o.best & 0xFFFFFFFFull);
        r->best_score = o.best_score;
        memcpy(r->best_T, o.best_T16, 64);
    }
    return TM_OK;
}
int tm_query_score_kernel_ms(tm_query* q, float* ms) {
    REQUIRE(q && ms && q->ran, "tm_query_score_kernel_ms: null/unrun query");
    TRY(bind(q->s->ctx));
    *ms = 0.f;
    if (!q->n_outer) return TM_OK;
    CU(cudaEventSynchronize(q->ev_s1));
    CU(cudaEventElapsedTime(ms, q->ev_s0, q->ev_s1));
    return TM_OK;
}
void* tm_query_best_key_device(tm_query* q) {
    return q ? (void*)&q->out.as<QueryOut>()->best : nullptr;
}
int tm_query_set_global_best(tm_query* q, uint64_t key) {
    REQUIRE(q && q->ran, "null/unrun query");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    QueryOut* out = q->out.as<QueryOut>();
    CU(cudaMemcpyAsync(&out->best, &key, 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(out->best_T16, 0, 64, c->stream));
    CU(cudaMemsetAsync(&out->best_score, 0, 8, c->stream));
    TRY(finalize_best(q));
    CU(cudaGetLastError());
    return TM_OK;
}

int tm_query_download(tm_query* q, uint64_t capacity, uint32_t* counts, double* scores,
                      float* T16s, uint8_t* valid, uint32_t* hyp_pair, uint8_t* dropped) {
    REQUIRE(q && q->ran, "null/unrun query");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    tm_query_result r;
    TRY(tm_query_result_get(q, &r));
    const uint64_t n = r.n_scored;
    if (n > capacity) return fail(TM_ERR_CAPACITY, "tm_query_download: capacity too small");
    if (!n) return TM_OK;
    if (counts) CU(cudaMemcpyAsync(counts, q->counts.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (valid) CU(cudaMemcpyAsync(valid, q->hyp_valid.p, n, cudaMemcpyDeviceToHost, c->stream));
    if (hyp_pair) CU(cudaMemcpyAsync(hyp_pair, q->hyp_pair.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (dropped) {
        if (q->p.early_out) CU(cudaMemcpyAsync(dropped, q->dropped.p, n, cudaMemcpyDeviceToHost, c->stream));
        else memset(dropped, 0, n);
    }
    std::vector<unsigned long long> fx;
    std::vector<uint8_t> dr;
    if (scores && q->p.early_out) {
        dr.resize(n);
        CU(cudaMemcpyAsync(dr.data(), q->dropped.p, n, cudaMemcpyDeviceToHost, c->stream));
    }
    if (scores) {
        TRY(ensure_scores(q));
        fx.resize(n);
        CU(cudaMemcpyAsync(fx.data(), q->scores.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    if (T16s) {
        DevBuf& d16 = c->scratch[0];
        TRY(d16.ensure(n * 64));
        launch_colmajor_from_rows(c->stream, q->T.as<float4>(), n, d16.as<float>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(T16s, d16.p, n * 64, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    if (scores)
        for (uint64_t i = 0; i < n; ++i) {
            double v = (double)fx[i] / SCORE_SCALE;
            scores[i] = (!dr.empty() && dr[i]) ? v : v / (double)q->m->dev.cloud.n;
        }
    return TM_OK;
}

int tm_query_icp_results(tm_query* q, uint32_t* hyp_ids, float* T16s, uint32_t* counts,
                         double* scores, uint32_t* iters) {
    REQUIRE(q && q->ran, "null/unrun query");
    const uint32_t k = q->p.icp_top_k;
    REQUIRE(k && q->p.max_icp_iterations, "query has no ICP stage");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    std::vector<long long> sums((size_t)k * ICP_NSUM);
    std::vector<uint32_t> ids(k);
    launch_colmajor_from_rows(c->stream, q->icp.Tbest.as<float4>(), k, q->icp_T16.as<float>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sums.data(), q->icp.sums_best.p, sums.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ids.data(), q->topk_ids.p, k * 4ull, cudaMemcpyDeviceToHost, c->stream));
    if (T16s) CU(cudaMemcpyAsync(T16s, q->icp_T16.p, k * 64ull, cudaMemcpyDeviceToHost, c->stream));
    if (iters) CU(cudaMemcpyAsync(iters, q->icp.iters.p, k * 4ull, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint32_t r = 0; r < k; ++r) {
        if (hyp_ids) hyp_ids[r] = ids[r];
        if (counts) counts[r] = (uint32_t)sums[(size_t)r * ICP_NSUM];
        if (scores)
            scores[r] = (double)sums[(size_t)r * ICP_NSUM + 16] / SCORE_SCALE / (double)q->m->dev.cloud.n;
    }
    return TM_OK;
}

// ------------------------------------------------------------------- NCCL
// The one collective of the path.  NCCL is resolved at run time
// with dlopen so the library loads in processes that never go multi-GPU.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclUint8_ = 1, ncclInt64_ = 4, ncclUint64_ = 5 };  // ncclDataType_t
enum { ncclSum_ = 0, ncclMax_ = 2 };       // ncclRedOp_t
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;
static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return TM_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names)
        if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) return fail(TM_ERR_NCCL, std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(
        lib, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        return fail(TM_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.lib = lib;
    return TM_OK;
}
#define NC(call)                                                                             \
    do {                                                                                     \
        int r_ = (call);                                                                     \
        if (r_ != 0)                                                                         \
            return fail(TM_ERR_NCCL, std::string(#call) + ": " +                            \
                                         (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?")); \
    } while (0)

struct tm_comm {
    tm_ctx* ctx;
    ncclComm_t comm;
    int rank, world;
    DevBuf stage;  // batched best-pose reduce: n keys, then n x (score, pose)
};

int tm_nccl_unique_id(uint8_t out[128]) {
    REQUIRE(out, "null out");
    TRY(nccl_load());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return TM_OK;
}
int tm_comm_create(tm_ctx* c, const uint8_t idb[128], int rank, int world, tm_comm** out) {
    REQUIRE(c && idb && out && world > 0 && rank >= 0 && rank < world, "tm_comm_create: bad argument");
    TRY(nccl_load());
    TRY(bind(c));
    ncclUniqueId id;
    memcpy(id.internal, idb, 128);
    tm_comm* cm = new tm_comm{c, nullptr, rank, world, DevBuf()};
    int r = g_nccl.CommInitRank(&cm->comm, world, id, rank);
    if (r != 0) {
        delete cm;
        return fail(TM_ERR_NCCL, std::string("ncclCommInitRank: ") +
                                     (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    }
    *out = cm;
    return TM_OK;
}
void tm_comm_destroy(tm_comm* cm) {
    if (!cm) return;
    cudaSetDevice(cm->ctx->device);
    if (cm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(cm->comm);
    cm->stage.release();
    delete cm;
}
// batched form for several queries over the same scene (BASELINE configs[3]: 16 models x one
// scene): ONE max all-reduce over the n packed keys and ONE sum all-reduce over the n x 72-byte
// (score, pose) records instead of 2n collectives
int tm_queries_allreduce_best(tm_query** qs, uint32_t n, tm_comm* cm) {
    REQUIRE(cm && (n == 0 || qs), "tm_queries_allreduce_best: null argument");
    if (!n) return TM_OK;
    tm_ctx* c = cm->ctx;
    for (uint32_t i = 0; i < n; ++i)
        REQUIRE(qs[i] && qs[i]->ran && qs[i]->s->ctx == c, "tm_queries_allreduce_best: bad query");
    TRY(bind(c));
    TRY(cm->stage.ensure((size_t)n * 8 + (size_t)n * 72));
    unsigned long long* keys = cm->stage.as<unsigned long long>();
    uint8_t* recs = reinterpret_cast<uint8_t*>(keys + n);
    for (uint32_t i = 0; i < n; ++i)
        CU(cudaMemcpyAsync(keys + i, &qs[i]->out.as<QueryOut>()->best, 8, cudaMemcpyDeviceToDevice, c->stream));
    NC(g_nccl.AllReduce(keys, keys, n, ncclUint64_, ncclMax_, cm->comm, c->stream));
    for (uint32_t i = 0; i < n; ++i) {
        tm_query* q = qs[i];
        QueryOut* out = q->out.as<QueryOut>();
        CU(cudaMemcpyAsync(&out->best, keys + i, 8, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemsetAsync(out->best_T16, 0, 64, c->stream));
        CU(cudaMemsetAsync(&out->best_score, 0, 8, c->stream));
        TRY(finalize_best(q));
        CU(cudaMemcpyAsync(recs + 72 * (size_t)i, &out->best_score, 72, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaGetLastError());
    NC(g_nccl.AllReduce(recs, recs, (size_t)n * 72, ncclUint8_, ncclSum_, cm->comm, c->stream));
    for (uint32_t i = 0; i < n; ++i)
        CU(cudaMemcpyAsync(&qs[i]->out.as<QueryOut>()->best_score, recs + 72 * (size_t)i, 72,
                           cudaMemcpyDeviceToDevice, c->stream));
    return TM_OK;
}
static int comm_allreduce_sum_i64(tm_comm* cm, void* buf, size_t count, cudaStream_t st) {
    NC(g_nccl.AllReduce(buf, buf, count, ncclInt64_, ncclSum_, cm->comm, st));
    return TM_OK;
}
int tm_icp_sharded(tm_scene* s, tm_model* m, tm_comm* cm, const float* T16s, uint32_t n,
                   uint32_t max_iterations, float dist_thres, uint32_t pt_begin, uint32_t pt_end,
                   uint64_t n_scene_total, uint32_t emulate_parts, float* T16s_out, uint32_t* counts,
                   double* scores, uint32_t* iters) {
    REQUIRE(s && m, "null handle");
    REQUIRE(pt_begin <= pt_end && pt_end <= s->dev.n, "tm_icp_sharded: bad point range");
    REQUIRE(n_scene_total >= (uint64_t)(pt_end - pt_begin), "tm_icp_sharded: n_scene_total too small");
    REQUIRE(!cm || cm->ctx == s->ctx, "communicator belongs to another context");
    REQUIRE(!(cm && emulate_parts > 1), "tm_icp_sharded: emulate_parts is for single-process runs");
    REQUIRE(max_iterations > 0, "tm_icp_sharded: max_iterations must be > 0");
    IcpSplit sp;
    sp.pt_begin = pt_begin; sp.pt_end = pt_end; sp.n_total = n_scene_total; sp.comm = cm;
    sp.emulate = std::max(1u, emulate_parts);
    return icp_run(s, m, T16s, n, max_iterations, dist_thres, T16s_out, counts, scores, iters, &sp);
}
int tm_query_allreduce_best(tm_query* q, tm_comm* cm) {
    REQUIRE(q && cm && q->ran, "tm_query_allreduce_best: bad argument");
    tm_ctx* c = q->s->ctx;
    REQUIRE(c == cm->ctx, "communicator belongs to another context");
    TRY(bind(c));
    QueryOut* out = q->out.as<QueryOut>();
    // max over ranks of (inliers << 32 | ~global id): 8 bytes, latency-bound
    NC(g_nccl.AllReduce(&out->best, &out->best, 1, ncclUint64_, ncclMax_, cm->comm, c->stream));
    // the owner re-exports the winning pose; everybody else contributes zeros
    CU(cudaMemsetAsync(out->best_T16, 0, 64, c->stream));
    CU(cudaMemsetAsync(&out->best_score, 0, 8, c->stream));
    TRY(finalize_best(q));
    CU(cudaGetLastError());
    // best_score (8 B) and best_T16 (64 B) are adjacent in QueryOut
    NC(g_nccl.AllReduce(&out->best_score, &out->best_score, 72, ncclUint8_, ncclSum_, cm->comm,
                        c->stream));
    return TM_OK;
}

}  // extern "C"
