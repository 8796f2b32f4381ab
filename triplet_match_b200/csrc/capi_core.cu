// capi_core.cu — contexts, resident clouds / models / scenes and the stage calls with host buffers
// (include/tm_b200.h up to the resident query).
#include "capi_internal.cuh"

namespace tmk {
std::atomic<unsigned long long> g_launch_count{0};
}

static thread_local std::string g_err;
int tm_fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

const Knobs& knobs() {
    static const Knobs k = [] {
        Knobs v;
        if (const char* e = getenv("TM_OCC")) v.occ = atoi(e) != 0;
        if (const char* e = getenv("TM_FUSED_GRID")) v.fused_grid = atoi(e) != 0 ? 1 : 0;
        if (const char* e = getenv("TM_SCORE_GRID")) v.score_grid = std::max(1, atoi(e));
        v.score_stats = getenv("TM_SCORE_STATS") != nullptr;
        if (const char* e = getenv("TM_SCORER")) v.scorer = atoi(e);
        if (const char* e = getenv("TM_EARLY_LEVELS")) v.early_levels = atoi(e) != 0;
        if (const char* e = getenv("TM_EARLY_MERGE")) v.early_merge = atoi(e) != 0;
        return v;
    }();
    return k;
}

int bind(tm_ctx* c) {
    CU(cudaSetDevice(c->device));
    return TM_OK;
}
// ModelDev for kernels that test against `thres`: the resident description plus, when it pays, the
// block-occupancy mask of that threshold (built on first use, cached per model).  TM_OCC=0 disables.
int model_dev_for(tm_ctx* c, tm_model* m, float thres, ModelDev* out) {
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
int pinned_ensure(tm_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_cap) return TM_OK;
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->pinned_cap = 0;
    CU(cudaMallocHost(&c->pinned, bytes));
    c->pinned_cap = bytes;
    ++c->pinned_gen;
    return TM_OK;
}

// `dist > thres` with dist = sqrtf(sq) (scene.hpp:464-465) <=> sq > S, where S is the
// largest float whose correctly rounded square root is <= thres.
float sq_threshold(float thres) {
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
    c->icp.release();
    c->icp_d16.release();
    c->icp_pack.release();
    c->icp_graph.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->pinned_gather) cudaFreeHost(c->pinned_gather);
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
    // fused grid (cell -> nearest model point, 16 B) when it stays L2-sized — or, up to 2 GiB, for the large grids of
    // surface models: the occupancy mask keeps the scorers away from all but the shell of cells around the surface,
    // and that shell is what has to stay in L2, not the grid (C3: 21.5 M cells = 344 MB fused, 117 -> 109 ms per 2^20
    // hypotheses against the index grid + a second gather)
    m->fused = cells * sizeof(float4) <= (2048ull << 20);
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
        DevBuf ka, va, kb, vb, hist, offs, p2, n2, t2, scr;
        auto release = [&] { for (DevBuf* b : {&ka, &va, &kb, &vb, &hist, &offs, &p2, &n2, &t2, &scr}) b->release(); };
        rc = ka.ensure((size_t)n * 4);
        if (!rc) rc = va.ensure((size_t)n * 4);
        if (!rc) rc = kb.ensure((size_t)n * 4);
        if (!rc) rc = vb.ensure((size_t)n * 4);
        if (!rc) rc = hist.ensure(((size_t)256 * nb + 1) * 4);
        if (!rc) rc = offs.ensure(((size_t)256 * nb + 1) * 4);
        if (!rc) rc = p2.ensure((size_t)n * 16);
        if (!rc) rc = n2.ensure((size_t)n * 16);
        if (!rc) rc = t2.ensure((size_t)n * 16);
        if (!rc) rc = scr.ensure(scan_scratch_bytes((uint64_t)256 * nb));
        if (rc) { release(); return bail(rc); }
        launch_morton_codes(c->stream, s->dev.pos, n, blo, inv, ka.as<uint32_t>(), va.as<uint32_t>());
        launch_radix_sort_pairs(c->stream, ka.as<uint32_t>(), va.as<uint32_t>(), kb.as<uint32_t>(), vb.as<uint32_t>(), n,
                                hist.as<uint32_t>(), offs.as<uint32_t>(), scr.as<unsigned long long>(), c->sm_count);
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
void pair_window(const tm_model* m, float min_df, float max_df, float& lower, float& upper) {
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
int ball_subsets_dev(tm_ctx* c, const CloudDev& scene, const uint32_t* d_centres,
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
    a.thres = thres;
    a.cell_reach = cell_reach_of(a.model);
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
uint32_t tm_early_level_begin(uint32_t n, int level) { return level_begin(n, level); }

int tm_correspondences_batch(tm_scene* s, tm_model* m, const float* T16s, uint32_t n_T, float dist_thres,
                             uint64_t* offsets, uint32_t* scene_corrs, uint32_t* model_corrs, uint64_t capacity,
                             double* scores) {
    REQUIRE(s && m && offsets && (n_T == 0 || T16s), "tm_correspondences_batch: null argument");
    REQUIRE((scene_corrs == nullptr) == (model_corrs == nullptr), "tm_correspondences_batch: pass both lists or neither");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    for (uint32_t t = 0; t <= n_T; ++t) offsets[t] = 0;
    if (scores) for (uint32_t t = 0; t < n_T; ++t) scores[t] = 0.0;
    if (!s->dev.n || !n_T) return TM_OK;
    REQUIRE((uint64_t)n_T * s->dev.n < (1ull << 32), "tm_correspondences_batch: too many transforms for one call");
    const uint32_t n_seg = (s->dev.n + CORR_SEG - 1) / CORR_SEG;
    const size_t slots = (size_t)n_T * n_seg;
    DevBuf &cnt = c->scratch[0], &off = c->scratch[1], &sc = c->scratch[2], &mc = c->scratch[3], &acc = c->scratch[4],
           &rows = c->scratch[5];
    TRY(cnt.ensure(slots * 4)); TRY(off.ensure((slots + 1) * 4)); TRY(acc.ensure(n_T * 8ull)); TRY(rows.ensure(n_T * 48ull));
    std::vector<float4> hrows(3 * (size_t)n_T);
    for (uint32_t t = 0; t < n_T; ++t) {
        const float* T16 = T16s + 16 * (size_t)t;
        hrows[3 * t] = make_float4(T16[0], T16[4], T16[8], T16[12]);
        hrows[3 * t + 1] = make_float4(T16[1], T16[5], T16[9], T16[13]);
        hrows[3 * t + 2] = make_float4(T16[2], T16[6], T16[10], T16[14]);
    }
    const float sqt = sq_threshold(dist_thres * m->dev.resolution);
    CU(cudaMemcpyAsync(rows.p, hrows.data(), n_T * 48ull, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(acc.p, 0, n_T * 8ull, c->stream));
    launch_corr_count(c->stream, s->dev, m->dev, rows.as<float4>(), n_T, sqt, n_seg, cnt.as<uint32_t>(),
                      acc.as<unsigned long long>(), m->fused);
    launch_exclusive_scan_u32(c->stream, cnt.as<uint32_t>(), off.as<uint32_t>(), slots);
    CU(cudaGetLastError());
    std::vector<uint32_t> hoff(slots + 1);
    std::vector<unsigned long long> fx(n_T);
    CU(cudaMemcpyAsync(hoff.data(), off.p, (slots + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(fx.data(), acc.p, n_T * 8ull, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint32_t t = 0; t <= n_T; ++t) offsets[t] = hoff[(size_t)t * n_seg];
    if (scores) for (uint32_t t = 0; t < n_T; ++t) scores[t] = (double)fx[t] / SCORE_SCALE / (double)m->dev.cloud.n;
    const uint64_t total = offsets[n_T];
    if (!scene_corrs || !total) return TM_OK;
    if (capacity < total) return fail(TM_ERR_CAPACITY, "tm_correspondences_batch: capacity too small (size with NULL lists first)");
    TRY(sc.ensure(total * 4)); TRY(mc.ensure(total * 4));
    launch_corr_fill(c->stream, s->dev, m->dev, rows.as<float4>(), n_T, sqt, n_seg, off.as<uint32_t>(), sc.as<uint32_t>(),
                     mc.as<uint32_t>(), m->fused);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(scene_corrs, sc.p, total * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(model_corrs, mc.p, total * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

int tm_correspondences(tm_scene* s, tm_model* m, const float* T16, float dist_thres,
                       uint32_t* scene_corrs, uint32_t* model_corrs, uint32_t* n_corr,
                       double* score) {
    REQUIRE(s && m && T16 && n_corr, "tm_correspondences: null argument");
    *n_corr = 0;
    uint64_t offsets[2] = {0, 0};
    // lists are optional here; the batch call wants both or neither
    const bool lists = scene_corrs && model_corrs;
    if (!lists && (scene_corrs || model_corrs)) {
        std::vector<uint32_t> other(s->dev.n ? s->dev.n : 1);
        TRY(tm_correspondences_batch(s, m, T16, 1, dist_thres, offsets, scene_corrs ? scene_corrs : other.data(),
                                     model_corrs ? model_corrs : other.data(), s->dev.n, score));
    } else {
        TRY(tm_correspondences_batch(s, m, T16, 1, dist_thres, offsets, scene_corrs, model_corrs, s->dev.n, score));
    }
    *n_corr = (uint32_t)offsets[1];
    return TM_OK;
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

// self-test hook: exclusive prefix sum of `in` (n x u32) with the chained multi-CTA scan the resident query uses for
// long pair lists; out: n + 1 x u64
int tm_ctx_scan_u64(tm_ctx* c, const uint32_t* in, uint64_t n, uint64_t* out) {
    REQUIRE(c && out && (n == 0 || in), "tm_ctx_scan_u64: null argument");
    TRY(bind(c));
    DevBuf &din = c->scratch[0], &dout = c->scratch[1], &scr = c->scratch[2];
    TRY(din.ensure(std::max<uint64_t>(n, 1) * 4)); TRY(dout.ensure((n + 1) * 8)); TRY(scr.ensure(scan_scratch_bytes(n)));
    if (n) CU(cudaMemcpyAsync(din.p, in, n * 4, cudaMemcpyHostToDevice, c->stream));
    launch_exclusive_scan_u64_chained(c->stream, din.as<uint32_t>(), dout.as<unsigned long long>(), n,
                                      scr.as<unsigned long long>(), c->sm_count);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, (n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TM_OK;
}

// self-test hook: the top-k selection of the resident query's ICP stage (by count descending, index ascending; zero
// counts and excluded entries never selected) on caller-supplied counts; ids: k x u32, 0xFFFFFFFF where fewer than k
// qualify
int tm_ctx_select_topk(tm_ctx* c, const uint32_t* counts, const uint8_t* excluded, uint32_t n, uint32_t k, uint32_t* ids) {
    REQUIRE(c && ids && (n == 0 || counts) && k > 0 && k <= 4096, "tm_ctx_select_topk: bad argument");
    TRY(bind(c));
    DevBuf &dc = c->scratch[0], &dx = c->scratch[1], &dn = c->scratch[2], &did = c->scratch[3], &scr = c->scratch[4];
    const uint64_t cap = std::max<uint32_t>(n, 1);
    TRY(dc.ensure(cap * 4)); TRY(dx.ensure(cap)); TRY(dn.ensure(4)); TRY(did.ensure(k * 4ull));
    TRY(scr.ensure(topk_scratch_bytes(cap, k)));
    if (n) CU(cudaMemcpyAsync(dc.p, counts, n * 4ull, cudaMemcpyHostToDevice, c->stream));
    if (n && excluded) CU(cudaMemcpyAsync(dx.p, excluded, n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dn.p, &n, 4, cudaMemcpyHostToDevice, c->stream));
    launch_select_topk(c->stream, dc.as<uint32_t>(), nullptr, excluded ? dx.as<uint8_t>() : nullptr, dn.as<uint32_t>(), cap, k,
                       did.as<uint32_t>(), scr.as<unsigned long long>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ids, did.p, k * 4ull, cudaMemcpyDeviceToHost, c->stream));
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
    DevBuf& scr = c->scratch[7];
    TRY(scr.ensure(scan_scratch_bytes(n)));
    launch_tangent_candidates(c->stream, s->dev.tgt, n, fl.as<uint32_t>());
    launch_exclusive_scan_u32_chained(c->stream, fl.as<uint32_t>(), off.as<uint32_t>(), n, scr.as<unsigned long long>(),
                                      c->sm_count);
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

}  // extern "C"
