// capi_icp.cu — icp_ (include/impl/scene.hpp:369-404): the accumulate / step loop on the device, for host
// pose lists (tm_icp, tm_icp_sharded, tm_icp_pose_sharded) and for the resident query's top-k stage
// (icp_enqueue).
//
// A refinement of k poses is 2 * (1 + max_iterations) small launches plus copies; at the sizes of BASELINE
// configs[4] (64 poses, 10 M points) their launch latency, not their work, sets the time.  So the host-list
// entry points keep their buffers in the context (no allocation per call), stage poses and results through
// the context's pinned block, and replay the whole sequence as ONE cached CUDA graph (IcpGraph).  With a
// communicator (scene-sharded sums all-reduced every iteration) the sequence is enqueued launch by launch.
#include "capi_internal.cuh"

static double icp_fix_scale(const tm_model* m, uint32_t n_scene, float thres) {
    // |s'|,|m'| <= r = half bbox diagonal + thres; n * r^2 * 2^bits < 2^62
    double r = (double)m->half_diag + (double)thres + 1e-6;
    double bound = std::max(1.0, (double)std::max(n_scene, 1u) * std::max(r * r, r));
    int bits = (int)std::floor(62.0 - std::log2(bound));
    bits = std::max(8, std::min(40, bits));
    return std::ldexp(1.0, bits);
}

// enqueue the ICP loop for k transforms already in b.Tcur with b.active set; mdev = model_dev_for(thres)
static int icp_enqueue_with(tm_ctx* c, const CloudDev& scene, const tm_model* m, const ModelDev& mdev, IcpBufs& b,
                            uint32_t k, uint32_t max_iterations, float thres, const IcpSplit& sp, uint32_t* n_kernels) {
    const float sqt = sq_threshold(thres);
    const double fs = icp_fix_scale(m, (uint32_t)std::min<uint64_t>(sp.n_total, 0xffffffffull), thres);
    CU(cudaMemsetAsync(b.sums_cur.p, 0, (size_t)k * ICP_NSUM * 8, c->stream));
    CU(cudaMemsetAsync(b.sums_best.p, 0, (size_t)k * ICP_NSUM * 8, c->stream));
    CU(cudaMemsetAsync(b.iters.p, 0, (size_t)k * 4, c->stream));
    const int grid = c->sm_count * 4;
    IcpState st = b.state();
    const uint32_t parts = std::max(1u, sp.emulate);
    const uint64_t span = sp.pt_end - sp.pt_begin;
    uint32_t nk = 0;
    for (uint32_t it = 0; it <= max_iterations; ++it) {
        for (uint32_t w = 0; w < parts; ++w) {
            const uint32_t b0 = sp.pt_begin + (uint32_t)(span * w / parts);
            const uint32_t b1 = sp.pt_begin + (uint32_t)(span * (w + 1) / parts);
            if (b1 > b0) {
                launch_icp_accumulate(c->stream, scene, mdev, st.Tcur, st.active, k, b0, b1, sqt, m->centre[0],
                                      m->centre[1], m->centre[2], fs, st.sums_cur, b.pair_list(), b.pair_count(), grid,
                                      m->fused);
                nk += 2;
            }
        }
        if (sp.comm) TRY(comm_allreduce_sum_i64(sp.comm, st.sums_cur, (size_t)k * ICP_NSUM, c->stream));
        launch_icp_step(c->stream, st, k, it == 0 ? 1 : 0, max_iterations, 1.0 / fs, m->centre[0], m->centre[1],
                        m->centre[2]);
        ++nk;
    }
    if (n_kernels) *n_kernels = nk;
    return TM_OK;
}

int icp_enqueue(tm_ctx* c, const CloudDev& scene, const tm_model* m, IcpBufs& b, uint32_t k, uint32_t max_iterations,
                float dist_thres, const IcpSplit* split) {
    const float thres = (2 * dist_thres) * m->dev.resolution;  // scene.hpp:373 + :413
    IcpSplit sp;
    if (split) sp = *split;
    else { sp.pt_end = scene.n; sp.n_total = scene.n; }
    ModelDev mdev;
    TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &mdev));
    TRY(b.pairs.ensure(icp_pairs_bytes(sp.pt_begin, sp.pt_end, k)));
    TRY(icp_enqueue_with(c, scene, m, mdev, b, k, max_iterations, thres, sp, nullptr));
    CU(cudaGetLastError());
    return TM_OK;
}

// layout of the pinned staging block of one refinement of n poses
struct IcpStage {
    float* T_in;        // n x 16, column-major
    float* T_out;       // n x 16
    long long* sums;    // n x ICP_NSUM
    uint32_t* iters;    // n
    static size_t bytes(uint32_t n) { return (size_t)n * (64 + 64 + ICP_NSUM * 8 + 4) + 64; }
    IcpStage(void* base, uint32_t n) {
        uint8_t* p = static_cast<uint8_t*>(base);
        T_in = reinterpret_cast<float*>(p);
        T_out = reinterpret_cast<float*>(p + (size_t)n * 64);
        sums = reinterpret_cast<long long*>(p + (size_t)n * 128);
        iters = reinterpret_cast<uint32_t*>(p + (size_t)n * (128 + ICP_NSUM * 8));
    }
};

// everything between "poses are in the pinned block" and "results are in the pinned block"
static int icp_sequence(tm_ctx* c, tm_scene* s, tm_model* m, const ModelDev& mdev, uint32_t n, uint32_t max_iterations,
                        float thres, const IcpSplit& sp, const IcpStage& hs, uint32_t* n_kernels) {
    IcpBufs& b = c->icp;
    CU(cudaMemcpyAsync(c->icp_d16.p, hs.T_in, (size_t)n * 64, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(b.active.p, 1, (size_t)n * 4, c->stream));  // any non-zero word = active
    launch_rows_from_colmajor(c->stream, c->icp_d16.as<float>(), n, b.Tcur.as<float4>());
    uint32_t nk = 0;
    TRY(icp_enqueue_with(c, s->dev, m, mdev, b, n, max_iterations, thres, sp, &nk));
    launch_colmajor_from_rows(c->stream, b.Tbest.as<float4>(), n, c->icp_d16.as<float>());
    CU(cudaMemcpyAsync(hs.T_out, c->icp_d16.p, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hs.sums, b.sums_best.p, (size_t)n * ICP_NSUM * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hs.iters, b.iters.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (n_kernels) *n_kernels = nk + 2;
    return TM_OK;
}

int icp_run(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations, float dist_thres,
            float* T16s_out, uint32_t* counts, double* scores, uint32_t* iters, const IcpSplit* split) {
    REQUIRE(s && m, "null handle");
    REQUIRE(n == 0 || (T16s && T16s_out && counts), "tm_icp: null buffer");
    tm_ctx* c = s->ctx;
    TRY(bind(c));
    if (!n) return TM_OK;
    if (max_iterations == 0 && !split) {  // scene.hpp:371: the match is returned unchanged
        memcpy(T16s_out, T16s, (size_t)n * 64);
        if (iters) memset(iters, 0, (size_t)n * 4);
        return tm_score(s, m, T16s, n, nullptr, nullptr, nullptr, 0, dist_thres, 0.f, 0, counts, scores, nullptr);
    }
    const float thres = (2 * dist_thres) * m->dev.resolution;  // scene.hpp:373 + :413
    IcpSplit sp;
    if (split) sp = *split;
    else { sp.pt_end = s->dev.n; sp.n_total = s->dev.n; }
    TRY(c->icp.ensure(n));
    TRY(c->icp.pairs.ensure(icp_pairs_bytes(sp.pt_begin, sp.pt_end, n)));
    TRY(c->icp_d16.ensure((size_t)n * 64));
    TRY(pinned_ensure(c, IcpStage::bytes(n)));
    ModelDev mdev;
    TRY(model_dev_for(c, m, thres, &mdev));  // may build the occupancy mask (synchronises): before any capture
    IcpStage hs(c->pinned, n);
    memcpy(hs.T_in, T16s, (size_t)n * 64);
    static const bool use_graph = [] {
        const char* e = getenv("TM_ICP_GRAPH");  // TM_ICP_GRAPH=0: launch by launch (development knob)
        return e ? atoi(e) != 0 : true;
    }();
    if (sp.comm || !use_graph) {
        TRY(icp_sequence(c, s, m, mdev, n, max_iterations, thres, sp, hs, nullptr));
        CU(cudaGetLastError());
    } else {
        IcpGraphKey key;
        key.scene_pos = s->dev.pos; key.model_vox = m->dev.voxel; key.occ = mdev.occ; key.bufs = c->icp.Tcur.p;
        key.pinned = c->pinned; key.d16 = c->icp_d16.p; key.pairs = c->icp.pairs.p; key.pairs_cap = c->icp.pairs.cap; key.pinned_gen = c->pinned_gen; key.scene_n = s->dev.n; key.k = n;
        key.max_iterations = max_iterations; key.pt_begin = sp.pt_begin; key.pt_end = sp.pt_end; key.emulate = sp.emulate;
        key.n_total = sp.n_total; key.thres = thres;
        IcpGraph& g = c->icp_graph;
        if (!g.exec || !(g.key == key)) {
            g.release();
            const unsigned long long before = g_launch_count.load(std::memory_order_relaxed);
            CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            uint32_t nk = 0;
            int rc = icp_sequence(c, s, m, mdev, n, max_iterations, thres, sp, hs, &nk);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            g_launch_count.store(before, std::memory_order_relaxed);  // captured, not launched
            if (rc) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            if (e != cudaSuccess) return fail(TM_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                g.exec = nullptr;
                return fail(TM_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
            }
            g.key = key;
            g.kernels = nk;
        }
        CU(cudaGraphLaunch(g.exec, c->stream));
        g_launch_count += g.kernels;
    }
    CU(cudaStreamSynchronize(c->stream));
    memcpy(T16s_out, hs.T_out, (size_t)n * 64);
    if (iters) memcpy(iters, hs.iters, (size_t)n * 4);
    for (uint32_t h = 0; h < n; ++h) {
        counts[h] = (uint32_t)hs.sums[(size_t)h * ICP_NSUM];
        if (scores) scores[h] = (double)hs.sums[(size_t)h * ICP_NSUM + 16] / SCORE_SCALE / (double)m->dev.cloud.n;
    }
    return TM_OK;
}

extern "C" {

int tm_icp(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations,
           float dist_thres, float* T16s_out, uint32_t* counts, double* scores, uint32_t* iters) {
    return icp_run(s, m, T16s, n, max_iterations, dist_thres, T16s_out, counts, scores, iters, nullptr);
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

// SURVEY §8e, first option for BASELINE configs[4]: the POSES are sharded, every rank holds the whole scene.
// Rank r refines poses [n*r/world, n*(r+1)/world) exactly as tm_icp would (no collective inside the loop) and
// one all-gather of 80-byte records publishes all n results on every rank.
int tm_icp_pose_sharded(tm_scene* s, tm_model* m, tm_comm* cm, uint32_t rank, uint32_t world, const float* T16s,
                        uint32_t n, uint32_t max_iterations, float dist_thres, float* T16s_out, uint32_t* counts,
                        double* scores, uint32_t* iters) {
    REQUIRE(s && m, "null handle");
    REQUIRE(n == 0 || (T16s && T16s_out && counts), "tm_icp_pose_sharded: null buffer");
    REQUIRE(!cm || cm->ctx == s->ctx, "communicator belongs to another context");
    if (cm) { rank = (uint32_t)cm->rank; world = (uint32_t)cm->world; }
    REQUIRE(world > 0 && rank < world, "tm_icp_pose_sharded: bad rank/world");
    if (!n) return TM_OK;
    const uint32_t b = (uint32_t)((uint64_t)n * rank / world), e = (uint32_t)((uint64_t)n * (rank + 1) / world);
    std::vector<double> sc_local(std::max(1u, e - b));
    if (e > b)
        TRY(icp_run(s, m, T16s + 16 * (size_t)b, e - b, max_iterations, dist_thres, T16s_out + 16 * (size_t)b, counts + b,
                    sc_local.data(), iters ? iters + b : nullptr, nullptr));
    if (scores)
        for (uint32_t h = b; h < e; ++h) scores[h] = sc_local[h - b];
    if (!cm) return TM_OK;  // single process: only [b, e) is written
    // all-gather: fixed slots of `per` records {T[16] f32, count u32, iters u32, score f64}
    tm_ctx* c = s->ctx;
    const uint32_t per = (n + world - 1) / world;
    constexpr size_t REC = 64 + 4 + 4 + 8;
    const size_t slot = (size_t)per * REC;
    // own pinned block: the refinement's cached graph copies to / from c->pinned, which must not be re-allocated
    // between replays
    const size_t need = (size_t)(world + 1) * slot;
    if (need > c->pinned_gather_cap) {
        if (c->pinned_gather) cudaFreeHost(c->pinned_gather);
        c->pinned_gather = nullptr;
        c->pinned_gather_cap = 0;
        CU(cudaMallocHost(&c->pinned_gather, need));
        c->pinned_gather_cap = need;
    }
    TRY(c->icp_pack.ensure(need));
    uint8_t* hsend = static_cast<uint8_t*>(c->pinned_gather);
    memset(hsend, 0, slot);
    for (uint32_t h = b; h < e; ++h) {
        uint8_t* r = hsend + (size_t)(h - b) * REC;
        memcpy(r, T16s_out + 16 * (size_t)h, 64);
        memcpy(r + 64, &counts[h], 4);
        const uint32_t it = iters ? iters[h] : 0u;
        memcpy(r + 68, &it, 4);
        memcpy(r + 72, &sc_local[h - b], 8);
    }
    uint8_t* dsend = c->icp_pack.as<uint8_t>();
    uint8_t* drecv = dsend + slot;
    CU(cudaMemcpyAsync(dsend, hsend, slot, cudaMemcpyHostToDevice, c->stream));
    TRY(comm_allgather_bytes(cm, dsend, drecv, slot, c->stream));
    uint8_t* hrecv = hsend + slot;
    CU(cudaMemcpyAsync(hrecv, drecv, (size_t)world * slot, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint32_t r = 0; r < world; ++r) {
        const uint32_t rb = (uint32_t)((uint64_t)n * r / world), re = (uint32_t)((uint64_t)n * (r + 1) / world);
        for (uint32_t h = rb; h < re; ++h) {
            const uint8_t* rec = hrecv + (size_t)r * slot + (size_t)(h - rb) * REC;
            memcpy(T16s_out + 16 * (size_t)h, rec, 64);
            memcpy(&counts[h], rec + 64, 4);
            if (iters) memcpy(&iters[h], rec + 68, 4);
            if (scores) memcpy(&scores[h], rec + 72, 8);
        }
    }
    return TM_OK;
}

}  // extern "C"
