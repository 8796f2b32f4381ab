// capi_icp.cu — icp_ (include/impl/scene.hpp:369-404): the accumulate / step loop on the device, for host
// pose lists (tm_icp, tm_icp_sharded) and for the resident query's top-k stage (icp_enqueue).
#include "capi_internal.cuh"

static double icp_fix_scale(const tm_model* m, uint32_t n_scene, float thres) {
    // |s'|,|m'| <= r = half bbox diagonal + thres; n * r^2 * 2^bits < 2^62
    double r = (double)m->half_diag + (double)thres + 1e-6;
    double bound = std::max(1.0, (double)std::max(n_scene, 1u) * std::max(r * r, r));
    int bits = (int)std::floor(62.0 - std::log2(bound));
    bits = std::max(8, std::min(40, bits));
    return std::ldexp(1.0, bits);
}

// enqueue the ICP loop for k transforms already in b.Tcur with b.active set
int icp_enqueue(tm_ctx* c, const CloudDev& scene, const tm_model* m, IcpBufs& b, uint32_t k,
                       uint32_t max_iterations, float dist_thres, const IcpSplit* split) {
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

int icp_run(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations,
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
}  // extern "C"
