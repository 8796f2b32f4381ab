// capi_query.cu — the resident query (tm_query_*): the whole recorded-list search enqueued on one stream with
// no host round trip: features + probe -> scan -> shard -> subsets -> hypotheses -> work list -> scoring ->
// argmax (-> top-k -> ICP) -> best-pose export.
#include <cstddef>

#include "capi_internal.cuh"


// (a1-a5) on the device for the whole recorded list: pair filter, feature, key, probe, and the exclusive scan of
// the hit counts (hyp_off).  Replicated on every rank: one thread per pair and one scan, a few tens of
// microseconds for the 2.6e5 pairs of an 8-GPU C2 list (tm_query_frontend_ms reports it).
static int enqueue_front_end(tm_query* q, unsigned long long* n_valid) {
    tm_ctx* c = q->s->ctx;
    float lower, upper;
    pair_window(q->m, q->p.min_diameter_factor, q->p.max_diameter_factor, lower, upper);
    const uint32_t limit = q->p.query_limit ? q->p.query_limit : 200;
    launch_pair_features_probe(c->stream, q->s->dev, q->m->dev, q->outer.as<uint32_t>(), q->pair_outer.as<uint32_t>(),
                               q->pair_j.as<uint32_t>(), q->n_pairs, lower, upper, limit, nullptr, nullptr,
                               q->valid.as<uint8_t>(), q->hit_begin.as<uint32_t>(), q->hit_count.as<uint32_t>(), n_valid);
    if (q->n_pairs == 0) CU(cudaMemsetAsync(q->hyp_off.p, 0, 8, c->stream));
    else
        launch_exclusive_scan_u64_chained(c->stream, q->hit_count.as<uint32_t>(), q->hyp_off.as<unsigned long long>(),
                                          q->n_pairs, q->scan_scratch.as<unsigned long long>(), c->sm_count);
    return TM_OK;
}

// Decide this rank's shard and size everything that depends on it.  The shard is a contiguous range of the
// global hypothesis list: equal hypothesis counts (default) or, with tm_query_set_balance(by_tests), equal
// hypothesis-point tests (balance_bounds_kernel).  Only the outer samples that own hypotheses of the shard are
// sized (radius-search counting pass), so N ranks do not repeat each other's sizing; the by-tests split needs
// the subset size of EVERY outer sample, which each rank measures for its count-based share and one
// ncclAllReduce(max) of n_outer x 4 bytes completes (without a communicator every rank sizes all of them).
static int size_for_shard(tm_query* q) {
    tm_ctx* c = q->s->ctx;
    const uint32_t n_outer = q->n_outer;
    const uint64_t n_pairs = q->n_pairs;
    const std::vector<uint32_t>& opo = q->opo_host;
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
    TRY(q->bounds.ensure(((size_t)q->world + 1) * 8)); TRY(q->bal_cum.ensure(((size_t)n_outer + 1) * 8));
    TRY(q->scan_scratch.ensure(scan_scratch_bytes(n_pairs)));
    QueryOut* out = q->out.as<QueryOut>();
    q->balanced = false;
    std::vector<uint32_t> gh(n_outer + 1, 0), sizes(n_outer + 1, 0);
    if (n_outer) {
        CU(cudaMemsetAsync(out, 0, sizeof(QueryOut), c->stream));
        TRY(enqueue_front_end(q, &out->n_valid));
        // count-based shard and the outer samples it touches (by-tests: unclipped, every outer sample must be sized
        // by the rank whose count-based share holds it)
        const bool by_tests = q->by_tests && q->world > 1;
        launch_shard_range(c->stream, q->hyp_off.as<unsigned long long>(), n_pairs, q->p.hyp_limit, q->rank, q->world,
                           by_tests ? ~0ull : q->cap_hyp, nullptr, out->shard, &out->n_local, &out->err);
        launch_group_hyp_ranges(c->stream, q->hyp_off.as<unsigned long long>(), q->outer_pair_off.as<uint32_t>(), n_outer,
                                out->shard, q->g_hyp.as<uint32_t>());
        const uint32_t* active = (by_tests && !q->comm) ? nullptr : q->g_hyp.as<uint32_t>();
        TRY(ball_subsets_dev(c, q->s->dev, q->outer.as<uint32_t>(), n_outer, active, q->m->dev.diameter, q->ball_counts,
                             q->ball_seg_off, q->sub_off, nullptr, nullptr));
        if (by_tests) {
            if (q->comm) TRY(comm_allreduce_max_u32(q->comm, q->ball_seg_off.p, n_outer, c->stream));
            launch_balance_bounds(c->stream, q->hyp_off.as<unsigned long long>(), n_pairs, q->p.hyp_limit,
                                  q->outer_pair_off.as<uint32_t>(), q->ball_seg_off.as<uint32_t>(), n_outer, q->world,
                                  q->bal_cum.as<unsigned long long>(), q->bounds.as<unsigned long long>());
            CU(cudaMemsetAsync(&out->err, 0, 4, c->stream));
            launch_shard_range(c->stream, q->hyp_off.as<unsigned long long>(), n_pairs, q->p.hyp_limit, q->rank, q->world,
                               q->cap_hyp, q->bounds.as<unsigned long long>(), out->shard, &out->n_local, &out->err);
            launch_group_hyp_ranges(c->stream, q->hyp_off.as<unsigned long long>(), q->outer_pair_off.as<uint32_t>(),
                                    n_outer, out->shard, q->g_hyp.as<uint32_t>());
            q->balanced = true;
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(gh.data(), q->g_hyp.p, (n_outer + 1) * 4ull, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(sizes.data(), q->ball_seg_off.p, n_outer * 4ull, cudaMemcpyDeviceToHost, c->stream));
    } else {
        TRY(q->sub_off.ensure(8));
        CU(cudaMemsetAsync(q->sub_off.p, 0, 8, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    uint64_t total = 0, items = 0, el_items = 0;
    q->max_sub = 0;
    for (uint32_t o = 0; o < n_outer; ++o) {
        const uint64_t nh = gh[o + 1] - gh[o];
        if (!nh) continue;  // not in this rank's shard: its subset row stays empty
        const uint64_t np = sizes[o];
        total += np;
        q->max_sub = (uint32_t)std::max<uint64_t>(q->max_sub, np);
        items += ((np + SCORE_TILE - 1) / SCORE_TILE) * ((nh + SCORE_HCHUNK - 1) / SCORE_HCHUNK + 1);
        el_items += el_items_bound(np, nh);
    }
    (void)opo;
    q->sub_total = total;
    TRY(q->sub_idx.ensure(std::max<uint64_t>(total, 1) * 4));
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
        if (knobs().early_levels) {
            REQUIRE(el_items < (1ull << 31), "too many work items");
            q->el_items_cap = (uint32_t)std::max<uint64_t>(el_items, 1);
            const size_t lg = std::max(n_outer, 1u);
            TRY(q->lvl_idx.ensure(std::max<uint64_t>(total, 1) * 4)); TRY(q->lvl_pos.ensure(std::max<uint64_t>(total, 1) * 4));
            TRY(q->el_n_items.ensure(lg * 4)); TRY(q->el_item_off.ensure((lg + 1) * 4));
            TRY(q->el_list.ensure(cap * 8)); TRY(q->el_goff.ensure((lg + 1) * 8));
            TRY(q->el_hist.ensure(walk_levels_hist_bytes(n_outer, q->max_sub)));
            TRY(q->el_items.ensure((size_t)q->el_items_cap * sizeof(WorkItem)));
            TRY(q->el_alive.ensure(cap)); TRY(q->el_corrs.ensure(cap * 4)); TRY(q->el_cnt.ensure(cap * 4 * EL_MAX_MERGE));
            TRY(q->el_minkey.ensure(cap * 4 * EL_MAX_MERGE)); TRY(q->el_irregular.ensure(cap * 4));
            TRY(q->el_ctrl.ensure((EL_LEVELS + 1) * 4));
        }
    }
    if (q->p.icp_top_k) {
        TRY(q->icp.ensure(q->p.icp_top_k));
        TRY(q->topk_ids.ensure(q->p.icp_top_k * 4ull));
        TRY(q->topk_keys.ensure(topk_scratch_bytes(q->cap_hyp, q->p.icp_top_k)));
        TRY(q->icp_T16.ensure(q->p.icp_top_k * 64ull));
    }
    q->sized_rank = q->rank;
    q->sized_world = q->world;
    q->need_size = false;
    // what the sizing pass left on the device is exactly what the next run would recompute
    q->front_ready = n_outer != 0;
    q->balls_ready = n_outer != 0 && !q->balanced;  // by-tests: the counts belong to the count-based share
    return TM_OK;
}

extern "C" {

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
    if (e == cudaSuccess) e = cudaEventCreate(&q->ev_f0);
    if (e == cudaSuccess) e = cudaEventCreate(&q->ev_f1);
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
    if (q->ev_f0) cudaEventDestroy(q->ev_f0);
    if (q->ev_f1) cudaEventDestroy(q->ev_f1);
    for (DevBuf* b :
         {&q->outer, &q->pair_outer, &q->pair_j, &q->outer_pair_off, &q->ball_counts,
          &q->ball_seg_off, &q->sub_off, &q->sub_idx, &q->sub_idx_walk, &q->valid, &q->hit_begin, &q->hit_count,
          &q->hyp_off, &q->g_hyp, &q->g_of_hyp, &q->T, &q->hyp_valid, &q->hyp_pair, &q->counts,
          &q->scores, &q->dropped, &q->n_items_g, &q->item_off, &q->items, &q->ctrl, &q->out,
          &q->topk_ids, &q->topk_keys, &q->icp_T16, &q->stats, &q->tile_lo, &q->tile_hi, &q->bounds, &q->bal_cum, &q->scan_scratch,
          &q->lvl_idx, &q->lvl_pos, &q->el_n_items, &q->el_item_off, &q->el_items, &q->el_alive, &q->el_corrs, &q->el_cnt,
          &q->el_minkey, &q->el_irregular, &q->el_ctrl, &q->el_list, &q->el_goff, &q->el_hist})
        b->release();
    q->icp.release();
    delete q;
}

int tm_query_set_shard(tm_query* q, uint32_t rank, uint32_t world) {
    REQUIRE(q && world > 0 && rank < world, "tm_query_set_shard: bad rank/world");
    if (q->pairs_set && (rank != q->sized_rank || world != q->sized_world)) q->need_size = true;
    q->rank = rank;
    q->world = world;
    return TM_OK;
}

int tm_query_set_balance(tm_query* q, int by_tests, tm_comm* comm) {
    REQUIRE(q, "tm_query_set_balance: null query");
    REQUIRE(!comm || comm->ctx == q->s->ctx, "communicator belongs to another context");
    if (q->pairs_set && ((by_tests != 0) != q->by_tests || comm != q->comm)) q->need_size = true;
    q->by_tests = by_tests != 0;
    q->comm = comm;
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
    q->opo_host = opo;
    q->pairs_set = true;
    TRY(size_for_shard(q));
    q->ran = false;
    return TM_OK;
}

}  // extern "C"

// export the pose and score of the hypothesis named by out->best (if this shard owns it).  After the count-only
// scorer the score of that one pose is summed here (score_best_kernel); scores[] stays empty until asked for.
int finalize_best(tm_query* q) {
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
static int enqueue_walker(tm_query* q, const ModelDev& md, uint32_t* counts, unsigned long long* scores, uint8_t* dropped,
                          unsigned long long* n_tests, const uint32_t* hyp_list, const uint32_t* n_list);
static int ensure_scores(tm_query* q) {
    if (q->scores_valid || !q->lazy || !q->n_outer) return TM_OK;
    tm_ctx* c = q->s->ctx;
    QueryOut* out = q->out.as<QueryOut>();
    DevBuf& cnt2 = c->scratch[5];
    TRY(cnt2.ensure(q->cap_hyp * 5));
    if (q->levels) {  // partial scores of dropped hypotheses depend on the walk: the walker produces them
        ModelDev md;
        TRY(model_dev_for(c, q->m, q->run_thres, &md));
        CU(cudaMemsetAsync(q->scores.p, 0, q->cap_hyp * 8, c->stream));
        TRY(enqueue_walker(q, md, cnt2.as<uint32_t>(), q->scores.as<unsigned long long>(),
                           cnt2.as<uint8_t>() + q->cap_hyp * 4, nullptr, nullptr, nullptr));
        CU(cudaGetLastError());
        q->scores_valid = true;
        return TM_OK;
    }
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

// project_(early_out = true) with one warp per hypothesis (score_early_drop_kernel): in subset order with tile boxes
// (early_out = 1) or over the evenly sampling walk (early_out = 2; rows permuted by walk_order_rows_kernel).
// hyp_list / n_list: walk only the listed hypotheses.
static int enqueue_walker(tm_query* q, const ModelDev& md, uint32_t* counts, unsigned long long* scores, uint8_t* dropped,
                          unsigned long long* n_tests, const uint32_t* hyp_list, const uint32_t* n_list) {
    tm_ctx* c = q->s->ctx;
    QueryOut* out = q->out.as<QueryOut>();
    EarlyArgs a;
    a.sub_idx = q->sub_idx.as<int32_t>();
    if (q->p.early_out == 2) {  // evenly sampling walk order (see tm_score)
        launch_walk_order_rows(c->stream, q->sub_idx.as<int32_t>(), q->sub_off.as<unsigned long long>(), q->n_outer,
                               q->max_sub, q->sub_idx_walk.as<int32_t>());
        a.sub_idx = q->sub_idx_walk.as<int32_t>();
    } else {
        a.tile_lo = q->tile_lo.as<float4>();
        a.tile_hi = q->tile_hi.as<float4>();
    }
    a.scene = q->s->dev;
    a.model = md;
    a.sub_off = q->sub_off.as<unsigned long long>();
    a.g_of_hyp = q->g_of_hyp.as<uint32_t>();
    a.T = q->T.as<float4>();
    a.n_hyp = (uint32_t)q->cap_hyp;  // grid bound; the kernel clips to n_local (or to the list's length)
    a.n_hyp_dev = n_list ? n_list : &out->n_local;
    a.hyp_list = hyp_list;
    a.n_tests = n_tests;
    a.sq_thres = q->run_sqt;
    a.accept_prob = q->p.accept_prob;
    a.early_out = 1;
    a.counts = counts;
    a.scores = scores;
    a.dropped = dropped;
    a.tested = nullptr;
    launch_score_early_drop(c->stream, a, q->m->fused);
    return TM_OK;
}

// The steps of the level-by-level early drop.  A scoring stage scores M consecutive levels and applies their
// checkpoints; a probe step applies one level's checkpoint from its first reaching element alone (found by walking the
// level in walk order, level_probe_kernel), so that the level is scored afterwards, for the survivors only.
// Default: level 0 is scored for everybody, checkpoint 1 — where most hypotheses that are going to be dropped are
// dropped — is probed, level 1 follows for the survivors, then four levels per stage and the (double-length) last level.  TM_EARLY_MERGE=0: one level per stage, no probe.
struct LevelStep {
    int probe;       // 1: probe the checkpoint of level L0
    int L0, M;       // scoring stage: levels L0 .. L0 + M - 1
    int skip_first;  // scoring stage whose first level's checkpoint was probed before
};
static std::vector<LevelStep> level_steps() {
    std::vector<LevelStep> v;
    if (!knobs().early_merge) {
        for (int L = 0; L < EL_LEVELS; ++L) v.push_back({0, L, 1, 0});
        return v;
    }
    v.push_back({0, 0, 1, 0});
    v.push_back({1, 1, 1, 0});
    v.push_back({0, 1, 1, 1});  // level 1 itself, for the survivors of its checkpoint
    for (int L = 2; L + 4 <= EL_LEVELS - 1; L += 4) v.push_back({0, L, 4, 0});
    v.push_back({0, EL_LEVELS - 1, 1, 0});  // the last level is twice as long as the others: its chunks fill tiles of their own
    return v;
}

// early_out = 2 level by level (k_early2.cu): regroup the subset rows by checkpoint range, then per stage (one or a few
// ranges) one tiled scoring launch over the hypotheses still alive and one checkpoint launch; the few hypotheses with a
// range that reaches nothing are walked one by one at the end.  Counts and drop flags equal the walker's.
static int enqueue_levels(tm_query* q, float thres, float sqt) {
    tm_ctx* c = q->s->ctx;
    QueryOut* out = q->out.as<QueryOut>();
    const tm_model* m = q->m;
    const uint32_t G = q->n_outer;
    const uint32_t cap = (uint32_t)q->cap_hyp;
    const std::vector<LevelStep> steps = level_steps();
    const unsigned long long* sub_off = q->sub_off.as<unsigned long long>();
    launch_group_of_hyp(c->stream, q->g_hyp.as<uint32_t>(), G, q->g_of_hyp.as<uint32_t>());
    launch_walk_levels(c->stream, q->sub_idx.as<int32_t>(), sub_off, G, q->max_sub, q->el_hist.as<uint32_t>(),
                       q->lvl_idx.as<int32_t>(), q->lvl_pos.as<uint32_t>());
    // the list of hypotheses still alive (two buffers, swapped after every stage) and its per-subset ranges: at first
    // every hypothesis of the shard, ranges = g_hyp
    uint32_t* hl[2] = {q->el_list.as<uint32_t>(), q->el_list.as<uint32_t>() + cap};
    uint32_t* goff[2] = {q->el_goff.as<uint32_t>(), q->el_goff.as<uint32_t>() + (G + 1)};
    launch_el_init(c->stream, &out->n_local, cap, q->el_alive.as<uint8_t>(), hl[0], q->el_corrs.as<uint32_t>(),
                   q->el_cnt.as<uint32_t>(), q->el_minkey.as<uint32_t>(), q->dropped.as<uint8_t>(), q->counts.as<uint32_t>());
    CU(cudaMemcpyAsync(goff[0], q->g_hyp.p, (G + 1) * 4ull, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemsetAsync(q->el_ctrl.p, 0, (EL_LEVELS + 1) * 4, c->stream));
    LevelArgs a;
    a.scene = q->s->dev;
    TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &a.model));
    a.lvl_idx = q->lvl_idx.as<int32_t>();
    a.lvl_pos = q->lvl_pos.as<uint32_t>();
    a.items = q->el_items.as<WorkItem>();
    a.n_items = q->el_item_off.as<uint32_t>() + G;
    a.T = q->T.as<float4>();
    a.cap = cap;
    a.lvl_cnt = q->el_cnt.as<uint32_t>();
    a.minkey = q->el_minkey.as<uint32_t>();
    a.sq_thres = sqt;
    if (knobs().score_stats) {
        TRY(q->stats.ensure(64));
        CU(cudaMemsetAsync(q->stats.p, 0, 64, c->stream));
        a.stats = q->stats.as<unsigned long long>();
    }
    EvalArgs e;
    e.g_of_hyp = q->g_of_hyp.as<uint32_t>();
    e.sub_off = sub_off;
    e.alive = q->el_alive.as<uint8_t>();
    e.corrs = q->el_corrs.as<uint32_t>();
    e.cap = cap;
    e.lvl_cnt = a.lvl_cnt;
    e.minkey = a.minkey;
    e.counts = q->counts.as<uint32_t>();
    e.dropped = q->dropped.as<uint8_t>();
    e.irregular = q->el_irregular.as<uint32_t>();
    e.n_irregular = q->el_ctrl.as<uint32_t>() + EL_LEVELS;
    e.n_tests = &out->n_tests;
    e.accept_bound = q->p.accept_prob * (float)m->dev.cloud.n;  // scene.hpp:500
    int& b = c->level_bps[m->fused ? 1 : 0];
    if (!b) b = score_level_max_blocks_per_sm(m->fused);
    const int grid = knobs().score_grid ? knobs().score_grid : c->sm_count * b;
    CU(cudaEventRecord(q->ev_s0, c->stream));
    ProbeArgs pr;
    pr.scene = a.scene;
    pr.model = a.model;
    pr.sub_idx_walk = q->sub_idx_walk.as<int32_t>();
    pr.sub_off = sub_off;
    pr.g_of_hyp = e.g_of_hyp;
    pr.T = a.T;
    pr.sq_thres = sqt;
    pr.minkey = a.minkey;  // slot 0
    bool walk_rows = false;
    int cur = 0, n_score = 0;
    for (size_t st = 0; st < steps.size(); ++st) {
        const LevelStep& s = steps[st];
        a.hl = e.hl = pr.hl = hl[cur];
        e.n_alive = pr.n_alive = goff[cur] + G;
        a.L0 = e.L0 = s.L0;
        e.M = s.M;
        e.probe = s.probe;
        e.skip_first = s.skip_first;
        if (s.probe) {
            if (!walk_rows) {  // the subset rows in walk order (also what the walker of the irregular hypotheses reads)
                launch_walk_order_rows(c->stream, q->sub_idx.as<int32_t>(), sub_off, G, q->max_sub, q->sub_idx_walk.as<int32_t>());
                walk_rows = true;
            }
            pr.level = s.L0;
            launch_level_probe(c->stream, pr, cap, m->fused);
        } else {
            // this stage's work list over the hypotheses still alive
            launch_el_work_count(c->stream, sub_off, goff[cur], G, s.L0, s.M, q->el_n_items.as<uint32_t>());
            launch_exclusive_scan_u32(c->stream, q->el_n_items.as<uint32_t>(), q->el_item_off.as<uint32_t>(), G);
            launch_el_work_fill(c->stream, sub_off, goff[cur], G, s.L0, s.M, q->el_item_off.as<uint32_t>(),
                                q->el_items.as<WorkItem>());
            a.work_counter = q->el_ctrl.as<uint32_t>() + n_score++;
            launch_score_level(c->stream, a, grid, m->fused, s.M);
        }
        launch_el_eval(c->stream, e, cap);
        if (st + 1 < steps.size() && !(st + 1 < steps.size() && steps[st + 1].probe)) {  // drop the dead from the list
            launch_el_alive_count(c->stream, hl[cur], goff[cur], G, e.alive, q->el_n_items.as<uint32_t>());
            launch_exclusive_scan_u32(c->stream, q->el_n_items.as<uint32_t>(), goff[cur ^ 1], G);
            launch_el_alive_fill(c->stream, hl[cur], goff[cur], G, e.alive, goff[cur ^ 1], hl[cur ^ 1]);
            cur ^= 1;
        }
    }
    TRY(enqueue_walker(q, a.model, q->counts.as<uint32_t>(), nullptr, q->dropped.as<uint8_t>(), &out->n_tests,
                       q->el_irregular.as<uint32_t>(), e.n_irregular));
    CU(cudaEventRecord(q->ev_s1, c->stream));
    q->lazy = true;
    q->levels = true;
    return TM_OK;
}

extern "C" {

int tm_query_run(tm_query* q) {
    REQUIRE(q, "null query");
    tm_ctx* c = q->s->ctx;
    TRY(bind(c));
    const tm_model* m = q->m;
    const CloudDev& sc = q->s->dev;
    if (q->need_size) TRY(size_for_shard(q));  // rank / world / balance changed after tm_query_set_pairs
    QueryOut* out = q->out.as<QueryOut>();
    // a run that directly follows tm_query_set_pairs finds the front end's outputs (hit counts, prefix, shard,
    // per-outer ranges) and, for count-based shards, the radius-search counts already on the device
    const bool reuse_front = q->front_ready, reuse_balls = q->front_ready && q->balls_ready;
    q->front_ready = q->balls_ready = false;
    const size_t run_part = offsetof(QueryOut, best);
    if (reuse_front) CU(cudaMemsetAsync(reinterpret_cast<uint8_t*>(out) + run_part, 0, sizeof(QueryOut) - run_part, c->stream));
    else CU(cudaMemsetAsync(out, 0, sizeof(QueryOut), c->stream));
    CU(cudaMemsetAsync(q->counts.p, 0, q->cap_hyp * 4, c->stream));
    CU(cudaMemsetAsync(q->scores.p, 0, q->cap_hyp * 8, c->stream));
    // (a1-a5) pair filter, feature, key, probe
    CU(cudaEventRecord(q->ev_f0, c->stream));
    if (!reuse_front) {
        TRY(enqueue_front_end(q, &out->n_valid));
        launch_shard_range(c->stream, q->hyp_off.as<unsigned long long>(), q->n_pairs, q->p.hyp_limit,
                           q->rank, q->world, q->cap_hyp, q->balanced ? q->bounds.as<unsigned long long>() : nullptr,
                           out->shard, &out->n_local, &out->err);
        launch_group_hyp_ranges(c->stream, q->hyp_off.as<unsigned long long>(),
                                q->outer_pair_off.as<uint32_t>(), q->n_outer, out->shard,
                                q->g_hyp.as<uint32_t>());
    }
    CU(cudaEventRecord(q->ev_f1, c->stream));
    // (a8) radius subsets, only of the outer samples that own hypotheses of this rank's shard
    // (g_hyp); the others get empty rows, so N ranks do not repeat each other's searches
    if (q->n_outer) {
        const float r2 = m->dev.diameter * m->dev.diameter;
        const uint32_t n_seg = (sc.n + BALL_SEG - 1) / BALL_SEG;
        if (!reuse_balls) {
            launch_ball_count(c->stream, sc, q->outer.as<uint32_t>(), q->n_outer, q->g_hyp.as<uint32_t>(), r2, n_seg,
                              q->ball_counts.as<uint32_t>());
            launch_ball_seg_scan(c->stream, q->ball_counts.as<uint32_t>(), q->n_outer, n_seg,
                                 q->ball_seg_off.as<uint32_t>());
            launch_exclusive_scan_u64(c->stream, q->ball_seg_off.as<uint32_t>(), q->sub_off.as<unsigned long long>(),
                                      q->n_outer);
        }
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
    q->levels = false;
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
            a.thres = thres;
            a.cell_reach = cell_reach_of(a.model);
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
        } else if (q->p.early_out == 2 && knobs().early_levels) {
            TRY(enqueue_levels(q, thres, sqt));
        } else {
            launch_group_of_hyp(c->stream, q->g_hyp.as<uint32_t>(), q->n_outer,
                                q->g_of_hyp.as<uint32_t>());
            ModelDev md;
            TRY(model_dev_for(c, const_cast<tm_model*>(m), thres, &md));
            if (q->p.early_out == 1)
                launch_subset_tile_boxes(c->stream, sc, q->sub_idx.as<int32_t>(), q->sub_off.as<unsigned long long>(),
                                         q->n_outer, q->max_sub, q->tile_lo.as<float4>(), q->tile_hi.as<float4>());
            CU(cudaEventRecord(q->ev_s0, c->stream));
            TRY(enqueue_walker(q, md, q->counts.as<uint32_t>(), q->scores.as<unsigned long long>(),
                               q->dropped.as<uint8_t>(), &out->n_tests, nullptr, nullptr));
            CU(cudaEventRecord(q->ev_s1, c->stream));
        }
    }
    // a hypothesis the early drop abandoned cannot win on its partial count (scene.hpp:330: it returned fewer
    // correspondences than the acceptance bound)
    launch_argmax(c->stream, q->counts.as<uint32_t>(), q->hyp_valid.as<uint8_t>(),
                  q->p.early_out ? q->dropped.as<uint8_t>() : nullptr, &out->n_local, out->shard, &out->best,
                  c->sm_count * 2);
    // (a12) ICP of the local top-k
    if (q->p.icp_top_k) {  // max_icp_iterations == 0: the top k unrefined, counted at the ICP threshold (scene.hpp:371)
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
        if (q->levels)
            fprintf(stderr, "[tm stats] level scheme: (tile,hyp) pairs %llu  of live hypotheses %llu (%.1f%%)  survive cull %llu "
                    "(%.1f%% of live; 64-point halves evaluated %llu = %.1f%%)  with a reaching element %llu\n", st[0], st[1],
                    st[0] ? 100.0 * st[1] / st[0] : 0.0, st[2], st[1] ? 100.0 * st[2] / st[1] : 0.0, st[4],
                    st[1] ? 50.0 * st[4] / st[1] : 0.0, st[3]);
        else
        fprintf(stderr, "[tm stats] (warp-tile,hyp) pairs %llu  survive cull %llu (%.1f%%; 64-point halves evaluated %.1f%%)  "
                "with inliers %llu (%.1f%%)  all-inlier tiles %llu  >=90%% %llu  inliers %llu\n",
                st[0], st[1], st[0] ? 100.0 * st[1] / st[0] : 0.0, st[0] ? 50.0 * st[6] / st[0] : 0.0, st[2],
                st[0] ? 100.0 * st[2] / st[0] : 0.0, st[3], st[4], st[5]);
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
int tm_query_frontend_ms(tm_query* q, float* ms) {
    REQUIRE(q && ms && q->ran, "tm_query_frontend_ms: null/unrun query");
    TRY(bind(q->s->ctx));
    CU(cudaEventSynchronize(q->ev_f1));
    CU(cudaEventElapsedTime(ms, q->ev_f0, q->ev_f1));
    return TM_OK;
}
int tm_query_early_walked(tm_query* q, uint32_t* n) {
    REQUIRE(q && n && q->ran, "tm_query_early_walked: null/unrun query");
    TRY(bind(q->s->ctx));
    *n = 0;
    if (!q->levels) return TM_OK;
    CU(cudaMemcpyAsync(n, q->el_ctrl.as<uint32_t>() + EL_LEVELS, 4, cudaMemcpyDeviceToHost, q->s->ctx->stream));
    CU(cudaStreamSynchronize(q->s->ctx->stream));
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
    REQUIRE(k, "query has no ICP stage");
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

}  // extern "C"
