// capi_internal.cuh — state and helpers shared by the capi_*.cu translation units (the implementation of
// include/tm_b200.h).  Host logic only; every data-parallel step is one of the kernels in k_*.cu.
//   capi_core.cu    contexts, cloud / model / scene upload, the stage calls with host buffers
//   capi_icp.cu     icp_ (tm_icp, tm_icp_sharded, tm_icp_poses) and its enqueue used by the resident query
//   capi_query.cu   the resident query pipeline (tm_query_*)
//   capi_nccl.cu    the NCCL loader, communicators and the best-pose all-reduce
#pragma once
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

using namespace tmk;

// ------------------------------------------------------------------- errors
int tm_fail(int code, const std::string& msg);  // records the thread's last error, returns code
#define fail tm_fail
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
    bool early_levels = true;      // TM_EARLY_LEVELS=0: early_out = 2 of tm_query_run through the one-warp-per-hypothesis
                                   // walker instead of the level-by-level tiled scorer (k_early2.cu); same results
    bool early_merge = true;       // TM_EARLY_MERGE=0: one checkpoint range per launch instead of up to four, no probe of
                                   // checkpoint 1; same results
};
const Knobs& knobs();

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

struct IcpBufs {
    DevBuf Tcur, Tbest, sums_cur, sums_best, iters, active;
    DevBuf pairs;  // surviving (segment, transform) pairs of one pass + their counter (last 8 bytes)
    void release() {
        for (DevBuf* b : {&Tcur, &Tbest, &sums_cur, &sums_best, &iters, &active, &pairs}) b->release();
    }
    uint2* pair_list() const { return pairs.as<uint2>(); }
    uint32_t* pair_count() const { return reinterpret_cast<uint32_t*>(pairs.as<uint8_t>() + pairs.cap - 8); }
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
// one cached CUDA graph of a whole refinement (tm_icp and friends): H2D of the poses, layout change, the
// 2 * (1 + max_iterations) accumulate / step launches, layout change, D2H of the results.  Replayed while the
// key (every pointer and scalar baked into the captured launches) is unchanged.
struct IcpGraphKey {
    const void* scene_pos = nullptr; const void* model_vox = nullptr; const void* occ = nullptr;
    const void* bufs = nullptr; const void* pinned = nullptr; const void* d16 = nullptr; const void* pairs = nullptr;
    size_t pairs_cap = 0;
    uint64_t pinned_gen = 0;
    uint32_t scene_n = 0, k = 0, max_iterations = 0, pt_begin = 0, pt_end = 0, emulate = 0;
    uint64_t n_total = 0;
    float thres = 0.f;
    bool operator==(const IcpGraphKey& o) const {
        return scene_pos == o.scene_pos && model_vox == o.model_vox && occ == o.occ && bufs == o.bufs &&
               pinned == o.pinned && d16 == o.d16 && pairs == o.pairs && pairs_cap == o.pairs_cap && pinned_gen == o.pinned_gen && scene_n == o.scene_n && k == o.k &&
               max_iterations == o.max_iterations && pt_begin == o.pt_begin && pt_end == o.pt_end &&
               emulate == o.emulate && n_total == o.n_total && thres == o.thres;
    }
};
struct IcpGraph {
    IcpGraphKey key;
    cudaGraphExec_t exec = nullptr;
    uint32_t kernels = 0;  // launches one replay stands for (tm_ctx_kernel_launches)
    void release() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
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
    uint64_t pinned_gen = 0;                 // bumped whenever the pinned block is re-allocated (a cached graph that
                                             // copies to / from it must be re-captured even if the address repeats)
    void* pinned_gather = nullptr;           // staging of the pose-sharded all-gather (kept apart from `pinned`)
    size_t pinned_gather_cap = 0;
    int score_bps[2][2] = {{0, 0}, {0, 0}};  // resident CTAs per SM of the scoring kernel [fused][with_score]
    int count_bps[2] = {0, 0};               // the same for the count-only kernel [fused]
    int level_bps[2] = {0, 0};               // ... and for the per-level scorer of the early drop [fused]
    IcpBufs icp;                             // refinement state of tm_icp*, grow-only (no allocation per call)
    DevBuf icp_d16, icp_pack;                // poses in / out (column-major), packed records of the pose-sharded gather
    IcpGraph icp_graph;
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

// 1.5 cell diagonals in model units, rounded up: how far a position can be from the centre of the cell voxel_query
// maps it to (truncation toward zero makes cell 0 two cells wide)
inline float cell_reach_of(const ModelDev& d) {
    const double diag = std::sqrt(1.0 / ((double)d.sx * d.sx) + 1.0 / ((double)d.sy * d.sy) + 1.0 / ((double)d.sz * d.sz));
    return (float)(1.5 * diag * 1.0001);
}
int bind(tm_ctx* c);
// ModelDev for kernels that test against `thres`: the resident description plus, when it pays, the
// block-occupancy mask of that threshold (built on first use, cached per model).  TM_OCC=0 disables.
int model_dev_for(tm_ctx* c, tm_model* m, float thres, ModelDev* out);
int pinned_ensure(tm_ctx* c, size_t bytes);
// `dist > thres` with dist = sqrtf(sq) (scene.hpp:464-465) <=> sq > S, where S is the
// largest float whose correctly rounded square root is <= thres.
float sq_threshold(float thres);
extern "C" void pair_window(const tm_model* m, float min_df, float max_df, float& lower, float& upper);
extern "C" int ball_subsets_dev(tm_ctx* c, const CloudDev& scene, const uint32_t* d_centres, uint32_t n_centres,
                     const uint32_t* active_ranges, float radius, DevBuf& counts, DevBuf& row_tot, DevBuf& row_off,
                     DevBuf* indices, uint64_t* total_out);

// --------------------------------------------------------------------- ICP
typedef struct ncclComm* ncclComm_t;
struct tm_comm {
    tm_ctx* ctx;
    ncclComm_t comm;
    int rank, world;
    DevBuf stage;  // batched best-pose reduce: n keys, then n x (score, pose)
};
int comm_allreduce_sum_i64(tm_comm* cm, void* buf, size_t count, cudaStream_t st);  // capi_nccl.cu
int comm_allreduce_max_u32(tm_comm* cm, void* buf, size_t count, cudaStream_t st);
int comm_allgather_bytes(tm_comm* cm, const void* send, void* recv, size_t bytes, cudaStream_t st);
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
int icp_enqueue(tm_ctx* c, const CloudDev& scene, const tm_model* m, IcpBufs& b, uint32_t k, uint32_t max_iterations,
                float dist_thres, const IcpSplit* split = nullptr);
int icp_run(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations, float dist_thres,
            float* T16s_out, uint32_t* counts, double* scores, uint32_t* iters, const IcpSplit* split);

// ------------------------------------------------------------ resident query
struct QueryOut {  // one contiguous device block, read back in one copy
    // written by the front end (pair features, probe, prefix, shard): kept when tm_query_run follows
    // tm_query_set_pairs directly
    unsigned long long shard[3];  // h_begin, h_end, H
    unsigned long long n_valid;
    uint32_t n_local;
    uint32_t err;
    // reset by every run
    unsigned long long best;
    unsigned long long n_tests;
    double best_score;
    float best_T16[16];  // directly behind best_score: the two travel as one 72-byte record
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
    // early_out = 2, level by level (k_early2.cu)
    DevBuf lvl_idx, lvl_pos, el_n_items, el_item_off, el_items, el_alive, el_corrs, el_cnt, el_minkey, el_irregular, el_ctrl,
        el_list, el_goff, el_hist;
    uint32_t el_items_cap = 0;
    bool levels = false;        // last run used the level scheme (scores[] on demand through the walker)
    uint32_t max_sub = 0;
    IcpBufs icp;
    QueryOut host_out;
    bool ran = false;
    bool lazy = false;          // last run used the count-only scorer: scores[] is filled on demand
    bool scores_valid = false;  // scores[] holds every hypothesis' score
    float run_thres = 0.f, run_sqt = 0.f;
    cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;  // around the scoring kernel
    cudaEvent_t ev_f0 = nullptr, ev_f1 = nullptr;  // around the replicated front end (features, probe, scan, shard)
    // sharding
    bool by_tests = false;       // tm_query_set_balance: equal hypothesis-point tests instead of equal counts
    bool balanced = false;       // bounds[] holds the by-tests split of the current list
    tm_comm* comm = nullptr;     // completes the per-outer subset sizes of the by-tests split (one all-reduce)
    DevBuf bounds, bal_cum, scan_scratch;
    bool pairs_set = false, need_size = false;
    bool front_ready = false;    // the front end's outputs for the current list and shard are on the device (set by
                                 // the sizing pass of tm_query_set_pairs, consumed by the next tm_query_run)
    bool balls_ready = false;    // ... and so are the radius-search counts of the shard's outer samples
    uint32_t sized_rank = 0, sized_world = 1;
    std::vector<uint32_t> opo_host;  // pairs per outer sample (prefix), kept for re-sizing
};
int finalize_best(tm_query* q);  // capi_query.cu
