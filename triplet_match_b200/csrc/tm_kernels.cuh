// tm_kernels.cuh — launch wrappers shared between the kernel files and capi_*.cu.
#pragma once
#include <atomic>

#include "tm_device.cuh"

namespace tmk {

// kernels launched by this process (every launch_* wrapper launches exactly one)
extern std::atomic<unsigned long long> g_launch_count;  // launches of this library's kernels (any host thread)

// ---- tuning constants ------------------------------------------------------
constexpr int SCORE_THREADS = 256;
#ifndef TM_SCORE_P
#define TM_SCORE_P 4
#endif
constexpr int SCORE_P = TM_SCORE_P;                     // scene points per thread
#ifndef TM_SCORE_MIN_BLOCKS
#define TM_SCORE_MIN_BLOCKS 4
#endif
constexpr int SCORE_MIN_BLOCKS = TM_SCORE_MIN_BLOCKS;   // __launch_bounds__ occupancy target
constexpr int SCORE_TILE = 32 * SCORE_P;                // points per (warp) work item
constexpr int SCORE_HSTAGE = 128;                       // hypotheses staged in smem at a time
constexpr int SCORE_HCHUNK = 2048;                      // hypotheses per work item
constexpr uint32_t BALL_SEG = 1024;                     // scene points per warp segment
constexpr uint32_t CORR_SEG = 1024;
constexpr int ICP_P = 4;
constexpr int ICP_NSUM = 17;

struct WorkItem {
    unsigned long long sub_begin;  // first subset position (or scene index if no index list)
    uint32_t npts;
    uint32_t hyp_begin, hyp_end;   // local hypothesis range
    uint32_t pad;
};

struct ScoreArgs {
    CloudDev scene;
    ModelDev model;
    const int32_t* sub_idx;  // null => identity
    const WorkItem* items;
    const uint32_t* n_items;
    uint32_t* work_counter;
    const float4* T;
    uint32_t* counts;
    unsigned long long* scores;
    float sq_thres;
    float thres = 0.f;       // the distance threshold itself (sqrtf(sq_thres) rounded up) and
    float cell_reach = 0.f;  // 1.5 cell diagonals in model units: the sphere cull of score_count_x2_kernel
    unsigned long long* stats;  // optional debug counters: [0] (warp,hyp) pairs, [1] survivors,
                                // [2] warp-tiles with an inlier, [3] all-inlier tiles, [4] >= 90 %, [5] inliers,
                                // [6] 64-point halves evaluated
};

struct EarlyArgs {
    CloudDev scene;
    ModelDev model;
    const int32_t* sub_idx;
    const unsigned long long* sub_off;
    const uint32_t* g_of_hyp;
    const float4* T;
    uint32_t n_hyp;              // upper bound (grid size)
    const uint32_t* n_hyp_dev;   // optional device-side count (<= n_hyp)
    unsigned long long* n_tests; // optional: sum of positions actually tried
    float sq_thres;
    float accept_prob;
    int early_out;
    uint32_t* counts;
    unsigned long long* scores;
    uint8_t* dropped;
    uint32_t* tested;
    // optional bounding boxes of every 32 consecutive subset positions (subset_tile_boxes):
    // tile t of group g is entry sub_off[g]/32 + g + t
    const float4* tile_lo = nullptr;
    const float4* tile_hi = nullptr;
    const uint32_t* hyp_list = nullptr;  // optional: warp w walks hypothesis hyp_list[w] (n_hyp_dev entries)
};

// k_early2.cu: the early drop over the evenly sampling walk, level by level
constexpr int EL_LEVELS = 19;  // walk-position ranges between the 18 checkpoints (scene.hpp:422-426); level_begin()
                               // in tm_device.cuh is written for this value
#ifndef TM_EL_HCHUNK
#define TM_EL_HCHUNK 256
#endif
constexpr int EL_HCHUNK = TM_EL_HCHUNK;  // hypotheses per work item: a level holds 1/20 of a subset's tiles, so the
                                         // items are made shorter than the full scorer's to keep the tail small
constexpr int EL_MAX_MERGE = 4;  // levels scored by one launch ("stage")
struct LevelArgs {
    CloudDev scene;
    ModelDev model;
    const int32_t* lvl_idx;    // subset rows regrouped by level (walk_levels_kernel): scene index ...
    const uint32_t* lvl_pos;   // ... and walk position of every element
    const WorkItem* items;     // this stage's work list; hypothesis ranges are positions in hl[]
    const uint32_t* n_items;   // (device) how many
    const uint32_t* hl;        // hypotheses still alive, grouped by subset, order kept
    int L0;                    // first level of the stage
    uint32_t* work_counter;    // this stage's
    const float4* T;
    uint32_t cap;              // hypothesis capacity = stride of the per-level accumulators
    uint32_t* lvl_cnt;         // [EL_MAX_MERGE][cap] inliers per level of the stage and hypothesis
    uint32_t* minkey;          // [EL_MAX_MERGE][cap] min over reaching elements of (walk position << 1) | !inlier
    float sq_thres;
    unsigned long long* stats = nullptr;  // optional debug counters: [0] (tile, hyp) pairs, [1] of live hypotheses,
                                          // [2] surviving the cull, [3] with a reaching element, [4] 64-point
                                          // halves evaluated
};
struct ProbeArgs {  // level_probe_kernel (k_score.cu)
    CloudDev scene;
    ModelDev model;
    const int32_t* sub_idx_walk;  // subset rows in walk order (walk_order_rows_kernel)
    const unsigned long long* sub_off;
    const uint32_t* g_of_hyp;
    const float4* T;
    const uint32_t* hl;           // hypotheses still alive ...
    const uint32_t* n_alive;      // (device) ... and how many
    int level;
    float sq_thres;
    uint32_t* minkey;             // [cap]: slot 0 of the per-level accumulators
};
struct EvalArgs {
    const uint32_t* hl;        // hypotheses still alive before this stage's checkpoints ...
    const uint32_t* n_alive;   // (device) ... and how many
    const uint32_t* g_of_hyp;
    const unsigned long long* sub_off;
    uint8_t* alive;
    uint32_t* corrs;           // inliers of the levels before the current stage
    uint32_t cap;
    uint32_t* lvl_cnt;
    uint32_t* minkey;
    uint32_t* counts;
    uint8_t* dropped;
    uint32_t* irregular;       // hypotheses to be walked one by one (a level without a reaching element)
    uint32_t* n_irregular;
    unsigned long long* n_tests;
    float accept_bound;
    int L0, M;
    int probe = 0;       // the checkpoint of level L0 from a probed key only: the level is scored later, nothing is added
    int skip_first = 0;  // level L0's checkpoint was already applied from a probe: only its count is added
};

struct IcpState {
    float4* Tcur;
    float4* Tbest;
    long long* sums_cur;
    long long* sums_best;
    uint32_t* iters;
    uint32_t* active;
};

// k_util.cu
void launch_pack_cloud(cudaStream_t st, const float* pos, const float* nrm, const float* tgt,
                       uint32_t stride, uint32_t n, const uint8_t* flags, int model_mode,
                       float4* opos, float4* onrm, float4* otgt);
void launch_set_mask(cudaStream_t st, float4* pos, uint32_t n, const uint8_t* mask);
void launch_exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint64_t n);
void launch_exclusive_scan_u64(cudaStream_t st, const uint32_t* in, unsigned long long* out,
                               uint64_t n);
size_t scan_scratch_bytes(uint64_t n);
void launch_exclusive_scan_u64_chained(cudaStream_t st, const uint32_t* in, unsigned long long* out, uint64_t n,
                                       unsigned long long* scratch, int max_ctas);
void launch_exclusive_scan_u32_chained(cudaStream_t st, const uint32_t* in, uint32_t* out, uint64_t n,
                                       unsigned long long* scratch, int max_ctas);
void launch_seg_bbox(cudaStream_t st, const float4* pos, uint32_t n, float4* lo, float4* hi);
void launch_ball_count(cudaStream_t st, const CloudDev& scene, const uint32_t* centres, uint32_t n_centres,
                       const uint32_t* active_ranges, float r2, uint32_t n_seg, uint32_t* counts);
void launch_ball_seg_scan(cudaStream_t st, uint32_t* counts, uint32_t n_centres, uint32_t n_seg,
                          uint32_t* row_total);
void launch_ball_fill(cudaStream_t st, const CloudDev& scene, const uint32_t* centres, uint32_t n_centres,
                      const uint32_t* active_ranges, float r2, uint32_t n_seg, const uint32_t* seg_local_off,
                      const unsigned long long* row_off, int32_t* indices);
// block_scratch == nullptr: brute force; else the pruned two-pass fill (voxel_fill_scratch_bytes floats)
void launch_voxel_fill(cudaStream_t st, const float4* mpos, uint32_t n, int ex, int ey, int ez,
                       float sx, float sy, float sz, float tx, float ty, float tz, uint32_t* voxel,
                       float* block_scratch);
size_t voxel_fill_scratch_bytes(int ex, int ey, int ez);
void launch_l2_gather(cudaStream_t st, const float4* buf, uint32_t n_cells_mask, uint32_t iters, float* out, int grid);
void launch_occupancy(cudaStream_t st, const uint32_t* voxel, const float4* mpos, int ex, int ey, int ez, float sx,
                      float sy, float sz, float tx, float ty, float tz, float reach2, int obx, int oby, uint32_t* occ);
void launch_fuse_grid(cudaStream_t st, const uint32_t* voxel, size_t total, const float4* mpos, float4* vcell);
void launch_model_ref(cudaStream_t st, const float4* mpos, const float4* mnrm, const float4* mtgt, uint32_t n,
                      float4* mref);
void launch_traits_project(cudaStream_t st, int kind, float4 r0, float4 r1, float4 r2, float radius,
                           float threshold, const float* xyz, uint64_t n, float* uvw, uint8_t* ok);
void launch_flush(cudaStream_t st, float4* buf, size_t n, float v);

// k_pairs.cu
void launch_pair_features_probe(cudaStream_t st, const CloudDev& scene, const ModelDev& model,
                                const uint32_t* outer, const uint32_t* pair_i,
                                const uint32_t* pair_j, uint64_t n, float lower, float upper,
                                uint32_t limit, float* feats, uint4* keys, uint8_t* valid,
                                uint32_t* hit_begin, uint32_t* hit_count,
                                unsigned long long* n_valid);
void launch_probe(cudaStream_t st, const ModelDev& model, const uint4* keys, const uint8_t* valid,
                  uint64_t n, uint32_t limit, uint32_t* hit_begin, uint32_t* hit_count);
void launch_gather_hits(cudaStream_t st, const ModelDev& model, const uint32_t* hit_begin,
                        const unsigned long long* offsets, uint64_t n, uint2* out);
void launch_hypotheses(cudaStream_t st, const CloudDev& scene, const ModelDev& model,
                       const uint32_t* outer, const uint32_t* pair_i, const uint32_t* pair_j,
                       uint64_t n_pairs, const unsigned long long* hyp_off,
                       const uint32_t* hit_begin, const uint2* hits, int force_up,
                       const unsigned long long* shard, float4* T, uint8_t* hyp_valid,
                       uint32_t* hyp_pair);
void launch_rows_from_colmajor(cudaStream_t st, const float* T16, uint64_t n, float4* rows);
void launch_colmajor_from_rows(cudaStream_t st, const float4* rows, uint64_t n, float* T16);

// k_score.cu
void launch_score_full(cudaStream_t st, const ScoreArgs& a, int grid, bool fused, bool with_score);
int score_full_max_blocks_per_sm(bool fused, bool with_score);
// k_score2.cu: count-only bulk scorer on the packed FP32 pipe (exact), lazy score of the selected pose
void launch_score_count_x2(cudaStream_t st, const ScoreArgs& a, int grid, bool fused);
int score_count_x2_max_blocks_per_sm(bool fused);
void launch_score_best(cudaStream_t st, const CloudDev& scene, const ModelDev& m, const int32_t* sub_idx,
                       const unsigned long long* sub_off, const uint32_t* g_hyp, uint32_t n_groups, const float4* T,
                       const unsigned long long* best_key, const unsigned long long* shard, float sq_thres,
                       unsigned long long* acc, bool fused);
void launch_work_count(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* g_hyp,
                       uint32_t n_groups, uint32_t* n_items_g, unsigned long long* n_tests);
void launch_work_fill(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* g_hyp,
                      uint32_t n_groups, const uint32_t* item_off, WorkItem* items);
void launch_walk_order_rows(cudaStream_t st, const int32_t* in, const unsigned long long* sub_off, uint32_t n_groups,
                            uint32_t max_sub, int32_t* out);
void launch_score_early_drop(cudaStream_t st, const EarlyArgs& a, bool fused);
void launch_level_probe(cudaStream_t st, const ProbeArgs& a, uint32_t n_bound, bool fused);
void launch_subset_tile_boxes(cudaStream_t st, const CloudDev& scene, const int32_t* sub_idx,
                              const unsigned long long* sub_off, uint32_t n_groups, uint32_t max_sub,
                              float4* tile_lo, float4* tile_hi);
// k_early2.cu
size_t walk_levels_hist_bytes(uint32_t n_groups, uint32_t max_sub);
void launch_walk_levels(cudaStream_t st, const int32_t* sub_idx, const unsigned long long* sub_off, uint32_t n_groups,
                        uint32_t max_sub, uint32_t* hist, int32_t* lvl_idx, uint32_t* lvl_pos);
void launch_el_work_count(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* goff, uint32_t n_groups,
                          int L0, int M, uint32_t* n_items);
void launch_el_work_fill(cudaStream_t st, const unsigned long long* sub_off, const uint32_t* goff, uint32_t n_groups,
                         int L0, int M, const uint32_t* item_off, WorkItem* items);
void launch_el_alive_count(cudaStream_t st, const uint32_t* hl, const uint32_t* goff, uint32_t n_groups, const uint8_t* alive,
                           uint32_t* cnt);
void launch_el_alive_fill(cudaStream_t st, const uint32_t* hl, const uint32_t* goff, uint32_t n_groups, const uint8_t* alive,
                          const uint32_t* goff_new, uint32_t* hl_new);
uint64_t el_items_bound(uint64_t n_points, uint64_t n_hyp);
void launch_score_level(cudaStream_t st, const LevelArgs& a, int grid, bool fused, int M);
int score_level_max_blocks_per_sm(bool fused);
void launch_el_eval(cudaStream_t st, const EvalArgs& a, uint32_t n_hyp_bound);
void launch_el_init(cudaStream_t st, const uint32_t* n_local, uint32_t cap, uint8_t* alive, uint32_t* hl, uint32_t* corrs,
                    uint32_t* lvl_cnt, uint32_t* minkey, uint8_t* dropped, uint32_t* counts);
void launch_argmax(cudaStream_t st, const uint32_t* counts, const uint8_t* valid, const uint8_t* excluded,
                   const uint32_t* n_local, const unsigned long long* h_begin,
                   unsigned long long* best, int grid);

// k_icp.cu
void launch_icp_accumulate(cudaStream_t st, const CloudDev& scene, const ModelDev& model,
                           const float4* T, const uint32_t* active, uint32_t n_hyp,
                           uint32_t pt_begin, uint32_t pt_end, float sq_thres, float cx, float cy,
                           float cz, double fix_scale, long long* sums, uint2* pairs, uint32_t* n_pairs, int grid,
                           bool fused);
size_t icp_pairs_bytes(uint32_t pt_begin, uint32_t pt_end, uint32_t n_hyp);
void launch_icp_step(cudaStream_t stream, const IcpState& st, uint32_t n_hyp, int first,
                     uint32_t max_iterations, double inv_scale, float cx, float cy, float cz);
void launch_corr_count(cudaStream_t st, const CloudDev& scene, const ModelDev& model, const float4* Trows, uint32_t n_T,
                       float sq_thres, uint32_t n_seg, uint32_t* counts, unsigned long long* score, bool fused);
void launch_corr_fill(cudaStream_t st, const CloudDev& scene, const ModelDev& model, const float4* Trows, uint32_t n_T,
                      float sq_thres, uint32_t n_seg, const uint32_t* seg_off, uint32_t* scene_corrs,
                      uint32_t* model_corrs, bool fused);

// k_uvicp.cu (the opencl/icp.cl path, a15)
void launch_uvicp_projection(cudaStream_t st, int projector, const float4* pnts, int n, const float4* image,
                             const int img_size[2], const int img_margin[2], const float* mat_align,
                             const float* mat_uvw, const float* mat_proj, const float* mat_norm,
                             float max_corr_dist, float4* out_positions, int* model_indices, int* scene_indices,
                             unsigned int* n_corr);
int uvicp_correlation_blocks(int n);
void launch_uvicp_correlation(cudaStream_t st, const float4* scene, const float4* model, const int* indices_scene,
                              const int* indices_model, int n, const float* centroid_scene,
                              const float* centroid_model, float* records, double* partials, double* cov9);

// k_sort.cu (Z-curve ordering of a cloud)
void launch_morton_codes(cudaStream_t st, const float4* pos, uint32_t n, const float lo[3], const float inv[3],
                         uint32_t* codes, uint32_t* idx);
uint32_t radix_blocks(uint32_t n);
void launch_radix_sort_pairs(cudaStream_t st, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                             uint32_t n, uint32_t* hist, uint32_t* offsets, unsigned long long* scan_scratch = nullptr,
                             int max_ctas = 1);
void launch_gather_cloud(cudaStream_t st, const float4* pos, const float4* nrm, const float4* tgt, const uint32_t* perm,
                         uint32_t n, float4* opos, float4* onrm, float4* otgt);

// k_model.cu (model::init pair enumeration)
void launch_model_pair_bounds(cudaStream_t st, const float* pos3, const float* tgt3, uint32_t T, float lower, float upper,
                              uint32_t* bounds, unsigned long long* count, int grid);
void launch_model_pair_keys(cudaStream_t st, const float* pos3, const float* tgt3, uint32_t T, float lower, float upper,
                            float fmn0, float fmx0, uint32_t steps, float angle_step, unsigned long long* keys, int grid);

// k_knn.cu (pre-processing: exact k-NN, principal-curvature tangent masks)
void launch_knn(cudaStream_t st, const CloudDev& cloud, const uint32_t* query, uint32_t n_query, uint32_t k,
                int32_t* out_idx, float* out_d2);
void launch_curvature(cudaStream_t st, const CloudDev& cloud, const uint32_t* query, uint32_t n_query, uint32_t k,
                      const int32_t* nbr, float* pc_min, float* pc_max, float* cov_out);
void launch_tangent_candidates(cudaStream_t st, const float4* tgt, uint32_t n, uint32_t* flags);
void launch_compact(cudaStream_t st, const uint32_t* flags, const uint32_t* offsets, uint32_t n, uint32_t* out);
void launch_tangent_mask(cudaStream_t st, float4* pos, uint32_t n, const uint32_t* cand, uint32_t n_cand,
                         const float* pc_min, const float* pc_max, float ratio, uint8_t* mask, int apply);

// k_query.cu (device-side glue of the resident query)
void launch_shard_range(cudaStream_t st, const unsigned long long* hyp_off, uint64_t n_pairs,
                        unsigned long long hyp_limit, uint32_t rank, uint32_t world,
                        unsigned long long capacity, const unsigned long long* bounds, unsigned long long* shard,
                        uint32_t* n_local, uint32_t* err);
void launch_balance_bounds(cudaStream_t st, const unsigned long long* hyp_off, uint64_t n_pairs,
                           unsigned long long hyp_limit, const uint32_t* outer_pair_off, const uint32_t* sizes,
                           uint32_t n_outer, uint32_t world, unsigned long long* cum, unsigned long long* bounds);
void launch_group_hyp_ranges(cudaStream_t st, const unsigned long long* hyp_off,
                             const uint32_t* outer_pair_off, uint32_t n_outer,
                             const unsigned long long* shard, uint32_t* g_hyp);
void launch_group_of_hyp(cudaStream_t st, const uint32_t* g_hyp, uint32_t n_groups,
                         uint32_t* g_of_hyp);
size_t topk_scratch_bytes(uint64_t capacity, uint32_t k);
void launch_select_topk(cudaStream_t st, const uint32_t* counts, const uint8_t* valid, const uint8_t* excluded,
                        const uint32_t* n_local, uint64_t capacity, uint32_t k, uint32_t* topk_ids,
                        unsigned long long* scratch_keys);
void launch_gather_rows(cudaStream_t st, const float4* T, const uint32_t* ids, uint32_t k,
                        float4* out, uint32_t* active);
void launch_finalize_best(cudaStream_t st, const unsigned long long* best,
                          const unsigned long long* shard, const float4* T,
                          const unsigned long long* scores, const unsigned long long* lazy_acc, uint32_t model_n,
                          float* best_T16, double* best_score);

}  // namespace tmk
