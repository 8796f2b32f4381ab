/* tm_b200.h — C-ABI of the B200-native triplet_match search path.
 *
 * The reference (richard-vock/triplet_match) has no FFI: its boundary is the C++
 * class API in include/model, include/scene, include/feature, include/discretize
 * and the *_traits headers.  This header is what the new model::impl /
 * scene::impl (include/triplet_match/, triplet_match_b200/host/) bind instead of
 * the reference's CPU loops; every entry point cites the reference code it
 * replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers unless a
 * parameter is documented as resident; opaque handles; every function returns
 * a tm_status (0 = ok) and never throws; tm_last_error() gives the message of
 * the calling thread's last failure.  A context owns one device and one
 * stream; a handle is used by one host thread at a time.  Transforms are
 * column-major float[16] (Eigen mat4f_t).  Clouds are passed as strided views
 * so that both packed xyz arrays (stride 3) and pcl::PointSurfel AoS records
 * (stride 12 floats; tangent = data_c[1..3], include/common:62-70) bind
 * without a copy on the host.  There is no CPU fallback: without a CUDA device
 * tm_ctx_create fails with TM_ERR_CUDA.
 */
#ifndef TM_B200_H
#define TM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tm_status {
    TM_OK = 0,
    TM_ERR_INVALID = 1,     /* bad argument (null, size, unsupported matrix) */
    TM_ERR_CUDA = 2,        /* CUDA runtime failure (message has the cudaError) */
    TM_ERR_CAPACITY = 3,    /* caller buffer / configured capacity too small  */
    TM_ERR_UNINITIALIZED = 4, /* model not initialised (include/impl/model.hpp:171-173) */
    TM_ERR_NCCL = 5
} tm_status;

typedef struct tm_ctx tm_ctx;
typedef struct tm_model tm_model;
typedef struct tm_scene tm_scene;
typedef struct tm_query tm_query;
typedef struct tm_comm tm_comm;

/* strided cloud view: point i has pos[i*stride + 0..2] etc. */
typedef struct tm_cloud_view {
    const float* pos;
    const float* nrm;
    const float* tgt;
    uint32_t stride; /* in floats; 3 = packed, 12 = pcl::PointSurfel */
    uint32_t n;
} tm_cloud_view;

/* What model::init (include/impl/model.hpp:16-167) produced on the host. */
typedef struct tm_model_desc {
    const uint32_t* voxel; /* voxel_data_, lin = k*ex*ey + j*ex + i (model.hpp:85) */
    int32_t extents[3];    /* extents_ (model.hpp:50) */
    float to_voxel[16];    /* to_voxel_ column-major (model.hpp:56-61); must be diag+translation */
    float resolution;      /* cloud()->resolution() (pointcloud.hpp:66-82) */
    float diameter;        /* diameter_ (model.hpp:39) */
    /* hash_map_t (include/model:25-26) flattened: unique keys, CSR offsets, (i,j)
     * values in equal_range order, at most query_limit kept per key */
    const uint32_t* keys;    /* n_keys x 4 */
    const uint32_t* offsets; /* n_keys + 1 */
    const uint32_t* pairs;   /* offsets[n_keys] x 2 */
    uint32_t n_keys;
    float feat_min[4], feat_max[4]; /* feat_bounds_ (model.hpp:122) */
    float distance_step_count;      /* discretization_params (include/discretize:8-12) */
    float angle_step;
} tm_model_desc;

/* ---- context ------------------------------------------------------------ */
const char* tm_last_error(void);
const char* tm_version(void);
int tm_ctx_create(int device, tm_ctx** out);
void tm_ctx_destroy(tm_ctx* ctx);
int tm_ctx_sync(tm_ctx* ctx);
void* tm_ctx_stream(tm_ctx* ctx);              /* cudaStream_t */
int tm_ctx_sm_count(tm_ctx* ctx);
int tm_timer_start(tm_ctx* ctx);               /* CUDA event on the context stream */
int tm_timer_stop(tm_ctx* ctx, float* ms);     /* records, synchronises, elapsed ms */
int tm_ctx_flush_l2(tm_ctx* ctx);              /* overwrites a 256 MiB scratch buffer */
uint64_t tm_ctx_kernel_launches(tm_ctx* ctx);
/* Micro-benchmark for the scorer's roofline: random 16-byte cell gathers (one 32-byte sector each)
 * over a working set of `working_set_bytes` (rounded down to a power of two; keep it below L2),
 * 8 loads in flight per lane.  Returns sector traffic in GB/s (32 B per load). */
int tm_ctx_measure_l2_gather(tm_ctx* ctx, uint64_t working_set_bytes, double* gb_per_s);  /* kernels launched so far on this context */

/* ---- resident model / scene -------------------------------------------- */
/* replaces the CPU-resident state probed by model::query / model::voxel_query
 * (include/impl/model.hpp:169-192) */
int tm_model_upload(tm_ctx* ctx, const tm_cloud_view* cloud, const tm_model_desc* desc,
                    tm_model** out);
void tm_model_destroy(tm_model* m);
/* GPU grid fill: exact 1-NN of every voxel centre (include/impl/model.hpp:81-94).
 * voxel_out: ex*ey*ez u32. */
int tm_voxel_fill(tm_ctx* ctx, const tm_cloud_view* cloud, const int32_t extents[3],
                  const float to_voxel[16], uint32_t* voxel_out);

/* scene cloud + tangent_mask_ (include/impl/scene.hpp:46-58); mask_ starts 0 */
int tm_scene_upload(tm_ctx* ctx, const tm_cloud_view* cloud, const uint8_t* tangent_mask,
                    tm_scene** out);
/* Same, but the device copy is put into Z-curve (Morton) order first: 30-bit codes over the cloud's
 * bounding box, stable radix sort on the device.  to_user[d] (scene n entries, may be NULL) = index in
 * the caller's cloud of device point d.  EVERY index exchanged with this scene afterwards (outer
 * samples, pair lists, masks, correspondences, k-NN) is a device index.  Results do not depend on the
 * order; run time does (radius search, k-NN and tile culling want consecutive points close in space). */
int tm_scene_upload_sorted(tm_ctx* ctx, const tm_cloud_view* cloud, const uint8_t* tangent_mask,
                           uint32_t* to_user, tm_scene** out);
int tm_scene_set_mask(tm_scene* s, const uint8_t* mask); /* mask_ (scene.hpp:87-90); NULL clears */
void tm_scene_destroy(tm_scene* s);

/* ---- stage calls, host buffers in / out -------------------------------- */
/* pair filter + feature + valid + discretize_feature for scene pairs (i,j):
 * include/impl/scene.hpp:290-302, include/impl/feature.hpp:15-88, src/discretize.cpp:19-30.
 * feats may be NULL. */
int tm_features(tm_scene* s, tm_model* m, const uint32_t* pair_i, const uint32_t* pair_j,
                uint64_t n, float min_diameter_factor, float max_diameter_factor, float* feats,
                uint32_t* keys, uint8_t* valid);
/* model::query + query_limit (include/impl/model.hpp:169-178, scene.hpp:304-311).
 * offsets: n+1 (CSR over pairs); hits: 2 x offsets[n] (m_i, m_j), may be NULL to size. */
int tm_probe(tm_model* m, const uint32_t* keys, const uint8_t* valid, uint64_t n, uint32_t limit,
             uint64_t* offsets, uint32_t* hits, uint64_t hits_capacity);
/* base_transform_ + force_up filter (include/impl/scene.hpp:312-319, 538-567) for
 * every hit of every pair.  T16s: 16 x offsets[n]; hyp_valid: offsets[n]. */
int tm_hypotheses(tm_scene* s, tm_model* m, const uint32_t* pair_i, const uint32_t* pair_j,
                  uint64_t n, const uint64_t* offsets, const uint32_t* hits, int force_up,
                  float* T16s, uint8_t* hyp_valid);
/* radius subset around scene points (scene.hpp:273): ascending indices with
 * ||p - c||^2 < radius^2.  indices may be NULL to size. */
int tm_ball_subsets(tm_scene* s, const uint32_t* centres, uint32_t n_centres, float radius,
                    uint64_t* offsets, int32_t* indices, uint64_t capacity);
/* project_ per hypothesis (include/impl/scene.hpp:411-510): inlier count
 * (= scene_corrs.size()), score, and (early_out != 0) the reference's early-drop
 * outcome.  hyp_sub[h] selects subset CSR row; hyp_sub == NULL scores against all
 * scene points (finish_find, scene.hpp:100-106).  scores/dropped may be NULL.
 * early_out: 0 = no drop; 1 = the drop test over the subset in the order given; 2 = the same test
 * over an evenly sampling walk of the subset (position p visits element (p * s) mod n with
 * s = tm_walk_stride(n)): the test extrapolates from the first 5 %, 10 %, ... of the walk
 * (scene.hpp:486-504), which presumes that a prefix samples the subset evenly; the reference
 * walks FLANN's unsorted radius-search order, ascending indices of a space-filling-curve-ordered
 * scene are spatially compact prefixes.
 * Scores are accumulated as 2^-36 fixed point (order-independent, reproducible across tilings
 * and GPUs): exact for per-point terms |ref . ref_n| < 2^27, i.e. for every rigid transform
 * (terms <= 1); a matrix that scales vectors by more than that is outside the score's domain
 * (inlier counts are unaffected). */
int tm_score(tm_scene* s, tm_model* m, const float* T16s, uint64_t n_hyp, const uint32_t* hyp_sub,
             const uint64_t* sub_offsets, const int32_t* sub_indices, uint32_t n_sub,
             float dist_thres, float accept_prob, int early_out, uint32_t* counts, double* scores,
             uint8_t* dropped);
/* stride of the evenly sampling walk: the integer nearest n / golden ratio that is coprime with n
 * (1 for n <= 2), so that p -> (p * s) mod n is a permutation whose prefixes spread over [0, n) */
uint32_t tm_walk_stride(uint32_t n);
/* early_out = 2 on the resident query: the 18 checkpoint thresholds tests[i] = uint32(0.05f * (i + 1) * n)
 * (scene.hpp:422-426) cut the walk positions [0, n) into 19 ranges; range `level` starts at position
 * tm_early_level_begin(n, level): 0 for level <= 0, max(tests[level - 1] - 1, 0) for 1..18, n for level >= 19.
 * Checkpoint i + 1 fires at the first reaching element at or after the start of range i + 1. */
uint32_t tm_early_level_begin(uint32_t n, int level);
/* correspondences of one transform over the whole scene (finish_find's
 * scene_corrs / model_corrs, ascending scene index).  Buffers sized scene n. */
int tm_correspondences(tm_scene* s, tm_model* m, const float* T16, float dist_thres,
                       uint32_t* scene_corrs, uint32_t* model_corrs, uint32_t* n_corr,
                       double* score);
/* The same for n_T transforms in one pass (the candidate lists of a find_parallel round, scene.hpp:162-190):
 * the correspondences of transform t are scene_corrs / model_corrs [offsets[t], offsets[t + 1]), ascending scene
 * index each; offsets has n_T + 1 entries, scores (optional) n_T.  scene_corrs == model_corrs == NULL sizes only
 * (offsets and scores are still filled); otherwise capacity >= offsets[n_T] entries or TM_ERR_CAPACITY. */
int tm_correspondences_batch(tm_scene* s, tm_model* m, const float* T16s, uint32_t n_T, float dist_thres,
                             uint64_t* offsets, uint32_t* scene_corrs, uint32_t* model_corrs, uint64_t capacity,
                             double* scores);
/* icp_ (include/impl/scene.hpp:369-404; replaces opencl/icp.cl icp_projection +
 * icp_correlation + the absent host reduction) for n start transforms. */
int tm_icp(tm_scene* s, tm_model* m, const float* T16s, uint32_t n, uint32_t max_iterations,
           float dist_thres, float* T16s_out, uint32_t* counts, double* scores, uint32_t* iters);
/* traits project closed forms (a14; replaces opencl/cylinder.cl uv_project).
 * kind: 0 cylinder, 1 plane, 2 plane2, 3 identity.  xyz/uvw packed n x 3; ok: n. */
int tm_traits_project(tm_ctx* ctx, int kind, const float g2l[16], float radius, float threshold,
                      const float* xyz, uint64_t n, float* uvw, uint8_t* ok);

/* ---- model::init pair enumeration (include/impl/model.hpp:100-149) on the device -------------
 * pos3 / tgt3: packed T x 3 positions and tangents of the tangent subset, in subset order.
 * tm_model_pair_bounds = pass 1: feat_min / feat_max over all ordered pairs that pass the
 * distance-window and collinearity filters (components 0..2; component 3 equals 0) and their
 * number.  tm_model_pair_keys = pass 2: for pair (a, b) at keys[a * T + b] the discretised key
 * packed as k0 | k1 << 24 | k2 << 44 (k3 == k0), or ~0 when the pair is filtered or not valid();
 * fmn0 / fmx0 = the distance bounds after valid_bounds(). */
int tm_model_pair_bounds(tm_ctx* ctx, const float* pos3, const float* tgt3, uint32_t T, float lower,
                         float upper, float feat_min[3], float feat_max[3], uint64_t* n_pass);
int tm_model_pair_keys(tm_ctx* ctx, const float* pos3, const float* tgt3, uint32_t T, float lower,
                       float upper, float fmn0, float fmx0, uint32_t steps, float angle_step,
                       uint64_t* keys);

/* ---- pre-processing the reference does with PCL/FLANN ----------------
 * Exact k nearest neighbours (k <= 32) of the resident cloud's points query_idx[0..n_query), the
 * query point included (pointcloud::knn_inclusive, include/impl/pointcloud.hpp:138-152).  Order:
 * ascending (d^2, index), d^2 = (dx*dx + dy*dy) + dz*dz.  out_idx: n_query x k (-1 pads clouds
 * smaller than k); out_d2 may be NULL.  Any cloud can be uploaded with tm_scene_upload for this. */
int tm_scene_knn(tm_scene* s, const uint32_t* query_idx, uint32_t n_query, uint32_t k, int32_t* out_idx,
                 float* out_d2);
/* pointcloud::curvature(k, idx) (include/impl/pointcloud.hpp:200-204 -> principal_curvatures
 * :3-44 + pcl::eigen33): pc_min = evs[1] / k, pc_max = evs[2] / k per query point.  cov9 (optional,
 * n_query x 9, row-major) returns the covariance of the projected normals. */
int tm_scene_curvature(tm_scene* s, const uint32_t* query_idx, uint32_t n_query, uint32_t k, float* pc_min,
                       float* pc_max, float* cov9);
/* The tangent criterion of scene.hpp:50 / model.hpp:98: mask[i] = ||tangent_i|| > 0.7 and
 * pc_min / pc_max < ratio (0.2) with k-NN curvature (k = detail::curvature_k = 30).  mask_out
 * (scene n bytes, may be NULL) receives the mask; apply != 0 also makes it the resident
 * tangent_mask_ of the scene (replacing the one given at upload).  n_tangent (optional) = popcount. */
int tm_scene_tangent_mask(tm_scene* s, uint32_t k, float ratio, uint8_t* mask_out, int apply,
                          uint32_t* n_tangent);

/* The orphaned OpenCL ICP path (a15).  tm_uvicp_projection = opencl/icp.cl:1-53 icp_projection
 * with the projector of opencl/cylinder.cl:1-25 (projector 0) or a linear projector
 * uv = mat_proj * loc (projector 1): per scene point pnts[i] (float4), nearest model sample =
 * the float4 stored in the pixel of the model's n_img = img_size[0]*img_size[1] "uv image".
 * Outputs as the kernel writes them: model_indices / scene_indices (-1 when rejected),
 * out_positions (uv_nrm, or (0,0,0,dist)); n_corr = number of emitted correspondences (what the
 * absent host would have counted).  Matrices are 16 floats, column-major (opencl/util.cl:1-9). */
int tm_uvicp_projection(tm_ctx* ctx, int projector, const float* pnts4, int32_t n, const float* image4,
                        const int32_t img_size[2], const int32_t img_margin[2], const float mat_align[16],
                        const float mat_uvw[16], const float mat_proj[16], const float mat_norm[16],
                        float max_corr_dist, float* out_positions4, int32_t* model_indices,
                        int32_t* scene_indices, uint32_t* n_corr);
/* opencl/icp.cl:55-86 icp_correlation fused with the reduction its host never had: cov9 =
 * sum over k of the nine float terms (scene[is[k]]-c_s) (x) (model[im[k]]-c_m) / (n-1), summed
 * in double (column-major 3x3: cov9[3*j+i] = scene_i * model_j).  records16 (n x 16 floats, the
 * kernel's float16 output) may be NULL. */
int tm_uvicp_correlation(tm_ctx* ctx, const float* scene4, uint32_t n_scene, const float* model4,
                         uint32_t n_model, const int32_t* indices_scene, const int32_t* indices_model, int32_t n,
                         const float centroid_scene[4], const float centroid_model[4], float* records16,
                         double cov9[9]);

/* self-test hook: exclusive prefix sum (n + 1 x u64) of n x u32 with the chained multi-CTA scan that computes the
 * hypothesis offsets of long recorded lists */
int tm_ctx_scan_u64(tm_ctx* ctx, const uint32_t* in, uint64_t n, uint64_t* out);
/* self-test hook: the top-k selection of the resident query's ICP stage on caller-supplied counts (count descending,
 * index ascending; zero counts and entries with excluded[i] != 0 are never selected; excluded may be NULL).
 * ids: k entries, 0xFFFFFFFF where fewer than k qualify. */
int tm_ctx_select_topk(tm_ctx* ctx, const uint32_t* counts, const uint8_t* excluded, uint32_t n, uint32_t k, uint32_t* ids);

/* ---- resident query: the whole recorded-list search in one call --------- */
typedef struct tm_query_params {
    float min_diameter_factor; /* sample_parameters (include/common:72-82) */
    float max_diameter_factor;
    int32_t force_up;
    uint32_t query_limit;      /* detail::query_limit = 200 (scene.hpp:19) */
    float dist_thres;
    float accept_prob;         /* model_match_factor */
    int32_t early_out;         /* 0 = finish_find semantics, 1 = reference early-drop in subset order,
                                  2 = early-drop over the evenly sampling walk (see tm_score), evaluated
                                  checkpoint range by checkpoint range with the tiled scorer */
    uint32_t icp_top_k;        /* 0 = no ICP stage */
    uint32_t max_icp_iterations;
    uint64_t max_hypotheses;   /* capacity; 0 = n_pairs * query_limit */
    uint64_t hyp_limit;        /* score only the first hyp_limit hypotheses (0 = all) */
} tm_query_params;

typedef struct tm_query_result {
    uint64_t n_pairs_valid;
    uint64_t n_hypotheses;      /* global, before sharding */
    uint64_t n_scored;          /* hypotheses scored by this shard */
    uint64_t n_tests;           /* hypothesis-point tests of this shard */
    uint64_t best_key;          /* (inliers << 32) | (0xFFFFFFFF - global hypothesis id); 0 = none */
    uint32_t best_hypothesis;
    uint32_t best_inliers;
    double best_score;
    float best_T[16];
} tm_query_result;

int tm_query_create(tm_scene* s, tm_model* m, const tm_query_params* p, tm_query** out);
void tm_query_destroy(tm_query* q);
/* recorded sample list: outer[o] = scene index of p1; pair k = (outer[pair_outer[k]], pair_j[k]),
 * pairs sorted by pair_outer.  Copies host -> device. */
int tm_query_set_pairs(tm_query* q, const uint32_t* outer, uint32_t n_outer,
                       const uint32_t* pair_outer, const uint32_t* pair_j, uint64_t n_pairs);
/* this rank scores hypotheses [rank*ceil(H/world), ...) of the global list */
int tm_query_set_shard(tm_query* q, uint32_t rank, uint32_t world);
/* How the global hypothesis list is cut into `world` contiguous shards.  by_tests == 0 (default): equal
 * hypothesis counts, [rank*ceil(H/world), ...).  by_tests != 0: equal hypothesis-point tests — a shard's work is
 * sum(|subset(outer)| x hypotheses(outer)), and a step takes as long as the slowest rank.  The split needs the
 * subset size of every outer sample: each rank measures the outer samples of its count-based share and one
 * ncclAllReduce(max) of n_outer x 4 bytes over `comm` completes the table (comm == NULL: every rank sizes all
 * outer samples itself).  All ranks must use the same setting; concatenating the shards in rank order still
 * gives the global list.  Leave headroom in max_hypotheses: shards no longer have equal counts. */
int tm_query_set_balance(tm_query* q, int by_tests, tm_comm* comm);
/* device time of the replicated front end of the last tm_query_run (pair features + probe over the whole list,
 * scan, shard and per-outer ranges) */
int tm_query_frontend_ms(tm_query* q, float* ms);
/* enqueue subsets -> features -> probe -> hypotheses -> score -> argmax (-> ICP) on the
 * context stream; inputs and outputs stay resident */
int tm_query_run(tm_query* q);
int tm_query_result_get(tm_query* q, tm_query_result* out); /* synchronises, small D2H */
/* device time of the scoring kernel of the last tm_query_run (CUDA events on the context
 * stream around that one launch) — the roofline numerator's denominator */
int tm_query_score_kernel_ms(tm_query* q, float* ms);
/* early_out = 2 runs the drop test checkpoint range by checkpoint range with the tiled scorer; a hypothesis one of
 * whose ranges reaches no grid cell is walked on its own afterwards.  *n = how many of the last run's were
 * (0 when the run did not use the level scheme).  Synchronises. */
int tm_query_early_walked(tm_query* q, uint32_t* n);
void* tm_query_best_key_device(tm_query* q); /* resident u64 for the NCCL max-reduce */
int tm_query_set_global_best(tm_query* q, uint64_t key); /* after the all-reduce */
/* full per-hypothesis arrays of this shard (parity tests); any pointer may be NULL */
int tm_query_download(tm_query* q, uint64_t capacity, uint32_t* counts, double* scores,
                      float* T16s, uint8_t* valid, uint32_t* hyp_pair, uint8_t* dropped);
/* top-k ICP results (k = icp_top_k): hypothesis ids, refined transforms, counts */
int tm_query_icp_results(tm_query* q, uint32_t* hyp_ids, float* T16s, uint32_t* counts,
                         double* scores, uint32_t* iters);

/* ---- the one collective: best-pose argmax over ranks ------- */
int tm_nccl_unique_id(uint8_t out[128]);
int tm_comm_create(tm_ctx* ctx, const uint8_t id[128], int rank, int world, tm_comm** out);
void tm_comm_destroy(tm_comm* c);
/* ncclAllReduce(max) of the packed best key, then broadcast of the winner's pose */
int tm_query_allreduce_best(tm_query* q, tm_comm* c);
/* the same for n queries at once (BASELINE configs[3]: 16 models x one scene): one max
 * all-reduce over the n keys and one over the n (score, pose) records. */
int tm_queries_allreduce_best(tm_query** qs, uint32_t n, tm_comm* c);
/* icp_ with the SCENE sharded across ranks (BASELINE configs[4]: top-64 hypotheses refined
 * against a 10 M-point scene on 8 GPUs): this rank accumulates n, sum s, sum m, sum s m^T and
 * the score over its resident points [pt_begin, pt_end) as 64-bit fixed point; one
 * ncclAllReduce(sum, int64) of n x 17 values per iteration makes every rank solve the same
 * rigid fit, so all ranks return identical transforms and the result equals tm_icp's on the
 * whole scene bit for bit.  n_scene_total = scene points over all ranks (it fixes the
 * fixed-point scale and must be the same everywhere).  comm == NULL runs single-process;
 * emulate_parts > 1 then accumulates the range as that many consecutive sub-ranges (how the
 * split-invariance is tested on one GPU). */
int tm_icp_sharded(tm_scene* s, tm_model* m, tm_comm* comm, const float* T16s, uint32_t n,
                   uint32_t max_iterations, float dist_thres, uint32_t pt_begin, uint32_t pt_end,
                   uint64_t n_scene_total, uint32_t emulate_parts, float* T16s_out, uint32_t* counts,
                   double* scores, uint32_t* iters);

/* icp_ with the POSES sharded across ranks (the first option for BASELINE configs[4]): every rank
 * holds the whole scene and refines poses [n*rank/world, n*(rank+1)/world) exactly as tm_icp does — no
 * collective inside the iteration loop — and ONE ncclAllGather of 80-byte records {pose, count, iterations,
 * score} publishes all n results on every rank, bit-identical to tm_icp of the whole list.  comm == NULL
 * runs the given (rank, world) slice in this process and writes only that slice of the outputs (how the
 * split is tested on one GPU and with a CPU-side gather); with a communicator rank / world come from it. */
int tm_icp_pose_sharded(tm_scene* s, tm_model* m, tm_comm* comm, uint32_t rank, uint32_t world, const float* T16s,
                        uint32_t n, uint32_t max_iterations, float dist_thres, float* T16s_out, uint32_t* counts,
                        double* scores, uint32_t* iters);

#ifdef __cplusplus
}
#endif
#endif /* TM_B200_H */
