/* tm_b200_host.h — host-side half of model<Point>::init behind the same C-ABI.
 *
 * model::init (include/impl/model.hpp:16-167) is the offline step that produces
 * what the search path probes: the voxel -> nearest-model-point grid and the
 * feature hash multimap.  tm_hostmodel_build restates it on the host (pair
 * enumeration, feature bounds, multimap in libstdc++ insertion order) and runs
 * the per-voxel 1-NN fill on the GPU when `ctx` is non-NULL (exact brute force,
 * lowest index wins ties) or on an exact host grid search when it is NULL.
 * The result feeds tm_model_upload / tm_model_create.
 */
#ifndef TM_B200_HOST_H
#define TM_B200_HOST_H

#include "tm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tm_hostmodel tm_hostmodel;

const char* tm_host_last_error(void);
/* pointcloud::resolution() (include/impl/pointcloud.hpp:66-82): running mean of 1-NN distances */
float tm_host_resolution(const tm_cloud_view* cloud);
/* curv_ok[i] != 0 replaces the PCL curvature-ratio test (model.hpp:98); NULL = all pass.
 * resolution <= 0 => computed.  cap = values kept per key (query_limit, scene.hpp:19). */
int tm_hostmodel_build(tm_ctx* ctx, const tm_cloud_view* cloud, const uint8_t* curv_ok,
                       float distance_step_count, float angle_step, float min_diameter_factor,
                       float max_diameter_factor, float resolution, uint32_t cap,
                       tm_hostmodel** out);
/* model::init(subset, params) (model.hpp:16-39): in_subset[i] != 0 marks the caller's subset (NULL = every point).
 * As in the reference the bounding box, diameter, voxel-grid geometry and the tangent subset come from the
 * finite points OF THE SUBSET, while the nearest-neighbour grid is filled from the whole cloud. */
int tm_hostmodel_build_subset(tm_ctx* ctx, const tm_cloud_view* cloud, const uint8_t* in_subset, const uint8_t* curv_ok,
                              float distance_step_count, float angle_step, float min_diameter_factor,
                              float max_diameter_factor, float resolution, uint32_t cap, tm_hostmodel** out);
void tm_hostmodel_destroy(tm_hostmodel* m);
/* Model blob: serialises what model::init produced (grid, hash table in insertion and in
 * equal_range order, bounds) so a model is built once and reloaded — the reference rebuilds both
 * on every run.  n_cloud_points ties the blob to its cloud; load verifies magic, version, an
 * FNV-1a checksum and every index before anything reaches the device. */
int tm_hostmodel_save(const tm_hostmodel* m, uint32_t n_cloud_points, const char* path);
int tm_hostmodel_load(const char* path, uint32_t n_cloud_points, tm_hostmodel** out);
/* pointers in *d stay valid while the host model lives */
void tm_hostmodel_desc(const tm_hostmodel* m, tm_model_desc* d);
void tm_hostmodel_counts(const tm_hostmodel* m, uint64_t* n_subset, uint64_t* n_entries,
                         uint64_t* n_keys, uint64_t* n_kept);
const uint32_t* tm_hostmodel_subset(const tm_hostmodel* m);      /* tangent subset (point_count()) */
const uint32_t* tm_hostmodel_entry_keys(const tm_hostmodel* m);  /* 4 x n_entries, insertion order */
const uint32_t* tm_hostmodel_entry_pairs(const tm_hostmodel* m); /* 2 x n_entries */
/* upload; hm == NULL reproduces "Cannot query uninitialized model" (model.hpp:171-173) */
int tm_model_create(tm_ctx* ctx, const tm_cloud_view* cloud, const tm_hostmodel* hm,
                    tm_model** out);

#ifdef __cplusplus
}
#endif
#endif /* TM_B200_HOST_H */
