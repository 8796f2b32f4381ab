// oracle/ref_cl_capi.cpp — TEST INFRASTRUCTURE: host loops over the reference's OpenCL kernels
// (opencl/icp.cl:1-53, 55-86), compiled from their own source by `make -C oracle ref` through
// oracle/shim/opencl/opencl_cxx.hpp.  Same C signatures as orc_cl_icp_projection / orc_cl_icp_correlation.
#include <cstdint>
#include <cstring>

#include "shim/opencl/opencl_cxx.hpp"

// defined by the piped translation unit (util.cl + cylinder.cl + icp.cl)
void icp_projection(const float4* pnts, int n, const float4* image, int2 img_size, int2 img_margin,
                    const float* mat_align, const float* mat_uvw, const float* mat_proj, const float* mat_norm,
                    float max_corr_dist, float4* out_positions, int* model_indices, int* scene_indices);
void icp_correlation(const float4* scene, const float4* model, const int* indices_scene, const int* indices_model, int n,
                     float4 centroid_scene, float4 centroid_model, float16* output);

extern "C" {
// `extra` more work-items than n are launched too (a padded NDRange): the kernel's own guard must stop them
uint32_t ref_cl_icp_projection(const float* pnts4, int n, int extra, const float* image4, const int32_t* img_size,
                               const int32_t* img_margin, const float* mat_align, const float* mat_uvw,
                               const float* mat_proj, const float* mat_norm, float max_corr_dist, float* out_positions4,
                               int32_t* model_indices, int32_t* scene_indices) {
    static_assert(sizeof(float4) == 16 && sizeof(float16) == 64, "vector layouts");
    for (int i = 0; i < n + extra; ++i) {
        cl_current_global_id() = (uint)i;
        icp_projection(reinterpret_cast<const float4*>(pnts4), n, reinterpret_cast<const float4*>(image4),
                       int2(img_size[0], img_size[1]), int2(img_margin[0], img_margin[1]), mat_align, mat_uvw, mat_proj,
                       mat_norm, max_corr_dist, reinterpret_cast<float4*>(out_positions4), model_indices, scene_indices);
    }
    uint32_t c = 0;
    for (int i = 0; i < n; ++i) c += model_indices[i] >= 0;
    return c;
}
void ref_cl_icp_correlation(const float* scene4, const float* model4, const int32_t* is, const int32_t* im, int n,
                            int extra, const float* cs, const float* cm, float* records16) {
    for (int i = 0; i < n + extra; ++i) {
        cl_current_global_id() = (uint)i;
        icp_correlation(reinterpret_cast<const float4*>(scene4), reinterpret_cast<const float4*>(model4), is, im, n,
                        float4(cs[0], cs[1], cs[2], cs[3]), float4(cm[0], cm[1], cm[2], cm[3]),
                        reinterpret_cast<float16*>(records16));
    }
}
}
