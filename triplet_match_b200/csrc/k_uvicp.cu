// k_uvicp.cu — the orphaned OpenCL ICP path (opencl/icp.cl, opencl/cylinder.cl, opencl/util.cl;
// SURVEY §8 row a15), re-created natively.  The reference ships these kernels without any host
// code; what they compute per element is kept bit for bit, what the absent host would have done
// around them (compaction of the -1 entries, the sum over the per-correspondence outer products)
// is fused in:
//   uvicp_projection_kernel    icp_projection (icp.cl:1-53) + uv_project (cylinder.cl:1-25):
//                              one thread per scene point, float4 coalesced load, the four
//                              4x4 matrices in kernel-parameter (constant) space, image gather,
//                              warp-ballot count of emitted correspondences.
//   uvicp_correlation_kernel   icp_correlation (icp.cl:55-86) + the reduction to the 3x3
//                              cross-covariance: every term is the reference's float expression
//                              (s*m)*norm; terms are summed in double by a fixed shuffle/shared
//                              tree and one partial per CTA, so the result is reproducible for a
//                              given n.  The 64-byte float16 records are optional.
#include "tm_kernels.cuh"

namespace tmk {

struct Mat44 {
    float m[16];  // column-major, as in util.cl
};

// util.cl:1-9 — left-to-right sums, no contraction
__device__ __forceinline__ float4 mat44_multiply(float4 p, const Mat44& a) {
    float4 r;
    r.x = ((a.m[0] * p.x + a.m[4] * p.y) + a.m[8] * p.z) + a.m[12] * p.w;
    r.y = ((a.m[1] * p.x + a.m[5] * p.y) + a.m[9] * p.z) + a.m[13] * p.w;
    r.z = ((a.m[2] * p.x + a.m[6] * p.y) + a.m[10] * p.z) + a.m[14] * p.w;
    r.w = ((a.m[3] * p.x + a.m[7] * p.y) + a.m[11] * p.z) + a.m[15] * p.w;
    return r;
}

// cylinder.cl:1-25; atan2pi(y, x) is taken as atan2f(y, x) / pi in binary32 (OpenCL builtin:
// third-party arithmetic, stated in DESIGN.md)
__device__ __forceinline__ float4 uv_project_cylinder(float4 loc, const Mat44& proj) {
    const float4 nc = mat44_multiply(loc, proj);
    float u = atan2f_full(nc.y, nc.x) / 3.14159274101257324219f;
    if (u < 0.f) u += 2.f;
    u /= 2.f;
    const float v = nc.z;
    const float w = sqrtf(nc.x * nc.x + nc.y * nc.y) - 1.0f;
    return make_float4(u, v, w, 1.f);
}

// projector 0: cylinder (cylinder.cl); 1: linear (plane-like: uv = mat_proj * loc)
template <int PROJ>
__global__ void __launch_bounds__(256)
    uvicp_projection_kernel(const float4* __restrict__ pnts, int n, const float4* __restrict__ image,
                            int2 img_size, int2 img_margin, Mat44 mat_align, Mat44 mat_uvw, Mat44 mat_proj,
                            Mat44 mat_norm, float max_corr_dist, float4* __restrict__ out_positions,
                            int* __restrict__ model_indices, int* __restrict__ scene_indices,
                            unsigned int* __restrict__ n_corr) {
    const unsigned int index = blockIdx.x * blockDim.x + threadIdx.x;
    bool emitted = false;
    if (index < (unsigned int)n) {
        const float4 pnt = pnts[index];
        const float4 loc = mat44_multiply(pnt, mat_align);
        const float4 uv = PROJ == 0 ? uv_project_cylinder(loc, mat_proj) : mat44_multiply(loc, mat_proj);
        const float4 uv_nrm = mat44_multiply(mat44_multiply(uv, mat_norm), mat_uvw);
        const float ex = (float)(img_size.x - 2 * img_margin.x - 1), ey = (float)(img_size.y - 2 * img_margin.y - 1);
        // convert_int2: round toward zero; saturating for out-of-range values (as cvt.rzi does)
        int px = (int)((unsigned)__float2int_rz(uv_nrm.x * ex) + (unsigned)img_margin.x);  // wrapping add
        int py = (int)((unsigned)__float2int_rz(uv_nrm.y * ey) + (unsigned)img_margin.y);
        if (py == img_size.y) py = img_size.y - 1;
        int mi = -1, si = -1;
        float4 op = make_float4(0.f, 0.f, 0.f, 0.f);
        if (px >= 0 && px < img_size.x && py >= 0 && py < img_size.y) {
            const int idx = py * img_size.x + px;
            const float4 im = __ldg(&image[idx]);
            const float dx = im.x - uv_nrm.x, dy = im.y - uv_nrm.y;
            const float dist = sqrtf(dx * dx + dy * dy);
            op.w = dist;
            if (dist < max_corr_dist) {
                mi = idx;
                si = (int)index;
                op = uv_nrm;
                emitted = true;
            }
        }
        model_indices[index] = mi;
        scene_indices[index] = si;
        out_positions[index] = op;
    }
    const unsigned int b = __ballot_sync(0xffffffffu, emitted);
    if (n_corr && (threadIdx.x & 31) == 0 && b) atomicAdd(n_corr, __popc(b));
}

__global__ void __launch_bounds__(256)
    uvicp_correlation_kernel(const float4* __restrict__ scene, const float4* __restrict__ model,
                             const int* __restrict__ indices_scene, const int* __restrict__ indices_model, int n,
                             float4 centroid_scene, float4 centroid_model, float* __restrict__ records,
                             double* __restrict__ partials) {
    __shared__ double s_part[8][9];
    const unsigned int index = blockIdx.x * blockDim.x + threadIdx.x;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = 0.f;
    if (index < (unsigned int)n) {
        const float4 s = scene[indices_scene[index]], m = model[indices_model[index]];
        const float sx = s.x - centroid_scene.x, sy = s.y - centroid_scene.y, sz = s.z - centroid_scene.z;
        const float mx = m.x - centroid_model.x, my = m.y - centroid_model.y, mz = m.z - centroid_model.z;
        const float norm = 1.f / (float)(n - 1);
        t[0] = sx * mx * norm; t[1] = sy * mx * norm; t[2] = sz * mx * norm;
        t[3] = sx * my * norm; t[4] = sy * my * norm; t[5] = sz * my * norm;
        t[6] = sx * mz * norm; t[7] = sy * mz * norm; t[8] = sz * mz * norm;
        if (records) {
            float4* o = reinterpret_cast<float4*>(records + 16 * (size_t)index);
            o[0] = make_float4(t[0], t[1], t[2], t[3]);
            o[1] = make_float4(t[4], t[5], t[6], t[7]);
            o[2] = make_float4(t[8], 0.f, 0.f, 0.f);
            o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = (double)t[k];
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
        partials[9 * (size_t)blockIdx.x + threadIdx.x] = v;
    }
}

// final fixed-order sum of the per-CTA partials (one thread per matrix entry)
__global__ void uvicp_correlation_finish_kernel(const double* __restrict__ partials, int n_blocks,
                                                double* __restrict__ cov9) {
    const int k = threadIdx.x;
    if (k >= 9) return;
    double v = 0.0;
    for (int b = 0; b < n_blocks; ++b) v += partials[9 * (size_t)b + k];
    cov9[k] = v;
}

static Mat44 mat_of(const float* p) {
    Mat44 m;
    for (int i = 0; i < 16; ++i) m.m[i] = p[i];
    return m;
}

void launch_uvicp_projection(cudaStream_t st, int projector, const float4* pnts, int n, const float4* image,
                             const int img_size[2], const int img_margin[2], const float* mat_align,
                             const float* mat_uvw, const float* mat_proj, const float* mat_norm,
                             float max_corr_dist, float4* out_positions, int* model_indices, int* scene_indices,
                             unsigned int* n_corr) {
    if (n <= 0) return;
    const int grid = (n + 255) / 256;
    const int2 is = make_int2(img_size[0], img_size[1]), im = make_int2(img_margin[0], img_margin[1]);
    if (projector == 0)
        uvicp_projection_kernel<0><<<grid, 256, 0, st>>>(pnts, n, image, is, im, mat_of(mat_align), mat_of(mat_uvw),
                                                         mat_of(mat_proj), mat_of(mat_norm), max_corr_dist,
                                                         out_positions, model_indices, scene_indices, n_corr);
    else
        uvicp_projection_kernel<1><<<grid, 256, 0, st>>>(pnts, n, image, is, im, mat_of(mat_align), mat_of(mat_uvw),
                                                         mat_of(mat_proj), mat_of(mat_norm), max_corr_dist,
                                                         out_positions, model_indices, scene_indices, n_corr);
    ++g_launch_count;
}

int uvicp_correlation_blocks(int n) { return n > 0 ? (n + 255) / 256 : 0; }

void launch_uvicp_correlation(cudaStream_t st, const float4* scene, const float4* model, const int* indices_scene,
                              const int* indices_model, int n, const float* centroid_scene,
                              const float* centroid_model, float* records, double* partials, double* cov9) {
    const int grid = uvicp_correlation_blocks(n);
    if (grid) {
        uvicp_correlation_kernel<<<grid, 256, 0, st>>>(
            scene, model, indices_scene, indices_model, n,
            make_float4(centroid_scene[0], centroid_scene[1], centroid_scene[2], centroid_scene[3]),
            make_float4(centroid_model[0], centroid_model[1], centroid_model[2], centroid_model[3]), records, partials);
        ++g_launch_count;
    }
    uvicp_correlation_finish_kernel<<<1, 32, 0, st>>>(partials, grid, cov9);
    ++g_launch_count;
}

}  // namespace tmk
