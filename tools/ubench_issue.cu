// ubench_issue.cu — issue-rate micro-benchmark behind DESIGN.md §4's roofline for the scorer:
// warp instructions per cycle per SM for FMUL+FADD (the scalar exact path), FFMA2 (the packed
// exact path of k_score2.cu) and F2I.TRUNC, each with 8 independent chains per thread and
// 8 resident warps per scheduler, timed with clock64() inside the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/ubench_issue tools/ubench_issue.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef unsigned long long p2;
__device__ __forceinline__ p2 ffma2(p2 a, p2 b, p2 c) {
    p2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int ITERS = 4096;

__global__ void k_scalar(float* out, float a, float b, long long* cyc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.001f + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] * a + b;  // -fmad=false: FMUL + FADD
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma(float* out, float a, float b, long long* cyc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.001f + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = __fmaf_rn(v[i], a, b);
            v[i] = __fmaf_rn(v[i], a, b);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_packed(float* out, p2 a, p2 b, long long* cyc) {
    p2 v[8];
    for (int i = 0; i < 8; ++i) v[i] = (p2)(threadIdx.x + i) * 0x0000100000001000ull + 0x3f8000003f800000ull;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = ffma2(v[i], a, b);
            v[i] = ffma2(v[i], a, b);
        }
    }
    long long t1 = clock64();
    p2 s = 0;
    for (int i = 0; i < 8; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s ^ (uint32_t)(s >> 32));
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_f2i(float* out, float a, long long* cyc) {
    float v[8];
    int acc[8];
    for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 0.37f + i; acc[i] = 0; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] += (int)v[i];      // F2I + IADD
            v[i] = __int_as_float(__float_as_int(v[i]) ^ (it & 3));  // LOP3 keeps the input changing
        }
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s + a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 1024, blocks = sms;  // 32 warps per SM = 8 per scheduler
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaMallocManaged(&cyc, sizeof(long long) * blocks);
    auto report = [&](const char* name, double warp_instr_per_warp) {
        cudaDeviceSynchronize();
        double c = 0;
        for (int b = 0; b < blocks; ++b) c += (double)cyc[b];
        c /= blocks;
        const double per_sm = warp_instr_per_warp * (threads / 32) / c;
        printf("{\"kernel\": \"%s\", \"cycles\": %.0f, \"warp_instr_per_clk_per_sm\": %.3f, \"per_scheduler\": %.3f}\n", name, c,
               per_sm, per_sm / 4);
    };
    const unsigned one = 0x3f800000u, tiny = 0x33800000u;
    const p2 a2 = ((p2)one << 32) | one, b2 = ((p2)tiny << 32) | tiny;
    for (int rep = 0; rep < 2; ++rep) {
        k_scalar<<<blocks, threads>>>(out, 1.0f, 1e-7f, cyc);
        report("fmul+fadd (2 instr per step)", 2.0 * 8 * ITERS);
        k_ffma<<<blocks, threads>>>(out, 1.0f, 1e-7f, cyc);
        report("ffma", 2.0 * 8 * ITERS);
        k_packed<<<blocks, threads>>>(out, a2, b2, cyc);
        report("ffma2 (two fp32 fma per instr)", 2.0 * 8 * ITERS);
        k_f2i<<<blocks, threads>>>(out, 0.f, cyc);
        report("f2i.trunc (+iadd +lop3: 3 instr per step)", 3.0 * 8 * ITERS);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
