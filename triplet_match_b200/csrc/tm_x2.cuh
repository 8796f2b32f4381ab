// tm_x2.cuh — the reference's no-FMA FP32 arithmetic on the packed FP32 pipe (FFMA2), exactly.
//   a * b  ==  fma(a, b, -0)   (one rounding of the exact product; (+0) + (-0) = +0, (-0) + (-0) = -0)
//   a + b  ==  fma(a,  1,  b)  (one rounding of the exact sum)
// The constants -0 and 1 arrive as kernel arguments, so ptxas cannot fold fma(a, b, -0) back into a multiply and
// contract it with the add that follows (which it does for mul.rn.f32x2 + add.rn.f32x2 even under -fmad=false).
#pragma once
#include <cstring>

#include "tm_device.cuh"

namespace tmk {

// ---- exact packed FP32 -----------------------------------------------------------------------------
// A pair of floats lives in one 64-bit register (p2) from load to last use, so ptxas allocates it an aligned
// register pair once instead of re-assembling it from two scalars for every FFMA2.
typedef unsigned long long p2;
__device__ __forceinline__ p2 pack2(float lo, float hi) {
    p2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(p2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi2(p2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ p2 ffma2(p2 a, p2 b, p2 c) {
    p2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ p2 splat(float v) { return pack2(v, v); }
struct X2 {  // the three opaque constants
    p2 nz, one, mone;
    __device__ __forceinline__ p2 mul(p2 a, p2 b) const { return ffma2(a, b, nz); }
    __device__ __forceinline__ p2 mul(float a, p2 b) const { return ffma2(splat(a), b, nz); }
    __device__ __forceinline__ p2 add(p2 a, p2 b) const { return ffma2(a, one, b); }
    __device__ __forceinline__ p2 add(p2 a, float b) const { return ffma2(a, one, splat(b)); }
    // row of Matrix4f * (x,y,z,1) for two points: ((r.x*x + r.y*y) + r.z*z) + r.w  (tm_device.cuh row_apply)
    __device__ __forceinline__ p2 row_apply(float4 r, p2 x, p2 y, p2 z) const {
        return add(add(add(mul(r.x, x), mul(r.y, y)), mul(r.z, z)), r.w);
    }
    // a0 + (a1 + a2) (tm_device.cuh sum3) of the squares
    __device__ __forceinline__ p2 sqnorm(p2 dx, p2 dy, p2 dz) const {
        return add(mul(dx, dx), add(mul(dy, dy), mul(dz, dz)));
    }
};

static inline p2 host_pair(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    return ((p2)b << 32) | b;
}

}  // namespace tmk
