// oracle/shim/opencl/opencl_cxx.hpp — TEST INFRASTRUCTURE.  Just enough of OpenCL C, written as C++,
// to compile the reference's opencl/{util,cylinder,icp}.cl AS THEY ARE (the Makefile pipes them through
// a sed that turns the vector literals `(float4)(a, b, c, d)` into constructor calls — nothing else)
// and run their kernels one work-item at a time on the host.
//
// What this pins and what it does not: the kernels' own statements (matrix layout, operation order,
// pixel arithmetic, bounds rule, the outputs written in each branch) are the reference's.  The
// OpenCL BUILT-INS are stand-ins defined here, and an OpenCL device may round them differently
// (atan2pi and length are <= a few ulp by the standard, convert_int2 of out-of-range values is
// implementation-defined): atan2pi(y, x) = atan2f(y, x) / pi_f, length(v) = sqrtf(x*x + y*y),
// convert_int2 = saturating truncation with NaN -> 0, no FMA contraction.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

#define __kernel
#define __global
typedef unsigned int uint;

struct float2 {  // plain aggregate: it sits inside float4's anonymous struct (.xy / .zw)
    float x, y;
};
struct int2 {
    int x, y;
    int2() = default;
    int2(int a, int b) : x(a), y(b) {}
};
struct float4 {
    union {
        struct { float x, y, z, w; };
        struct { float2 xy, zw; };
    };
    float4() = default;
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
struct float16 {
    float s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sa, sb, sc, sd, se, sf;
};
inline float2 operator-(float2 a, float2 b) { return float2{a.x - b.x, a.y - b.y}; }
inline float2 operator*(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
// integer vector arithmetic wraps in OpenCL C; unsigned detour keeps that defined in C++
inline int wrap_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
inline int wrap_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
inline int wrap_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
inline int2 operator-(int2 a, int2 b) { return int2(wrap_sub(a.x, b.x), wrap_sub(a.y, b.y)); }
inline int2 operator+(int2 a, int2 b) { return int2(wrap_add(a.x, b.x), wrap_add(a.y, b.y)); }
inline int2 operator*(int s, int2 a) { return int2(wrap_mul(s, a.x), wrap_mul(s, a.y)); }
inline float2 convert_float2(int2 a) { return float2{(float)a.x, (float)a.y}; }
inline int convert_int_sat_rtz(float v) {
    if (!(v == v)) return 0;
    if (v >= 2147483648.f) return std::numeric_limits<int>::max();
    if (v <= -2147483648.f) return std::numeric_limits<int>::min();
    return (int)v;
}
inline int2 convert_int2(float2 a) { return int2(convert_int_sat_rtz(a.x), convert_int_sat_rtz(a.y)); }
inline float length(float2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
inline float atan2pi(float y, float x) { return atan2f(y, x) / 3.14159274101257324219f; }

// the work-item the host loop is currently running
inline uint& cl_current_global_id() {
    static thread_local uint id = 0;
    return id;
}
inline uint get_global_id(int) { return cl_current_global_id(); }
