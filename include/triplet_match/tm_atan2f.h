// tm_atan2f.h — atan2f / atanf of the reference platform's libm, restated.
//
// The reference calls libm's atan2f (include/impl/feature.hpp:7 via angle(); cylinder
// project, include/impl/cylinder_traits.hpp:106).  On the reference's platform (x86-64,
// glibc 2.39: sysdeps/ieee754/flt-32/{e_atan2f,s_atanf}.c, the fdlibm single-precision
// algorithm, built without FMA) that is plain IEEE binary32 +,-,*,/ — restated here so the
// device, the host C++ and the oracle produce the same bits as the reference's CPU path.
// tests/test_oracle_math.py checks it against the host's libm bit for bit (10^7 inputs).
// Compile with FMA contraction off (-fmad=false / -ffp-contract=off).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TM_HD __host__ __device__ inline
#else
#define TM_HD inline
#endif

namespace tm_math {

TM_HD float from_bits(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
TM_HD uint32_t to_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

// atanf: argument reduction to |t| < 7/16 around 0.5, 1, 1.5, inf + odd/even split of an
// 11-term polynomial in t^2
TM_HD float atanf_libm(float x) {
    const uint32_t hx = to_bits(x), ix = hx & 0x7fffffffu;
    const bool neg = (hx >> 31) != 0;
    float hi, lo;
    int id;
    if (ix >= 0x4c000000u) {  // |x| >= 2^25 (or inf / NaN)
        if (ix > 0x7f800000u) return x + x;
        const float r = from_bits(0x3fc90fdau) + from_bits(0x33a22168u);
        return neg ? -r : r;
    }
    if (ix < 0x3ee00000u) {  // |x| < 7/16
        if (ix < 0x31000000u) return x;  // |x| < 2^-29
        id = -1;
        hi = lo = 0.f;
    } else {
        x = from_bits(ix);
        if (ix < 0x3f980000u) {      // |x| < 19/16
            if (ix < 0x3f300000u) {  // 7/16 <= |x| < 11/16
                id = 0;
                hi = from_bits(0x3eed6338u);
                lo = from_bits(0x31ac3769u);
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1;
                hi = from_bits(0x3f490fdau);
                lo = from_bits(0x33222168u);
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else if (ix < 0x401c0000u) {  // |x| < 39/16
            id = 2;
            hi = from_bits(0x3f7b985eu);
            lo = from_bits(0x33140fb4u);
            x = (x - 1.5f) / (1.0f + 1.5f * x);
        } else {
            id = 3;
            hi = from_bits(0x3fc90fdau);
            lo = from_bits(0x33a22168u);
            x = -1.0f / x;
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float a0 = from_bits(0x3eaaaaabu), a1 = from_bits(0xbe4ccccdu), a2 = from_bits(0x3e124925u),
                a3 = from_bits(0xbde38e38u), a4 = from_bits(0x3dba2e6eu), a5 = from_bits(0xbd9d8795u),
                a6 = from_bits(0x3d886b35u), a7 = from_bits(0xbd6ef16bu), a8 = from_bits(0x3d4bda59u),
                a9 = from_bits(0xbd15a221u), a10 = from_bits(0x3c8569d7u);
    const float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    const float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return neg ? -r : r;
}

// atan2f, all quadrants and special values
TM_HD float atan2f_libm(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi = from_bits(0x40490fdbu), pi_o_2 = from_bits(0x3fc90fdbu), pi_o_4 = from_bits(0x3f490fdbu),
                pi_lo = from_bits(0xb3bbbd2eu);
    const uint32_t hx = to_bits(x), hy = to_bits(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;  // NaN
    if (hx == 0x3f800000u) return atanf_libm(y);             // x == 1
    const uint32_t m = (hy >> 31) | ((hx >> 30) & 2u);        // 2*sign(x) + sign(y)
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = ((int32_t)iy - (int32_t)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;             // |y/x| > 2^60
    else if ((hx >> 31) && k < -60) z = 0.0f;          // |y|/x < -2^60
    else z = atanf_libm(from_bits(to_bits(y / x) & 0x7fffffffu));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

}  // namespace tm_math
