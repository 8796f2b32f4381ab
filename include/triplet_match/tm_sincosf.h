// tm_sincosf.h — sinf / cosf for |x| <= 1.6, identical bits on the host, the device and in the oracle.
//
// Where it is used: pcl::eigen33's closed form (the principal-curvature criterion pc_min / pc_max < 0.2 behind
// the tangent masks, include/impl/pointcloud.hpp:3-44 -> scene.hpp:50, model.hpp:98) evaluates cos and sin of
// theta = atan2(sqrt(-q), half_b) / 3, i.e. theta in [0, pi/3].  The reference gets them from its platform's
// libm; the device's cosf / sinf differ from that in the last bit now and then, which moves an eigenvalue by
// 1e-7 relative and can flip a mask bit of a point that sits on the 0.2 boundary — and a mask bit changes every
// "bit-exact" quantity downstream.  So device and oracle share this routine instead (like tm_atan2f.h).
//
// Algorithm: the Taylor series in binary64, Horner form, plain * and + (no FMA: compile with contraction off),
// rounded once to binary32.  For |x| <= 1.6 the truncation error is below 5e-16 relative, so the result is the
// correctly rounded binary32 value except when the exact value lies within ~1e-8 ulp of a rounding boundary.
// It equals (float)sin((double)x) / (float)cos((double)x) for every one of the 1.07e9 binary32 values of [0, 1.6]
// (checked exhaustively).  glibc's sinf / cosf (double-precision polynomial, up to 0.56 ulp, not correctly rounded)
// differ from it on about 1.5 % of uniformly drawn arguments, never by more than 1 ulp; tests/test_oracle_math.py
// measures both.
#pragma once

#if defined(__CUDACC__)
#define TM_SC_HD __host__ __device__ inline
#else
#define TM_SC_HD inline
#endif

namespace tm_math {

// sin x = x + x * z * S(z), z = x^2, S(z) = -1/3! + z/5! - z^2/7! + ... (through x^21)
TM_SC_HD float sinf_small(float xf) {
    const double x = (double)xf, z = x * x;
    double s = -1.0 / 51090942171709440000.0;                 // -1/21!
    s = s * z + 1.0 / 121645100408832000.0;                   //  1/19!
    s = s * z + -1.0 / 355687428096000.0;                     // -1/17!
    s = s * z + 1.0 / 1307674368000.0;                        //  1/15!
    s = s * z + -1.0 / 6227020800.0;                          // -1/13!
    s = s * z + 1.0 / 39916800.0;                             //  1/11!
    s = s * z + -1.0 / 362880.0;                              // -1/9!
    s = s * z + 1.0 / 5040.0;                                 //  1/7!
    s = s * z + -1.0 / 120.0;                                 // -1/5!
    s = s * z + 1.0 / 6.0;                                    //  1/3!  (sign folded below)
    return (float)(x - (x * z) * s);
}
// cos x = 1 - z * C(z), C(z) = 1/2! - z/4! + z^2/6! - ... (through x^20)
TM_SC_HD float cosf_small(float xf) {
    const double x = (double)xf, z = x * x;
    double c = -1.0 / 2432902008176640000.0;                  // -1/20!
    c = c * z + 1.0 / 6402373705728000.0;                     //  1/18!
    c = c * z + -1.0 / 20922789888000.0;                      // -1/16!
    c = c * z + 1.0 / 87178291200.0;                          //  1/14!
    c = c * z + -1.0 / 479001600.0;                           // -1/12!
    c = c * z + 1.0 / 3628800.0;                              //  1/10!
    c = c * z + -1.0 / 40320.0;                               // -1/8!
    c = c * z + 1.0 / 720.0;                                  //  1/6!
    c = c * z + -1.0 / 24.0;                                  // -1/4!
    c = c * z + 1.0 / 2.0;                                    //  1/2!
    return (float)(1.0 - z * c);
}

}  // namespace tm_math
