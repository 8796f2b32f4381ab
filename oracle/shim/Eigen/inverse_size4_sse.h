// oracle/shim/Eigen/inverse_size4_sse.h — TEST INFRASTRUCTURE.
// Eigen is absent from this image.  The reference calls Matrix4f::inverse() once on the hot path's
// data (model.hpp:63 `mat4f_t inv = to_voxel_.inverse()`, used for the voxel centres at :87).  For a
// fixed 4x4 float matrix on x86-64 (SSE2 is baseline, no -march in CMakeLists.txt:28-40) Eigen 3.3
// dispatches to compute_inverse_size4<Architecture::SSE, float, ...> (Eigen/src/LU/arch/Inverse_SSE.h;
// Eigen 3.4 carries the same arithmetic as Eigen/src/LU/arch/InverseSize4.h), the 2x2-block cofactor
// routine published by Intel (AP-928 "Streaming SIMD Extensions - Inverse of 4x4 Matrix", the
// divide-and-conquer variant).  This file restates that routine lane by lane: every _mm_* operation is
// an IEEE binary32 operation per lane, so a scalar emulation of the same shuffles, products, sums and
// the one division reproduces its roundings.  Restated from the published algorithm, not copied from
// Eigen (no Eigen source is available here); what pins it:
//   * tests/test_voxel_centre.py: M * inverse(M) = I to rounding on random general matrices (a wrong
//     shuffle would not give an inverse), and the closed form used by the oracle and the product for
//     to_voxel_ = diag(s) + t (include/triplet_match/tm_voxel_centre.h) equals this routine bit for bit.
#pragma once

namespace eigen_shim_sse {

struct V4 {
    float v[4];
};
inline V4 mul(V4 a, V4 b) { return {{a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]}}; }
inline V4 add(V4 a, V4 b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]}}; }
inline V4 sub(V4 a, V4 b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2], a.v[3] - b.v[3]}}; }
// scalar forms: lane 0 = op(a0, b0), lanes 1..3 = a's
inline V4 add_ss(V4 a, V4 b) { a.v[0] = a.v[0] + b.v[0]; return a; }
inline V4 sub_ss(V4 a, V4 b) { a.v[0] = a.v[0] - b.v[0]; return a; }
inline V4 mul_ss(V4 a, V4 b) { a.v[0] = a.v[0] * b.v[0]; return a; }
inline V4 div_ss(V4 a, V4 b) { a.v[0] = a.v[0] / b.v[0]; return a; }
// _mm_shuffle_ps(a, b, imm): lanes (a[imm & 3], a[(imm >> 2) & 3], b[(imm >> 4) & 3], b[(imm >> 6) & 3])
inline V4 shuf(V4 a, V4 b, int imm) {
    return {{a.v[imm & 3], a.v[(imm >> 2) & 3], b.v[(imm >> 4) & 3], b.v[(imm >> 6) & 3]}};
}
inline V4 movelh(V4 a, V4 b) { return {{a.v[0], a.v[1], b.v[0], b.v[1]}}; }
inline V4 movehl(V4 a, V4 b) { return {{b.v[2], b.v[3], a.v[2], a.v[3]}}; }
inline V4 neg_lanes(V4 a, bool n0, bool n1, bool n2, bool n3) {  // xor with a sign mask
    return {{n0 ? -a.v[0] : a.v[0], n1 ? -a.v[1] : a.v[1], n2 ? -a.v[2] : a.v[2], n3 ? -a.v[3] : a.v[3]}};
}

// lines[k] = the k-th group of four consecutive stored coefficients (a column of a column-major
// Matrix4f); the result is written in the same layout.
inline void inverse4(const float lines[4][4], float out[4][4]) {
    V4 L1 = {{lines[0][0], lines[0][1], lines[0][2], lines[0][3]}}, L2 = {{lines[1][0], lines[1][1], lines[1][2], lines[1][3]}},
       L3 = {{lines[2][0], lines[2][1], lines[2][2], lines[2][3]}}, L4 = {{lines[3][0], lines[3][1], lines[3][2], lines[3][3]}};
    // four 2x2 sub-matrices, each in one register (storage orders of source and result match)
    V4 A = movelh(L1, L2), B = movehl(L2, L1), C = movelh(L3, L4), D = movehl(L4, L3);
    // AB = A# * B, DC = D# * C
    V4 AB = mul(shuf(A, A, 0x0F), B);
    AB = sub(AB, mul(shuf(A, A, 0xA5), shuf(B, B, 0x4E)));
    V4 DC = mul(shuf(D, D, 0x0F), C);
    DC = sub(DC, mul(shuf(D, D, 0xA5), shuf(C, C, 0x4E)));
    // determinants of the sub-matrices (lane 0)
    V4 dA = mul(shuf(A, A, 0x5F), A);
    dA = sub_ss(dA, movehl(dA, dA));
    V4 dB = mul(shuf(B, B, 0x5F), B);
    dB = sub_ss(dB, movehl(dB, dB));
    V4 dC = mul(shuf(C, C, 0x5F), C);
    dC = sub_ss(dC, movehl(dC, dC));
    V4 dD = mul(shuf(D, D, 0x5F), D);
    dD = sub_ss(dD, movehl(dD, dD));
    // d = trace(AB * DC)
    V4 d = mul(shuf(DC, DC, 0xD8), AB);
    // iD = C * A# * B, iA = B * D# * C
    V4 iD = mul(shuf(C, C, 0xA0), movelh(AB, AB));
    iD = add(iD, mul(shuf(C, C, 0xF5), movehl(AB, AB)));
    V4 iA = mul(shuf(B, B, 0xA0), movelh(DC, DC));
    iA = add(iA, mul(shuf(B, B, 0xF5), movehl(DC, DC)));
    d = add(d, movehl(d, d));
    d = add_ss(d, shuf(d, d, 1));
    V4 d1 = mul_ss(dA, dD), d2 = mul_ss(dB, dC);
    // iD = D * |A| - C * A# * B,  iA = A * |D| - B * D# * C
    iD = sub(mul(D, shuf(dA, dA, 0)), iD);
    iA = sub(mul(A, shuf(dD, dD, 0)), iA);
    // det = |A| |D| + |B| |C| - trace(A# B D# C), rd = 1 / det (a true division, not rcpps)
    V4 det = sub_ss(add_ss(d1, d2), d);
    V4 one = {{1.f, 0.f, 0.f, 0.f}};
    V4 rd = div_ss(one, det);
    // iB = D * (A# B)#,  iC = A * (D# C)#
    V4 iB = mul(D, shuf(AB, AB, 0x33));
    iB = sub(iB, mul(shuf(D, D, 0xB1), shuf(AB, AB, 0x66)));
    V4 iC = mul(A, shuf(DC, DC, 0x33));
    iC = sub(iC, mul(shuf(A, A, 0xB1), shuf(DC, DC, 0x66)));
    rd = shuf(rd, rd, 0);
    rd = neg_lanes(rd, false, true, true, false);  // sign mask (+, -, -, +)
    // iB = C * |B| - D * B# * A,  iC = B * |C| - A * C# * D
    iB = sub(mul(C, shuf(dB, dB, 0)), iB);
    iC = sub(mul(B, shuf(dC, dC, 0)), iC);
    iA = mul(rd, iA);
    iB = mul(rd, iB);
    iC = mul(rd, iC);
    iD = mul(rd, iD);
    V4 r0 = shuf(iA, iB, 0x77), r1 = shuf(iA, iB, 0x22), r2 = shuf(iC, iD, 0x77), r3 = shuf(iC, iD, 0x22);
    for (int k = 0; k < 4; ++k) {
        out[0][k] = r0.v[k];
        out[1][k] = r1.v[k];
        out[2][k] = r2.v[k];
        out[3][k] = r3.v[k];
    }
}

}  // namespace eigen_shim_sse
