/* tm_voxel_centre.h — world position of a voxel centre, as model::init computes it.
 *
 * Reference: include/impl/model.hpp:63 `mat4f_t inv = to_voxel_.inverse();` and :87
 * `uvw = (inv * vec4f_t(i, j, k, 1.f)).head(3)`, followed by a 1-NN query at uvw.  Which model point
 * is nearest to a cell centre decides the voxel grid, so the centre has to be the reference's float,
 * not merely a close one: two model points can be 1e-8 apart in distance from a centre.
 *
 * to_voxel_ is diag(sx, sy, sz) plus a translation (model.hpp:56-61).  Matrix4f::inverse() is Eigen's
 * SSE 2x2-block cofactor routine (Intel AP-928; Eigen/src/LU/arch/Inverse_SSE.h) — restated lane by
 * lane in oracle/shim/Eigen/inverse_size4_sse.h.  For this matrix shape every product with a structural
 * zero is an exact zero, and what is left of that routine is the closed form below (same products, same
 * order, one division); tests/test_voxel_centre.py checks the two against each other bit for bit.
 * The product inv * (i, j, k, 1) is Eigen's packet order ((c0*i + c1*j) + c2*k) + c3*1, of which only
 * a*i + b survives per axis.
 *
 * Compile with contraction off (-fmad=false / -ffp-contract=off): a*i + b must round twice. */
#ifndef TM_VOXEL_CENTRE_H
#define TM_VOXEL_CENTRE_H

#if defined(__CUDACC__)
#define TM_VC_HD __host__ __device__
#else
#define TM_VC_HD
#endif

typedef struct tm_centre_map {
    float a[3]; /* inverse scale per axis: inv(0,0), inv(1,1), inv(2,2) */
    float b[3]; /* inverse translation per axis: inv(0,3), inv(1,3), inv(2,3) */
} tm_centre_map;

/* s = diagonal of to_voxel_, t = its translation column */
static inline tm_centre_map tm_voxel_centre_map(const float s[3], const float t[3]) {
    tm_centre_map m;
    const float sxy = s[0] * s[1];      /* |A| of the upper-left 2x2 block */
    const float det = sxy * s[2];       /* |A| * |D|; the other two terms of the determinant are exact zeros */
    const float rd = 1.0f / det;
    const float p = s[2] * t[0];        /* D# * C, second row */
    const float q = s[2] * t[1];
    m.a[0] = rd * (s[1] * s[2]);
    m.a[1] = rd * (s[0] * s[2]);
    m.a[2] = rd * sxy;
    /* written as the routine leaves them, so that a zero translation keeps the routine's sign of zero */
    m.b[0] = (-rd) * (0.0f - (0.0f - s[1] * p));
    m.b[1] = rd * (0.0f - s[0] * q);
    m.b[2] = (-rd) * (t[2] * sxy);
    return m;
}

/* centre of cell index i along one axis */
TM_VC_HD static inline float tm_voxel_centre(float a, float b, int i) { return a * (float)i + b; }

#endif /* TM_VOXEL_CENTRE_H */
