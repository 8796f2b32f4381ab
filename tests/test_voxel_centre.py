"""Voxel-centre arithmetic of model::init (include/impl/model.hpp:63 `to_voxel_.inverse()`, :87
`inv * (i, j, k, 1)`): which model point is nearest to a cell centre decides the voxel grid, so
oracle, shim-compiled reference and product must compute the same float for every centre.

  * oracle/shim/Eigen/inverse_size4_sse.h restates Eigen's SSE Matrix4f::inverse() (Intel AP-928
    2x2-block cofactor routine) lane by lane; it is what the reference sources see when compiled here.
    Checked as an inverse on random general matrices (a wrong shuffle would not invert).
  * For to_voxel_ = diag(s) + t that routine collapses to a closed form; the oracle's copy
    (oracle.hpp voxel_centre_map) and the product's (include/triplet_match/tm_voxel_centre.h) must equal
    the routine bit for bit, signs of zero included.
  * The product header is compiled here with gcc into a scratch .so (contraction off)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libtm_ref.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def header_lib(built):
    src = (
        '#include "%s"\n'
        "void hdr_centre_map(const float* s, const float* t, float* a, float* b) {\n"
        "  tm_centre_map m = tm_voxel_centre_map(s, t);\n"
        "  for (int k = 0; k < 3; ++k) { a[k] = m.a[k]; b[k] = m.b[k]; }\n"
        "}\n"
        "float hdr_centre(float a, float b, int i) { return tm_voxel_centre(a, b, i); }\n"
    ) % os.path.join(ROOT, "include", "triplet_match", "tm_voxel_centre.h")
    d = tempfile.mkdtemp(prefix="tm_vc_")
    c, so = os.path.join(d, "vc.c"), os.path.join(d, "vc.so")
    with open(c, "w") as f:
        f.write(src)
    subprocess.run(["gcc", "-O3", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, c], check=True)
    L = C.CDLL(so)
    L.hdr_centre.restype = C.c_float
    L.hdr_centre.argtypes = [C.c_float, C.c_float, C.c_int]
    return L


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.5, 400.0, size=(n, 3)).astype(np.float32)
    t = rng.uniform(-500.0, 500.0, size=(n, 3)).astype(np.float32)
    s[::7, 2] = 1.0      # degenerate axis: scale 1 (model.hpp:52-55)
    t[::11, 1] = 0.0     # sign of zero
    s[::13] = np.float32(199.99998)  # equal scales on all axes (isotropic clouds)
    return s, t


def test_oracle_equals_product_header(header_lib):
    L = po.load()
    s, t = _cases(20000, 3)
    for k in range(s.shape[0]):
        a0, b0, a1, b1 = (np.zeros(3, np.float32) for _ in range(4))
        L.orc_voxel_centre_map(_p(s[k]), _p(t[k]), _p(a0), _p(b0))
        header_lib.hdr_centre_map(_p(s[k]), _p(t[k]), _p(a1), _p(b1))
        assert np.array_equal(a0.view(np.uint32), a1.view(np.uint32))
        assert np.array_equal(b0.view(np.uint32), b1.view(np.uint32))
    # a * i + b rounds twice (no FMA)
    a, b = np.float32(0.005000001), np.float32(-0.0524999)
    for i in (0, 1, 7, 210, 4095):
        assert np.float32(header_lib.hdr_centre(a, b, i)) == np.float32(np.float32(a * np.float32(i)) + b)


def test_sse_inverse_routine(built):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libtm_ref.so not built (no /root/reference in this environment)")
    R = C.CDLL(REF)
    L = po.load()
    rng = np.random.default_rng(5)
    # (1) it inverts: random general matrices
    worst = 0.0
    for _ in range(3000):
        m = rng.uniform(-2, 2, size=16).astype(np.float32)
        inv = np.zeros(16, np.float32)
        R.ref_matrix4f_inverse(_p(m), _p(inv))
        M, I = m.reshape(4, 4).T.astype(np.float64), inv.reshape(4, 4).T.astype(np.float64)
        worst = max(worst, np.abs(M @ I - np.eye(4)).max() / (1.0 + np.abs(I).max()))
    assert worst < 1e-5
    # (2) diag + translation: the closed form is the routine
    s, t = _cases(20000, 9)
    for k in range(s.shape[0]):
        m = np.zeros(16, np.float32)
        m[0], m[5], m[10], m[15] = s[k, 0], s[k, 1], s[k, 2], 1.0
        m[12:15] = t[k]
        inv = np.zeros(16, np.float32)
        R.ref_matrix4f_inverse(_p(m), _p(inv))
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        L.orc_voxel_centre_map(_p(s[k]), _p(t[k]), _p(a), _p(b))
        assert np.array_equal(inv[[0, 5, 10]].view(np.uint32), a.view(np.uint32))
        assert np.array_equal(inv[12:15].view(np.uint32), b.view(np.uint32))
        off = np.delete(inv, [0, 5, 10, 12, 13, 14, 15])
        assert not off.any()  # structural zeros stay zeros (either sign)
