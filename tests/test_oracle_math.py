"""CPU tests of the oracle's primitives against independent restatements (numpy / pure
Python) and against libm.  The reference has no tests or golden vectors."""
import math

import numpy as np
import pytest

from oracle import pyoracle as po


def murmur3_x86_32_ref(data: bytes, seed: int) -> int:
    """Published MurmurHash3_x86_32 (Appleby), pure Python, for 4-byte-multiple inputs."""
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed
    for i in range(0, len(data), 4):
        k = int.from_bytes(data[i:i + 4], "little")
        k = (k * c1) & 0xFFFFFFFF
        k = ((k << 15) | (k >> 17)) & 0xFFFFFFFF
        k = (k * c2) & 0xFFFFFFFF
        h ^= k
        h = ((h << 13) | (h >> 19)) & 0xFFFFFFFF
        h = (h * 5 + 0xE6546B64) & 0xFFFFFFFF
    h ^= len(data)
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def test_murmur_matches_published_algorithm():
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 2**32, size=(500, 4), dtype=np.uint64).astype(np.uint32)
    keys[0] = 0
    keys[1] = 0xFFFFFFFF
    for k in keys:
        assert po.murmur4(k) == murmur3_x86_32_ref(k.tobytes(), 42)


def test_murmur_known_answers():
    # MurmurHash3_x86_32 of 16 zero bytes / of the words 1,2,3,4 with seed 42
    assert po.murmur4([0, 0, 0, 0]) == murmur3_x86_32_ref(bytes(16), 42)
    assert po.murmur4([1, 2, 3, 4]) == 0x3F7F5D44


def test_discretize_edges():
    L = po.load()
    # src/discretize.cpp:19-25: <0 -> 0, >=1 -> steps-1, else trunc(nval*steps)
    assert L.orc_discretize_range(-1.0, 0.0, 2.0, 20) == 0
    assert L.orc_discretize_range(2.0, 0.0, 2.0, 20) == 19
    assert L.orc_discretize_range(5.0, 0.0, 2.0, 20) == 19
    assert L.orc_discretize_range(0.999, 0.0, 2.0, 20) == 9
    assert L.orc_discretize_range(1.0, 0.0, 2.0, 20) == 10
    # src/discretize.cpp:27-30
    assert L.orc_discretize_step(0.0, 0.17453292) == 0
    assert L.orc_discretize_step(1.5707964, 0.17453292) == 9
    rng = np.random.default_rng(0)
    v = rng.random(2000).astype(np.float32) * 3 - 0.5
    for x in v:
        nval = (np.float32(x) - np.float32(0.25)) / np.float32(1.75)
        exp = 0 if nval < 0 else (19 if nval >= 1 else int(np.float32(nval * np.float32(20))))
        assert L.orc_discretize_range(float(x), 0.25, 1.75, 20) == exp


def test_atan2f_restatement_equals_libm():
    """The oracle's atan2f restates glibc 2.39's binary32 algorithm (the reference platform's
    libm, include/impl/feature.hpp:7).  It was pinned against this image's libm on 10^8 inputs;
    this test repeats the comparison on the host it runs on."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    y = np.abs(rng.standard_normal(n)).astype(np.float32)
    x = np.abs(rng.standard_normal(n)).astype(np.float32)
    y[:100] = 0
    x[100:200] = 0
    y[200:300] *= 1e-20
    x[300:400] *= 1e-20
    y[400:500] = np.inf
    x[500:600] = np.inf
    x[600:700] = 1.0
    # raw bit patterns (every exponent, denormals, NaN) for the first quadrant
    yb = rng.integers(0, 0x7FC00001, size=n // 4, dtype=np.int64).astype(np.uint32).view(np.float32)
    xb = rng.integers(0, 0x7FC00001, size=n // 4, dtype=np.int64).astype(np.uint32).view(np.float32)
    y, x = np.concatenate([y, yb]), np.concatenate([x, xb])
    ours, libm = po.atan2f_q1_batch(y, x)
    nan = np.isnan(ours) & np.isnan(libm)
    same = (ours.view(np.uint32) == libm.view(np.uint32)) | nan
    assert same.all(), f"{int((~same).sum())} of {y.size} inputs differ from this host's libm atan2f"
    # accuracy: within 1 ulp of the correctly rounded value
    fin = np.isfinite(ours) & np.isfinite(x) & np.isfinite(y)
    cr = np.arctan2(y[fin].astype(np.float64), x[fin].astype(np.float64)).astype(np.float32)
    d = np.abs(ours[fin].view(np.int32).astype(np.int64) - cr.view(np.int32).astype(np.int64))
    assert d.max() <= 1


def test_atan2f_all_quadrants_equals_libm():
    rng = np.random.default_rng(1)
    n = 500000
    y = rng.standard_normal(n).astype(np.float32)
    x = rng.standard_normal(n).astype(np.float32)
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 3e38, -3e38, 1e-30], np.float32)
    yy, xx = np.meshgrid(sp, sp)
    y, x = np.concatenate([y, yy.ravel()]), np.concatenate([x, xx.ravel()])
    ours, libm = po.atan2f_q1_batch(y, x)  # atan2f_q1 is the full four-quadrant restatement
    nan = np.isnan(ours) & np.isnan(libm)
    same = (ours.view(np.uint32) == libm.view(np.uint32)) | nan
    assert same.all(), f"{int((~same).sum())} of {y.size} inputs differ from this host's libm atan2f"


def _upper_x86(tried, nsub, corrs):
    """The reference expression `uint32_t upper = -1.0 - static_cast<uint32_t>((x*n+tmp)/N)`
    as x86-64/gcc evaluates it (cvttsd2si to 64 bit, low 32 bits)."""
    N = -2.0 - tried
    x = -2.0 - nsub
    n = -1.0 - corrs
    tmp = math.sqrt((x * n * (N - x) * (N - n)) / (N - 1.0))
    v = (x * n + tmp) / N
    a = int(v) & 0xFFFFFFFF  # trunc toward zero, wrap
    b = -1.0 - float(a)
    return int(b) & 0xFFFFFFFF


def test_sincosf_small_is_correctly_rounded_and_close_to_libm():
    """tm_sincosf.h (shared by device, host and oracle for pcl::eigen33's cos / sin of theta in [0, pi/3]): equal to the
    binary64 libm value rounded once — i.e. the correctly rounded binary32 result — on every sampled input (exhaustively
    verified for all 1.07e9 floats of [0, 1.6] when the header was written), and within 1 ulp of the host's sinf / cosf."""
    import ctypes as C
    L = po.load()
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(0.0, 1.6, 2_000_000), rng.uniform(0.0, 1.0472, 1_000_000),
                        [0.0, 1e-30, 1e-8, 0.5, 1.0, 1.0471976, 1.5707964, 1.6]]).astype(np.float32)
    s, c, sl, cl = (np.zeros(x.size, np.float32) for _ in range(4))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.orc_sincosf_small_batch(p(x), C.c_uint64(x.size), p(s), p(c), p(sl), p(cl))
    assert np.array_equal(s.view(np.uint32), np.sin(x.astype(np.float64)).astype(np.float32).view(np.uint32))
    assert np.array_equal(c.view(np.uint32), np.cos(x.astype(np.float64)).astype(np.float32).view(np.uint32))
    ds = np.abs(s.view(np.int32).astype(np.int64) - sl.view(np.int32))
    dc = np.abs(c.view(np.int32).astype(np.int64) - cl.view(np.int32))
    assert ds.max() <= 1 and dc.max() <= 1
    assert (ds != 0).mean() < 5e-2 and (dc != 0).mean() < 5e-2  # glibc is within 0.56 ulp, not correctly rounded


def test_early_drop_bound_restatement():
    L = po.load()
    rng = np.random.default_rng(3)
    for _ in range(3000):
        nsub = int(rng.integers(1, 200000))
        tried = int(rng.integers(1, nsub + 1))
        corrs = int(rng.integers(0, tried + 1))
        assert L.orc_early_drop_upper(tried, nsub, corrs) == _upper_x86(tried, nsub, corrs)
    # survey claim: upper == trunc(|v|) - 1 whenever trunc(|v|) >= 1
    assert L.orc_early_drop_upper(259, 5191, 0) == 38


def test_early_drop_checkpoints():
    for nsub in (0, 1, 7, 10, 19, 20, 21, 100, 5191, 65536, 1000003):
        exp = [int(np.float32(np.float32(np.float32(0.05) * np.float32(i + 1)) * np.float32(nsub)))
               for i in range(18)]
        assert po.early_drop_tests(nsub).tolist() == exp


def test_feature_against_float64():
    rng = np.random.default_rng(5)
    for _ in range(300):
        p0, p1 = rng.standard_normal(3), rng.standard_normal(3)
        t0 = rng.standard_normal(3); t0 /= np.linalg.norm(t0)
        t1 = rng.standard_normal(3); t1 /= np.linalg.norm(t1)
        f = po.feature(p0, t0, p1, t1)
        p0f, p1f, t0f, t1f = [np.float32(a).astype(np.float64) for a in (p0, p1, t0, t1)]
        d = p1f - p0f
        ang = lambda a, b: math.atan2(np.linalg.norm(np.cross(a, b)), abs(a @ b))
        exp = [np.linalg.norm(d), ang(d, t0f), ang(d, t1f), np.linalg.norm(d)]
        assert np.allclose(f, exp, rtol=2e-6, atol=2e-6)
        assert f[3] == f[0] and 0 <= f[1] <= math.pi / 2 + 1e-6


def test_base_transform_maps_frames():
    rng = np.random.default_rng(7)
    for _ in range(200):
        si, sj, ti, tj = (rng.standard_normal(3) for _ in range(4))
        st, tt = rng.standard_normal(3), rng.standard_normal(3)
        T = po.base_transform(si, sj, st, ti, tj, tt).reshape(4, 4).T.astype(np.float64)
        R, t = T[:3, :3], T[:3, 3]
        si32, sj32 = np.float32(si).astype(np.float64), np.float32(sj).astype(np.float64)
        ti32, tj32 = np.float32(ti).astype(np.float64), np.float32(tj).astype(np.float64)
        assert np.allclose(R @ si32 + t, ti32, atol=2e-5)            # origin -> origin
        assert np.allclose(R.T @ R, np.eye(3), atol=2e-5)            # rotation
        ua = (sj32 - si32) / np.linalg.norm(sj32 - si32)
        ub = (tj32 - ti32) / np.linalg.norm(tj32 - ti32)
        assert np.allclose(R @ ua, ub, atol=2e-5)                    # pair direction aligned
        assert np.array_equal(T[3], [0, 0, 0, 1])


def test_umeyama_recovers_rigid_motion():
    rng = np.random.default_rng(9)
    for _ in range(20):
        src = rng.standard_normal((200, 3))
        ax = rng.standard_normal(3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(-3, 3)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
        t = rng.standard_normal(3)
        dst = src @ R.T + t
        T = po.umeyama(src, dst).reshape(4, 4).T
        assert np.allclose(T[:3, :3], R, atol=1e-5) and np.allclose(T[:3, 3], t, atol=1e-5)


def test_octant_rule():
    L = po.load()
    c = np.zeros(3, dtype=np.float32)
    for bits in range(8):
        p = np.array([1 if bits & 1 else -1, 1 if bits & 2 else -1, 1 if bits & 4 else -1], dtype=np.float32)
        assert L.orc_get_octant(c.ctypes.data_as(po.C.c_void_p), p.ctypes.data_as(po.C.c_void_p)) == bits
    # pos == center goes to the low octant (strict '>', include/impl/octree.hpp:11-17)
    assert L.orc_get_octant(c.ctypes.data_as(po.C.c_void_p), c.ctypes.data_as(po.C.c_void_p)) == 0
