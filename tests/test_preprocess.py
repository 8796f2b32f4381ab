"""the pre-processing the reference does with PCL/FLANN — k-NN, principal
curvatures, the tangent criterion (scene.hpp:46-58, model.hpp:68-71,96-99 ->
pointcloud.hpp:3-44,138-152,200-204).
CPU: oracle vs the reference's own principal_curvatures (oracle/_ref) and vs numpy.
GPU (-m gpu): tm_scene_knn / tm_scene_curvature / tm_scene_tangent_mask vs the oracle:
neighbour lists and covariances bit-exact; eigenvalues to 1e-5 relative (device cos/sin vs libm);
masks identical away from the 0.2 ratio boundary."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from oracle import pyoracle as po
from triplet_match_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libtm_ref.so")
F = np.float32


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _noisy(cloud, seed, sigma=0.05):
    """Synthetic clouds have exact normals (degenerate covariance); perturb them like a real estimate."""
    rng = np.random.default_rng(seed)
    nrm = cloud.nrm.astype(np.float64) + sigma * rng.standard_normal(cloud.nrm.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return synth.Cloud(cloud.pos, nrm.astype(F), cloud.tgt, cloud.tangent_mask, cloud.poses)


def test_knn_oracle_vs_numpy(built):
    m = synth.freeform_model(seed=9, n_points=700, radius=0.2)
    q = np.arange(0, m.n, 37, dtype=np.uint32)
    idx, d2 = po.knn(m.pos, q, 30)
    for w, qi in enumerate(q):
        d = m.pos - m.pos[qi]
        dd = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(F) + (d[:, 2] * d[:, 2]).astype(F)).astype(F)
        order = np.lexsort((np.arange(m.n), dd))[:30]
        assert np.array_equal(idx[w], order) and np.array_equal(d2[w], dd[order])
        assert idx[w, 0] == qi  # inclusive
    # k larger than the cloud pads with -1
    idx, _ = po.knn(m.pos[:5], np.array([2], np.uint32), 8)
    assert (idx[0, :5] >= 0).all() and (idx[0, 5:] == -1).all()


def test_eigen33_oracle_vs_numpy(built):
    rng = np.random.default_rng(1)
    for _ in range(500):
        a = rng.standard_normal((3, 3)) * (10 ** rng.uniform(-3, 1))
        cov = (a @ a.T).astype(F)
        if rng.random() < 0.2:
            cov = (np.outer(a[0], a[0]) + 1e-9 * np.eye(3)).astype(F)  # rank-1-ish (projected normals)
        ev = po.eigen33(cov)
        ref = np.linalg.eigvalsh(cov.astype(np.float64))
        assert np.all(np.diff(ev) >= -1e-6 * abs(ref).max())  # roots2 can leave a -1 ulp middle root, as in PCL
        assert np.allclose(ev, ref, rtol=2e-3, atol=2e-3 * max(1e-30, abs(ref).max()))  # float closed form: absolute error ~1e-3 of the largest root
    assert np.array_equal(po.eigen33(np.zeros((3, 3), F)), np.zeros(3, F))


@pytest.mark.parametrize("name", ["plane_small", "cylinder_small", "freeform_small"])
def test_curvature_oracle_equals_reference(built, name):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    L = C.CDLL(REF)
    m, *_ = common.config(name)
    m = _noisy(m, 3)
    q = np.arange(0, m.n, 11, dtype=np.uint32)
    k = 30
    mn_r, mx_r = np.zeros(q.size, F), np.zeros(q.size, F)
    nbr_r = np.zeros((q.size, k), np.int32)
    pos, nrm = np.ascontiguousarray(m.pos, F), np.ascontiguousarray(m.nrm, F)
    L.ref_curvature(_p(pos), _p(nrm), C.c_uint32(m.n), _p(q), C.c_uint32(q.size), C.c_uint32(k), _p(mn_r), _p(mx_r), _p(nbr_r))
    nbr, _ = po.knn(m.pos, q, k)
    assert np.array_equal(nbr, nbr_r)
    mn, mx, cov = po.curvature(m.pos, m.nrm, q, k)
    # the reference's projection / running centroid / covariance order, then the same eigen33 stand-in
    assert np.array_equal(mn.view(np.uint32), mn_r.view(np.uint32))
    assert np.array_equal(mx.view(np.uint32), mx_r.view(np.uint32))
    assert (mx > 0).all() and np.isfinite(mn / mx).all()


def _crease_scene(n_points=30000):
    """Pyramid model (real creases along its edges) + a scene with posed copies: the only synthetic
    clouds here on which the curvature-ratio criterion is meaningful."""
    m = synth.pyramid_model(seed=7, size=0.3, height=0.12, res=0.01)
    s = synth.make_scene(seed=8, model=m, n_points=n_points, n_copies=3, extent=1.2, flat_copies=False)
    return m, s.take(synth.morton_order(s.pos))


def test_tangent_mask_oracle(built):
    m, s2 = _crease_scene()
    mask, cand, mn, mx = po.tangent_mask(s2)
    nrm_t = np.linalg.norm(s2.tgt, axis=1)
    assert np.array_equal(cand, np.flatnonzero(nrm_t.astype(F) > F(0.7)))
    assert mask.sum() > 100 and not mask[nrm_t < 0.7].any()
    # crease points pass, the floor's random "false feature" tangents (flat normals) do not
    mm, *_ = po.tangent_mask(m)
    assert mm.sum() >= 0.9 * m.tangent_mask.sum() and 0 < mask.sum() < cand.size
    # exact normals: projected normals of a plane are all zero -> 0/0 -> never a tangent point (the NaN trap)
    pl = synth.plane_model(seed=2, size=0.2, res=0.01, n_curves=3)
    flat = synth.Cloud(np.concatenate([pl.pos[:, :2], np.zeros((pl.n, 1), F)], 1), pl.nrm, pl.tgt, pl.tangent_mask)
    mk, *_ = po.tangent_mask(flat)
    assert mk.sum() == 0


# ------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx(built):
    from triplet_match_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,morton", [("plane_small", True), ("cylinder_small", False), ("freeform_small", True),
                                         ("crease", True), ("crease", False)])
def test_gpu_knn_curvature_mask(ctx, name, morton):
    from triplet_match_b200 import capi
    if name == "crease":
        m, s = _crease_scene()
    else:
        m, s, *_ = common.config(name)
        s = _noisy(s, 5)
    if not morton:
        s = s.take(synth.shuffle_perm(3, 1, s.n))  # loose segment boxes: still exact
    gs = capi.Scene(ctx, s.pos, s.nrm, s.tgt, s.tangent_mask)
    rng = np.random.default_rng(2)
    q = np.unique(np.concatenate([rng.integers(0, s.n, 300), [0, s.n - 1]])).astype(np.uint32)
    for k in (1, 2, 30, 32):
        gi, gd = gs.knn(q, k)
        oi, od = po.knn(s.pos, q, k)
        assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)), k
    mn, mx, cov = gs.curvature(q, 30)
    omn, omx, ocov = po.curvature(s.pos, s.nrm, q, 30)
    assert np.array_equal(cov.view(np.uint32), ocov.view(np.uint32))
    # eigenvalues bit for bit: pcl::eigen33's cos / sin come from the shared tm_sincosf.h on both sides
    assert np.array_equal(mx.view(np.uint32), omx.view(np.uint32)) and np.array_equal(mn.view(np.uint32), omn.view(np.uint32))
    mask, cnt = gs.compute_tangent_mask(30, 0.2, apply=False)
    omask, cand, c_mn, c_mx = po.tangent_mask(s)
    assert np.array_equal(mask, omask) and cnt == int(mask.sum())  # the whole mask, boundary points included
    if name == "crease":
        assert cnt > 100
    gs.close()


@pytest.mark.gpu
def test_gpu_tangent_mask_apply_feeds_the_search(ctx):
    """apply=1 makes the computed mask the scene's tangent_mask_: pair features then behave as with
    that mask uploaded."""
    from triplet_match_b200 import capi
    m, s, om, osc, rec = common.config("cylinder_small")
    s2 = _noisy(s, 6, sigma=0.3)  # strong normal noise: some points pass the ratio test by chance
    gm = common.upload_model(ctx, m, om)
    ga = capi.Scene(ctx, s2.pos, s2.nrm, s2.tgt, np.zeros(s2.n, np.uint8))
    mask, cnt = ga.compute_tangent_mask(30, 0.2, apply=True)
    gb = capi.Scene(ctx, s2.pos, s2.nrm, s2.tgt, mask)
    fa = ga.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
    fb = gb.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
    assert np.array_equal(fa[2], fb[2]) and np.array_equal(fa[1], fb[1])
    assert cnt == int(mask.sum())
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    ca, _, _ = ga.score(gm, T[:64])
    cb, _, _ = gb.score(gm, T[:64])
    assert np.array_equal(ca, cb)
    ga.close(); gb.close(); gm.close()


@pytest.mark.gpu
def test_gpu_knn_edge_cases(ctx):
    from triplet_match_b200 import capi
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [np.nan, 0, 0], [0, 0, 0]], F)  # NaN point, duplicate
    z = np.zeros_like(pts)
    gs = capi.Scene(ctx, pts, z, z, np.zeros(5, np.uint8))
    gi, gd = gs.knn(np.array([0, 4], np.uint32), 8)
    oi, od = po.knn(pts, np.array([0, 4], np.uint32), 8)
    assert np.array_equal(gi, oi)
    assert gi[0, 0] == 0 and gi[0, 1] == 4 and (gi[0, 4:] == -1).all()  # ties -> lower index; NaN excluded
    with pytest.raises(capi.TmError):
        gs.knn(np.array([9], np.uint32), 4)
    with pytest.raises(capi.TmError):
        gs.knn(np.array([0], np.uint32), 33)
    gs.close()


# ---- Z-curve ordering on the device (tm_scene_upload_sorted) ---------------------------------
def _morton30(pos):
    p = pos.astype(np.float32)
    fin = np.isfinite(p).all(1)
    lo = p[fin].min(0)
    hi = p[fin].max(0)
    d = (hi - lo).astype(np.float32)
    inv = np.where(d > 0, np.float32(1) / np.where(d > 0, d, 1), 0).astype(np.float32)
    t = ((p - lo) * inv).astype(np.float32)
    t = np.where(t >= 0, t, 0)  # also NaN
    t = np.minimum(t, 1).astype(np.float32)
    q = (t * np.float32(1023)).astype(np.uint32)

    def spread(v):
        v = v & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 255, 2048, 2049, 50000, 300001])
def test_gpu_sorted_upload_is_a_stable_morton_sort(ctx, n):
    from triplet_match_b200 import capi
    rng = np.random.default_rng(n)
    pos = rng.random((n, 3)).astype(F) * np.array([3.0, 1.0, 0.2], F)
    if n > 100:
        pos[5] = pos[6]                 # equal codes: stability
        pos[7, 0] = np.nan              # non-finite point: code of the clamped coordinates, still present
    nrm = rng.standard_normal((n, 3)).astype(F)
    tgt = rng.standard_normal((n, 3)).astype(F)
    tm = (rng.random(n) < 0.3).astype(np.uint8)
    gs = capi.Scene(ctx, pos, nrm, tgt, tm, sort=True)
    perm = gs.to_user
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    codes = _morton30(pos)
    exp = np.argsort(codes, kind="stable").astype(np.uint32)
    assert np.array_equal(perm, exp)
    # the resident arrays are the permuted cloud: k-NN of device point d == k-NN of user point perm[d]
    if n >= 255:
        q = rng.integers(0, n, 40).astype(np.uint32)
        gi, gd = gs.knn(q, 4)
        ref = capi.Scene(ctx, pos[perm], nrm[perm], tgt[perm], tm[perm])
        ri, rd = ref.knn(q, 4)
        assert np.array_equal(gi, ri) and np.array_equal(gd.view(np.uint32), rd.view(np.uint32))
        m1, c1 = gs.compute_tangent_mask(8, 0.2, apply=False)
        m2, c2 = ref.compute_tangent_mask(8, 0.2, apply=False)
        assert np.array_equal(m1, m2)
        ref.close()
    gs.close()
