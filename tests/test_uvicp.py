"""Row a15: the orphaned OpenCL ICP path (opencl/icp.cl, cylinder.cl, util.cl).
CPU: the oracle's per-work-item restatement against an independent numpy float32 restatement.
GPU (-m gpu): tm_uvicp_projection / tm_uvicp_correlation against the oracle, bit for bit
(the fused covariance sum: 1e-12 relative, double tree vs sequential double sum)."""
import numpy as np
import pytest

from oracle import pyoracle as po

F = np.float32


def _mm(p, mat):  # util.cl:1-9, p: (n,4) float32, mat: 16 floats column-major
    m = mat.astype(F)
    out = np.empty_like(p)
    for r in range(4):
        out[:, r] = ((m[r] * p[:, 0] + m[4 + r] * p[:, 1]) + m[8 + r] * p[:, 2]) + m[12 + r] * p[:, 3]
    return out


def _setup(seed, n=5000, img=(64, 48), margin=(2, 3), radius=0.5, height=1.2):
    """A cylinder of `radius` along z: model image of uv samples, scene points near the surface."""
    rng = np.random.default_rng(seed)
    w, h = img
    # model "uv image": pixel (x,y) stores the uv of a model sample that falls into it (+ jitter)
    ext = np.array([w - 2 * margin[0] - 1, h - 2 * margin[1] - 1], F)
    image = np.zeros((h, w, 4), F)
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    image[..., 0] = ((xs - margin[0] + rng.random((h, w))) / ext[0]).astype(F)
    image[..., 1] = ((ys - margin[1] + rng.random((h, w))) / ext[1]).astype(F)
    image[..., 3] = 1
    # scene: points on the posed cylinder + noise + clutter
    th = rng.random(n) * 2 * np.pi
    z = rng.random(n) * height
    rr = radius * (1 + 0.02 * rng.standard_normal(n))
    local = np.stack([rr * np.cos(th), rr * np.sin(th), z, np.ones(n)], 1)
    local[: n // 10, :3] = rng.standard_normal((n // 10, 3)) * 2  # clutter
    ang = 0.7
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    pose = np.eye(4)
    pose[:3, :3] = R @ Rx
    pose[:3, 3] = [0.3, -1.0, 2.0]
    pnts = (local @ pose.T).astype(F)
    pnts[:, 3] = 1
    mat_align = np.linalg.inv(pose).T.reshape(-1).astype(F)  # column-major of inverse pose
    proj = np.diag([1 / radius, 1 / radius, 1.0, 1.0])       # cyl2ncoord: radius -> 1
    mat_proj = proj.T.reshape(-1).astype(F)
    norm = np.diag([1.0, 1 / height, 1.0, 1.0])              # v in [0,1]
    mat_norm = norm.T.reshape(-1).astype(F)
    mat_uvw = np.eye(4).T.reshape(-1).astype(F)
    return pnts, image.reshape(-1, 4), np.array(img, np.int32), np.array(margin, np.int32), mat_align, mat_uvw, mat_proj, mat_norm


def _numpy_projection(pnts, image, sz, mg, mat_align, mat_uvw, mat_proj, mat_norm, maxd, atan2f):
    loc = _mm(pnts, mat_align)
    nc = _mm(loc, mat_proj)
    u = (atan2f(nc[:, 1], nc[:, 0]) / F(3.14159274101257324219)).astype(F)
    u = np.where(u < 0, (u + F(2)).astype(F), u)
    u = (u / F(2)).astype(F)
    w = (np.sqrt((nc[:, 0] * nc[:, 0] + nc[:, 1] * nc[:, 1]).astype(F)).astype(F) - F(1)).astype(F)
    uv = np.stack([u, nc[:, 2], w, np.ones_like(u)], 1).astype(F)
    uvn = _mm(_mm(uv, mat_norm), mat_uvw)
    ext = (sz - 2 * mg - 1).astype(F)
    px = np.trunc((uvn[:, :2] * ext).astype(F)).astype(np.int64) + mg
    px[:, 1] = np.where(px[:, 1] == sz[1], sz[1] - 1, px[:, 1])
    n = pnts.shape[0]
    mi = -np.ones(n, np.int32)
    si = -np.ones(n, np.int32)
    op = np.zeros((n, 4), F)
    inb = (px[:, 0] >= 0) & (px[:, 0] < sz[0]) & (px[:, 1] >= 0) & (px[:, 1] < sz[1])
    idx = np.where(inb, px[:, 1] * sz[0] + px[:, 0], 0)
    d = image[idx, :2] - uvn[:, :2]
    dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(F)).astype(F)
    op[inb, 3] = dist[inb]
    hit = inb & (dist < F(maxd))
    mi[hit] = idx[hit]
    si[hit] = np.flatnonzero(hit)
    op[hit] = uvn[hit]
    return op, mi, si


def _atan2f(y, x):
    o, _ = po.atan2f_q1_batch(np.ascontiguousarray(y, F), np.ascontiguousarray(x, F))
    return o


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_projection_vs_numpy(built, seed):
    a = _setup(seed)
    for maxd in (0.02, 0.2):
        op, mi, si, c = po.cl_icp_projection(0, *a, maxd)
        op2, mi2, si2 = _numpy_projection(*a, maxd, _atan2f)
        assert np.array_equal(mi, mi2) and np.array_equal(si, si2)
        assert np.array_equal(op.view(np.uint32), op2.view(np.uint32))
        assert c == (mi >= 0).sum() and 100 < c < a[0].shape[0]


def test_oracle_projection_edge_cases(built):
    pnts, image, sz, mg, *mats = _setup(3, n=64)
    pnts[0, :3] = np.nan
    pnts[1, :3] = 1e30
    pnts[2, :3] = -1e30
    pnts[3, :3] = 0  # on the axis: atan2(0,0), w = -1
    op, mi, si, c = po.cl_icp_projection(0, pnts, image, sz, mg, *mats, 0.5)
    assert mi[0] == -1 and mi[1] == -1 and mi[2] == -1
    # empty input
    op, mi, si, c = po.cl_icp_projection(0, np.zeros((0, 4), F), image, sz, mg, *mats, 0.5)
    assert c == 0 and mi.size == 0


def test_oracle_correlation_vs_numpy(built):
    rng = np.random.default_rng(5)
    scene = rng.standard_normal((3000, 4)).astype(F)
    model = rng.standard_normal((2000, 4)).astype(F)
    n = 1500
    is_ = rng.integers(0, 3000, n).astype(np.int32)
    im_ = rng.integers(0, 2000, n).astype(np.int32)
    cs = np.append(scene[is_, :3].mean(0), 0).astype(F)
    cm = np.append(model[im_, :3].mean(0), 0).astype(F)
    rec, cov = po.cl_icp_correlation(scene, model, is_, im_, cs, cm)
    s = (scene[is_, :3] - cs[:3]).astype(F)
    m = (model[im_, :3] - cm[:3]).astype(F)
    norm = F(1) / F(n - 1)
    exp = np.zeros((n, 16), F)
    for j in range(3):
        for i in range(3):
            exp[:, 3 * j + i] = ((s[:, i] * m[:, j]).astype(F) * norm).astype(F)
    assert np.array_equal(rec.view(np.uint32), exp.view(np.uint32))
    assert np.allclose(cov, exp[:, :9].astype(np.float64).sum(0), rtol=1e-12, atol=1e-15)
    # it is the cross-covariance of the pairs (scene x model), up to float rounding
    ref = (s.astype(np.float64).T @ m.astype(np.float64)) / (n - 1)
    assert np.allclose(cov.reshape(3, 3).T, ref, atol=1e-5)


# ------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx(built):
    from triplet_match_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n", [(1, 5000), (2, 200000), (4, 1)])
def test_gpu_projection_equals_oracle(ctx, seed, n):
    a = _setup(seed, n=max(n, 10))
    a = (a[0][:n],) + a[1:]
    for projector in (0, 1):
        for maxd in (0.02, 0.3):
            op, mi, si, c = ctx.uvicp_projection(projector, *a, maxd)
            op2, mi2, si2, c2 = po.cl_icp_projection(projector, *a, maxd)
            assert c == c2
            assert np.array_equal(mi, mi2) and np.array_equal(si, si2)
            assert np.array_equal(op.view(np.uint32), op2.view(np.uint32))


@pytest.mark.gpu
def test_gpu_projection_edge_cases(ctx):
    pnts, image, sz, mg, *mats = _setup(3, n=64)
    pnts[0, :3] = np.nan
    pnts[1, :3] = 1e30
    pnts[2, :3] = -1e30
    pnts[3, :3] = 0
    pnts[4, :3] = np.inf
    op, mi, si, c = ctx.uvicp_projection(0, pnts, image, sz, mg, *mats, 0.5)
    op2, mi2, si2, c2 = po.cl_icp_projection(0, pnts, image, sz, mg, *mats, 0.5)
    assert c == c2 and np.array_equal(mi, mi2) and np.array_equal(si, si2)
    fin = np.isfinite(op2).all(1)
    assert np.array_equal(op[fin].view(np.uint32), op2[fin].view(np.uint32))
    op, mi, si, c = ctx.uvicp_projection(0, np.zeros((0, 4), F), image, sz, mg, *mats, 0.5)
    assert c == 0 and mi.size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 100000])
def test_gpu_correlation_equals_oracle(ctx, n):
    rng = np.random.default_rng(n + 7)
    scene = rng.standard_normal((30000, 4)).astype(F)
    model = rng.standard_normal((2000, 4)).astype(F)
    is_ = rng.integers(0, 30000, n).astype(np.int32)
    im_ = rng.integers(0, 2000, n).astype(np.int32)
    cs = rng.standard_normal(4).astype(F)
    cm = rng.standard_normal(4).astype(F)
    rec, cov = ctx.uvicp_correlation(scene, model, is_, im_, cs, cm)
    rec2, cov2 = po.cl_icp_correlation(scene, model, is_, im_, cs, cm)
    if n == 1:  # norm = 1/0 = inf: records are +-inf / NaN on both sides
        assert np.array_equal(np.isnan(rec), np.isnan(rec2))
        return
    assert np.array_equal(rec.view(np.uint32), rec2.view(np.uint32))
    assert np.allclose(cov, cov2, rtol=1e-12, atol=1e-18)
    _, cov3 = ctx.uvicp_correlation(scene, model, is_, im_, cs, cm, want_records=False)
    assert np.array_equal(cov, cov3)  # reproducible, records optional


@pytest.mark.gpu
def test_gpu_correlation_rejects_bad_indices(ctx):
    from triplet_match_b200 import capi
    scene = np.zeros((10, 4), F)
    model = np.zeros((5, 4), F)
    with pytest.raises(capi.TmError):
        ctx.uvicp_correlation(scene, model, np.array([0, 10], np.int32), np.array([0, 1], np.int32), np.zeros(4, F), np.zeros(4, F))
