"""Pins rows a1-a4 of the oracle against the REFERENCE's own sources: src/discretize.cpp,
include/impl/discretize.hpp and include/impl/feature.hpp compiled where they lie against the
header stand-ins of oracle/shim/ (recipe: oracle/Makefile target `ref` -> oracle/_ref/).
The shared library is built in the authoring container (where /root/reference exists) and
travels with the snapshot; the test skips when it is absent."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libtm_ref.so")


@pytest.fixture(scope="module")
def ref(built):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libtm_ref.so not built (no /root/reference in this environment)")
    L = C.CDLL(REF)
    L.ref_murmur4.restype = C.c_uint32
    L.ref_std_hash4.restype = C.c_uint64
    L.ref_discretize_range.restype = C.c_uint32
    L.ref_discretize_range.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32]
    L.ref_discretize_step.restype = C.c_uint32
    L.ref_discretize_step.argtypes = [C.c_float, C.c_float]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_murmur_and_hash(ref):
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2**32, size=(3000, 4), dtype=np.uint64).astype(np.uint32)
    keys[:20] = rng.integers(0, 20, size=(20, 4))  # realistic small keys
    for k in keys:
        h = ref.ref_murmur4(_p(k))
        assert h == po.murmur4(k)
        assert ref.ref_std_hash4(_p(k)) == h  # std::hash widens the 32-bit murmur to size_t


def test_discretize(ref):
    L = po.load()
    rng = np.random.default_rng(1)
    for _ in range(4000):
        v, mn, rg = (np.float32(x) for x in rng.standard_normal(3) * 2)
        rg = np.float32(abs(rg) + 1e-3)
        steps = int(rng.integers(1, 64))
        assert ref.ref_discretize_range(v, mn, rg, steps) == L.orc_discretize_range(v, mn, rg, steps)
        a, st = np.float32(abs(v)), np.float32(abs(mn) + 0.01)
        assert ref.ref_discretize_step(a, st) == L.orc_discretize_step(a, st)


def test_feature_valid_discretize_feature(ref):
    rng = np.random.default_rng(2)
    for _ in range(3000):
        p0, p1 = rng.standard_normal(3), rng.standard_normal(3)
        t0 = rng.standard_normal(3); t0 /= np.linalg.norm(t0)
        t1 = rng.standard_normal(3); t1 /= np.linalg.norm(t1)
        if rng.random() < 0.1:
            t1 = t0  # parallel tangents
        if rng.random() < 0.05:
            t0 = (p1 - p0) / np.linalg.norm(p1 - p0)  # tangent along the pair direction
        inp = np.concatenate([p0, t0, p1, t1]).astype(np.float32)
        f_ref = np.zeros(4, dtype=np.float32)
        ref.ref_feature(_p(inp), _p(f_ref))
        # the reference calls libm atan2f; the oracle in libm mode must agree bit for bit
        f_orc = po.feature(inp[0:3], inp[3:6], inp[6:9], inp[9:12], use_libm=True)
        assert np.array_equal(f_ref.view(np.uint32), f_orc.view(np.uint32))
        # ... and so must the restated atan2f every kernel shares (glibc's binary32 algorithm)
        f_sw = po.feature(inp[0:3], inp[3:6], inp[6:9], inp[9:12])
        assert np.array_equal(f_sw.view(np.uint32), f_ref.view(np.uint32))
        mn = np.array([0.3, 0, 0, 0.3], dtype=np.float32) * np.float32(rng.random() + 0.5)
        mx = mn + np.array([2.5, 3.2, 3.2, 2.5], dtype=np.float32) * np.float32(rng.random() + 0.2)
        # oracle side through the same public helpers the pipeline uses
        nb_mn, nb_mx = np.zeros(4, np.float32), np.zeros(4, np.float32)
        ref.ref_valid_bounds(_p(mn), _p(mx), C.c_float(0.0), C.c_float(1.0), _p(nb_mn), _p(nb_mx))
        d0 = np.float32(mx[0] - mn[0]); d3 = np.float32(mx[3] - mn[3])
        assert nb_mn[0] == np.float32(mn[0] + np.float32(0.0) * d0) and nb_mx[0] == np.float32(mn[0] + np.float32(1.0) * d0)
        assert nb_mx[3] == np.float32(mn[3] + np.float32(1.0) * d3) and nb_mn[1] == mn[1] and nb_mx[2] == mx[2]
        v_ref = ref.ref_valid(_p(f_ref), _p(mn), _p(mx))
        v_exp = int((mn[0] <= f_ref[0] <= mx[0]) and 0 <= f_ref[1] <= np.float32(np.pi) and 0 <= f_ref[2] <= np.float32(np.pi))
        assert v_ref == v_exp
        key = np.zeros(4, dtype=np.uint32)
        ref.ref_discretize_feature(_p(f_ref), _p(mn), _p(mx), C.c_float(20.0), C.c_float(0.17453292), _p(key))
        L = po.load()
        exp = [L.orc_discretize_range(f_ref[0], mn[0], d0, 20), L.orc_discretize_step(f_ref[1], 0.17453292),
               L.orc_discretize_step(f_ref[2], 0.17453292), L.orc_discretize_range(f_ref[3], mn[0], d0, 20)]
        assert key.tolist() == exp


# ---- rows a5-a12: the reference's model::init / query / voxel_query and scene::impl::
# base_transform_ / project_ / icp_, compiled from /root/reference against oracle/shim ----
import common  # noqa: E402

CONFIGS = ["plane_small", "cylinder_small", "freeform_small"]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefModel:
    def __init__(self, L, m):
        L.ref_model_create.restype = C.c_void_p
        L.ref_model_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_float,
                                       C.c_float, C.c_float]
        L.ref_model_query.restype = C.c_uint32
        self.L = L
        self.pos, self.nrm, self.tgt = _f32(m.pos), _f32(m.nrm), _f32(m.tgt)
        self.h = C.c_void_p(L.ref_model_create(_p(self.pos), _p(self.nrm), _p(self.tgt), m.n, 20.0, 0.17453292, 0.2, 1.0))
        f10, tv, i5 = np.zeros(10, np.float32), np.zeros(16, np.float32), np.zeros(5, np.int32)
        L.ref_model_info(self.h, _p(f10), _p(tv), _p(i5))
        self.resolution, self.diameter = f10[0], f10[1]
        self.feat_min, self.feat_max = f10[2:6].copy(), f10[6:10].copy()
        self.to_voxel16, self.extents, self.margin, self.point_count = tv, i5[:3].copy(), int(i5[3]), int(i5[4])

    def query(self, f, limit=200):
        out = np.zeros((max(limit, 1), 2), dtype=np.uint32)
        n = self.L.ref_model_query(self.h, _p(_f32(f)), C.c_uint32(limit), _p(out))
        return out[:n]

    def voxel_query(self, p4):
        o = C.c_uint32()
        return int(o.value) if self.L.ref_model_voxel_query(self.h, _p(_f32(p4)), C.byref(o)) else None


class RefScene:
    def __init__(self, L, s, mask=None):
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p]
        L.ref_project.restype = C.c_uint32
        L.ref_icp.restype = C.c_uint32
        self.L = L
        self.pos, self.nrm, self.tgt = _f32(s.pos), _f32(s.nrm), _f32(s.tgt)
        tm = np.ascontiguousarray(s.tangent_mask, dtype=np.uint8)
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.h = C.c_void_p(L.ref_scene_create(_p(self.pos), _p(self.nrm), _p(self.tgt), s.n, _p(tm),
                                               None if mk is None else _p(mk)))

    def project(self, rm, subset, T16, accept=0.5, dist_thres=1.0, early_out=False):
        sub = np.ascontiguousarray(subset, dtype=np.int32)
        sc, mc = np.zeros(max(sub.size, 1), np.uint32), np.zeros(max(sub.size, 1), np.uint32)
        score, saved = C.c_double(), C.c_uint32()
        n = self.L.ref_project(self.h, rm.h, _p(sub), C.c_uint64(sub.size), _p(_f32(T16)), C.c_float(accept),
                               C.c_float(dist_thres), C.c_int(int(early_out)), _p(sc), _p(mc), C.byref(score),
                               C.byref(saved))
        return dict(count=int(n), scene_corrs=sc[:n].copy(), model_corrs=mc[:n].copy(), score=score.value,
                    saved=int(saved.value))


@pytest.fixture(scope="module", params=CONFIGS)
def refcfg(request, ref):
    m, s, om, osc, rec = common.config(request.param)
    rm = RefModel(ref, m)
    rs = RefScene(ref, s)
    return request.param, m, s, om, osc, rec, rm, rs


def test_reference_model_init(refcfg):
    name, m, s, om, osc, rec, rm, rs = refcfg
    assert np.float32(rm.resolution) == np.float32(om.resolution)
    assert np.float32(rm.diameter) == np.float32(om.diameter)
    assert np.array_equal(rm.extents, om.extents) and rm.margin == om.margin
    assert np.array_equal(rm.to_voxel16.view(np.uint32), om.to_voxel16.view(np.uint32))
    assert np.array_equal(rm.feat_min.view(np.uint32), om.feat_min.view(np.uint32))
    assert np.array_equal(rm.feat_max.view(np.uint32), om.feat_max.view(np.uint32))
    assert rm.point_count == om.n_subset
    # voxel grid through the reference's voxel_query at sampled cells (+ outside positions).  The
    # reference places voxel centres with Matrix4f::inverse() (model.hpp:63,87; the shim restates Eigen's
    # SSE cofactor routine lane by lane), the oracle with that routine's closed form for diag + translation:
    # the grids must be identical, near-ties included.
    ex = om.extents.astype(np.int64)
    rng = np.random.default_rng(0)
    cells = rng.choice(int(ex.prod()), size=min(4000, int(ex.prod())), replace=False)
    for lin in cells:
        k, r = divmod(int(lin), int(ex[0] * ex[1]))
        j, i = divmod(r, int(ex[0]))
        centre = (np.array([i, j, k], np.float32) + np.float32(0.25) - om.trans) / om.scale
        got = rm.voxel_query([centre[0], centre[1], centre[2], 1.0])
        exp = om.voxel_query([centre[0], centre[1], centre[2], 1.0])
        assert got == exp
    for p in ([1e3, 0, 0, 1], [-1e3, 0, 0, 1], [np.nan, 0, 0, 1]):
        assert rm.voxel_query(p) is None and om.voxel_query(p) is None


def test_reference_model_init_with_subset(ref):
    """model::init(subset, params) (model.hpp:16-61): bounding box, diameter, grid geometry and the tangent subset come
    from the caller's subset, the nearest-neighbour grid from the whole cloud.  Reference == oracle == the product's host
    build (CPU path, no device), incl. the grid and the hash table."""
    import common
    from triplet_match_b200 import capi
    m, s, om_full, osc, rec = common.config("cylinder_small")
    sub = np.flatnonzero(m.pos[:, 2] > np.median(m.pos[:, 2])).astype(np.uint32)  # the upper half of the cylinder
    ref.ref_model_create_subset.restype = C.c_void_p
    ref.ref_model_create_subset.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_float] * 4
    pos, nrm, tgt = _f32(m.pos), _f32(m.nrm), _f32(m.tgt)
    h = C.c_void_p(ref.ref_model_create_subset(_p(pos), _p(nrm), _p(tgt), m.n, _p(sub), sub.size, 20.0, 0.17453292, 0.2, 1.0))
    f10, tv, i5 = np.zeros(10, np.float32), np.zeros(16, np.float32), np.zeros(5, np.int32)
    ref.ref_model_info(h, _p(f10), _p(tv), _p(i5))
    om = po.OModel(m, subset=sub)
    assert np.float32(om.diameter) == f10[1] and np.float32(om.diameter) < np.float32(om_full.diameter)
    assert np.array_equal(om.extents, i5[:3]) and om.n_subset == int(i5[4]) and 0 < om.n_subset < om_full.n_subset
    assert np.array_equal(om.to_voxel16.view(np.uint32), tv.view(np.uint32))
    assert np.array_equal(om.feat_min.view(np.uint32), f10[2:6].view(np.uint32))
    assert np.array_equal(om.feat_max.view(np.uint32), f10[6:10].view(np.uint32))
    ref.ref_model_voxels.restype = C.c_uint64
    grid = np.zeros(int(np.prod(om.extents.astype(np.int64))), np.uint32)
    assert ref.ref_model_voxels(h, _p(grid)) == 0
    assert np.array_equal(grid, om.voxel)
    ref.ref_model_destroy(h)
    hm = capi.HostModel(None, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, subset=sub, resolution=om.resolution)
    assert np.float32(hm.diameter) == np.float32(om.diameter) and np.array_equal(hm.extents, om.extents)
    assert np.array_equal(hm.to_voxel16.view(np.uint32), om.to_voxel16.view(np.uint32))
    assert np.array_equal(hm.voxel, om.voxel) and hm.n_subset == om.n_subset
    keys, offsets, pairs = om.table(200)
    assert np.array_equal(hm.keys, keys) and np.array_equal(hm.offsets, offsets) and np.array_equal(hm.pairs, pairs)
    hm.close()


def test_reference_query_order_and_limit(refcfg):
    """Hash-hit order of the reference's own unordered_multimap + std::hash + query_limit loop."""
    name, m, s, om, osc, rec, rm, rs = refcfg
    feats, keys, valid = osc.pair_features(om, rec.pair_i, rec.pair_j)
    n = 0
    for f in feats[valid.astype(bool)][:60]:
        for limit in (200, 7):
            a, b = rm.query(f, limit), om.query(f, limit)
            assert np.array_equal(a, b)
            n += a.shape[0]
    assert n > 100


def test_reference_base_transform(refcfg):
    name, m, s, om, osc, rec, rm, rs = refcfg
    rng = np.random.default_rng(3)
    T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    for h in rng.choice(T.shape[0], size=min(300, T.shape[0]), replace=False):
        i, j = rec.pair_i[hp[h]], rec.pair_j[hp[h]]
        inp = _f32(np.concatenate([s.pos[i], s.pos[j], s.tgt[i], m.pos[mi[h]], m.pos[mj[h]], m.tgt[mi[h]]]))
        out = np.zeros(16, np.float32)
        rs.L.ref_base_transform(rs.h, _p(inp), _p(out))
        assert np.array_equal(out.view(np.uint32), T[h].view(np.uint32))
    for _ in range(200):  # arbitrary (incl. degenerate) inputs
        inp = _f32(rng.standard_normal(18))
        if rng.random() < 0.1:
            inp[6:9] = inp[3:6] - inp[0:3]  # tangent parallel to the pair direction -> singular frame
        out = np.zeros(16, np.float32)
        rs.L.ref_base_transform(rs.h, _p(inp), _p(out))
        exp = po.base_transform(inp[0:3], inp[3:6], inp[6:9], inp[9:12], inp[12:15], inp[15:18])
        assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))


def test_reference_project(refcfg):
    name, m, s, om, osc, rec, rm, rs = refcfg
    T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T, nthreads=4)
    rng = np.random.default_rng(4)
    pick = np.unique(np.concatenate([np.argsort(-cnt.astype(np.int64))[:12],
                                     rng.choice(T.shape[0], size=min(40, T.shape[0]), replace=False)]))
    for h in pick:
        sub = osc.ball_subset(int(rec.outer[rec.pair_outer[hp[h]]]), om.diameter)
        for eo in (False, True):
            a = rs.project(rm, sub, T[h], early_out=eo)
            b = osc.project(om, sub, T[h], early_out=eo)
            assert a["count"] == b["count"], (name, int(h), eo)
            assert np.array_equal(a["scene_corrs"], b["scene_corrs"])
            assert np.array_equal(a["model_corrs"], b["model_corrs"])
            assert a["score"] == b["score"] and a["saved"] == b["saved"]
    # early-drop on short / shuffled subsets (checkpoint chaining, UB casts as compiled by gcc)
    sub_full = osc.ball_subset(int(rec.outer[0]), om.diameter)
    perm = rng.permutation(sub_full)
    for n in (0, 1, 2, 7, 19, 20, 21, 40, 333, len(perm)):
        for h in pick[:6]:
            a = rs.project(rm, perm[:n], T[h], early_out=True)
            b = osc.project(om, perm[:n], T[h], early_out=True)
            assert (a["count"], a["saved"], a["score"]) == (b["count"], b["saved"], b["score"]), (n, int(h))


def test_reference_project_with_mask(refcfg, ref):
    name, m, s, om, osc, rec, rm, rs = refcfg
    rng = np.random.default_rng(5)
    mask = (rng.random(s.n) < 0.25).astype(np.uint8)
    rs2 = RefScene(ref, s, mask)
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T, nthreads=4)
    osc.set_mask(mask)
    try:
        for h in np.argsort(-cnt.astype(np.int64))[:10]:
            sub = osc.ball_subset(int(rec.outer[rec.pair_outer[hp[h]]]), om.diameter)
            for eo in (False, True):
                a, b = rs2.project(rm, sub, T[h], early_out=eo), osc.project(om, sub, T[h], early_out=eo)
                assert (a["count"], a["saved"], a["score"]) == (b["count"], b["saved"], b["score"])
    finally:
        osc.set_mask(np.zeros(s.n, dtype=np.uint8))


def test_reference_icp_control_flow(refcfg):
    """icp_ loop / stop rule / 2*dist_thres of the reference vs the oracle (the rigid solve is
    the same stand-in on both sides, so counts and poses must agree exactly)."""
    name, m, s, om, osc, rec, rm, rs = refcfg
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T, nthreads=4)
    for h in np.argsort(-cnt.astype(np.int64), kind="stable")[:3]:
        for iters in (1, 5):
            out, score = np.zeros(16, np.float32), C.c_double()
            n = rs.L.ref_icp(rs.h, rm.h, _p(_f32(T[h])), C.c_uint32(iters), C.c_float(1.0), C.c_float(0.5), _p(out),
                             C.byref(score))
            oT, on, osx, oit = osc.icp(om, T[h], iters, 1.0)
            assert n == on and np.array_equal(out.view(np.uint32), oT.view(np.uint32))
            assert score.value == osx


def test_reference_resolution(ref):
    from triplet_match_b200 import synth
    ref.ref_resolution.restype = C.c_float
    c = synth.freeform_model(seed=9, n_points=400, radius=0.2)
    pos = _f32(c.pos)
    assert np.float32(ref.ref_resolution(_p(pos), C.c_uint32(c.n))) == np.float32(
        po.load().orc_resolution(_p(pos), C.c_uint32(c.n)))


# ---- row a14: the traits' closed forms, compiled from the reference's impl/*_traits.hpp ----
def _traits_cases():
    rng = np.random.default_rng(8)
    for kind in (0, 1, 2, 3):
        for trial in range(4):
            q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            g = np.eye(4)
            g[:3, :3] = q
            g[:3, 3] = rng.standard_normal(3)
            g16 = _f32(g.T.reshape(-1))  # column-major
            l16 = _f32(np.linalg.inv(g).T.reshape(-1))
            radius, thr = np.float32(0.4 + 0.2 * trial), np.float32(0.15)
            pts = rng.standard_normal((400, 3)).astype(np.float32)
            if kind == 0:  # points near the cylinder surface (some inside the threshold, some not), all four quadrants
                loc = np.stack([np.cos(pts[:, 0] * 3) * (radius + 0.2 * pts[:, 1]), np.sin(pts[:, 0] * 3) * (radius + 0.2 * pts[:, 1]), pts[:, 2], np.ones(400)], 1)
                pts = (loc @ np.linalg.inv(g).T)[:, :3].astype(np.float32)
                pts[:4] = (np.array([[radius, 0, 0, 1], [-radius, 0, 0, 1], [0, radius, 0, 1], [0, -radius, 0, 1]]) @ np.linalg.inv(g).T)[:, :3].astype(np.float32)
            elif kind == 1:
                pts[:, :] = ((np.concatenate([pts[:, :2], 0.2 * pts[:, 2:3], np.ones((400, 1))], 1)) @ np.linalg.inv(g).T)[:, :3].astype(np.float32)
            nrm = rng.standard_normal((400, 3))
            nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
            tgt = np.cross(nrm, rng.standard_normal((400, 3)))
            tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
            yield kind, g16, l16, radius, thr, pts, _f32(nrm), _f32(tgt)


def _ref_traits_rows(ref, kind, g16, l16, radius, thr, pts, nrm, tgt):
    ref.ref_traits.restype = C.c_int
    rows = np.zeros((pts.shape[0], 14), np.float32)
    for i in range(pts.shape[0]):
        out = np.zeros(15, np.float32)
        ok = ref.ref_traits(C.c_int(kind), _p(g16), _p(l16), C.c_float(radius), C.c_float(thr), _p(pts[i]), _p(nrm[i]), _p(tgt[i]), _p(out))
        rows[i, 0] = ok
        rows[i, 1:] = out[:13]
    return rows


def test_reference_traits_project(ref):
    """oracle restatement of the four `project` closed forms (what tm_traits_project is tested against)."""
    for kind, g16, l16, radius, thr, pts, nrm, tgt in _traits_cases():
        uvw_o, ok_o = po.traits_project(kind, g16, radius, thr, pts)
        rows = _ref_traits_rows(ref, kind, g16, l16, radius, thr, pts, nrm, tgt)
        assert np.array_equal(rows[:, 0] != 0, ok_o != 0), kind
        sel = ok_o != 0
        assert np.array_equal(rows[sel, 1:4].view(np.uint32), uvw_o[sel].view(np.uint32)), kind
        assert sel.sum() > 50 and (kind >= 2 or sel.sum() < pts.shape[0])


def test_reference_traits_dropin_headers(ref, tmp_path):
    """include/triplet_match/*_traits (host code of the drop-in) against the reference's own
    project / unproject / tangent / normal / intrinsic_distance, bit for bit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "triplet_match_b200")
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_dropin.cpp"), "-o", exe, "-L" + libdir,
                           "-ltriplet_match_b200", "-Wl,-rpath," + libdir])
    names = ["ok", "u", "v", "w", "back_x", "back_y", "back_z", "tan_x", "tan_y", "tan_z", "nrm_x", "nrm_y", "nrm_z", "dist"]
    for c, (kind, g16, l16, radius, thr, pts, nrm, tgt) in enumerate(_traits_cases()):
        inp, outp = str(tmp_path / f"t{c}.in"), str(tmp_path / f"t{c}.out")
        with open(inp, "wb") as f:
            f.write(np.int32(kind).tobytes() + g16.tobytes() + l16.tobytes() + np.float32(radius).tobytes() + np.float32(thr).tobytes())
            f.write(np.uint32(pts.shape[0]).tobytes())
            f.write(np.ascontiguousarray(np.concatenate([pts, nrm, tgt], axis=1), dtype=np.float32).tobytes())
        subprocess.check_call([exe, "traits", inp, outp])
        got = np.fromfile(outp, dtype=np.float32).reshape(-1, 14)
        want = _ref_traits_rows(ref, kind, g16, l16, radius, thr, pts, nrm, tgt)
        for col in range(14):
            bad = np.nonzero(got[:, col].view(np.uint32) != want[:, col].view(np.uint32))[0]
            assert bad.size == 0, (kind, names[col], bad[:5], got[bad[:5], col], want[bad[:5], col])


# ---- row a16: octree build + traversals, compiled from include/octree(.ipp) + impl/octree.hpp ----
def test_reference_octree_dropin_headers(ref, tmp_path):
    """include/triplet_match/octree against the reference's own tree: same nodes (depth, kind, box bits,
    leaf contents) in the same order for all five traversals, three criteria, with and without subset."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "triplet_match_b200")
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_dropin.cpp"), "-o", exe, "-L" + libdir,
                           "-ltriplet_match_b200", "-Wl,-rpath," + libdir])
    ref.ref_octree.restype = C.c_uint32
    rng = np.random.default_rng(16)
    pts = rng.standard_normal((3000, 3)).astype(np.float32)
    pts[:50] = pts[50:100]  # duplicates: a max_point_count tree bottoms out on the depth limit
    pts[100:140, 0] = 0.25  # points on a splitting plane go to the low octant (strict '>')
    one = pts[:1].copy()
    cases = []
    for crit_kind, crit_value, max_depth in ((2, 20.0, 6), (2, 1.0, 3), (1, 0.8, 8), (0, 0.3, 8), (2, 5000.0, 4)):
        for subset in (None, rng.permutation(3000)[:700].astype(np.uint32)):
            for trav, level in ((0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (4, 2), (4, 9)):
                cases.append((pts, subset, max_depth, crit_kind, crit_value, trav, level))
    cases.append((one, None, 4, 2, 8.0, 3, 0))  # a single leaf: branch_traverse yields the root leaf (reference quirk)
    cases.append((one, None, 4, 2, 8.0, 2, 0))
    n_nodes = 0
    for c, (p, subset, max_depth, crit_kind, crit_value, trav, level) in enumerate(cases):
        cap = 1 << 16
        rows = np.zeros((cap, 12), np.float64)
        depth = C.c_uint32()
        sub_p = _p(subset) if subset is not None else None
        n = ref.ref_octree(_p(p), C.c_uint32(p.shape[0]), sub_p, C.c_uint32(0 if subset is None else subset.size),
                           C.c_uint32(max_depth), C.c_int(crit_kind), C.c_float(crit_value), C.c_int(trav),
                           C.c_uint32(level), _p(rows), C.c_uint32(cap), C.byref(depth))
        assert n <= cap
        inp, outp = str(tmp_path / f"o{c}.in"), str(tmp_path / f"o{c}.out")
        with open(inp, "wb") as f:
            f.write(np.uint32(p.shape[0]).tobytes() + np.ascontiguousarray(p).tobytes())
            if subset is None:
                f.write(np.uint32(0xffffffff).tobytes())
            else:
                f.write(np.uint32(subset.size).tobytes() + subset.tobytes())
            f.write(np.uint32(max_depth).tobytes() + np.int32(crit_kind).tobytes() + np.float32(crit_value).tobytes())
            f.write(np.int32(trav).tobytes() + np.uint32(level).tobytes())
        subprocess.check_call([exe, "octree", inp, outp])
        raw = open(outp, "rb").read()
        got_depth, got_n = np.frombuffer(raw[:8], dtype=np.uint32)
        got = np.frombuffer(raw[8:], dtype=np.float64).reshape(-1, 12)
        assert (got_depth, got_n) == (depth.value, n), (c, got_depth, got_n, depth.value, n)
        assert np.array_equal(got.view(np.uint64), rows[:n].view(np.uint64)), c
        n_nodes += n
    assert n_nodes > 10000


# ---- row a15: opencl/{util,cylinder,icp}.cl compiled as C++ (built-ins are stand-ins, see the shim) ----
REF_CL = os.path.join(os.path.dirname(REF), "libtm_ref_cl.so")


def test_reference_opencl_kernels(built):
    """The oracle's restatement of icp_projection / icp_correlation against the reference's own kernel
    source run one work-item at a time (with a padded range: the kernels' own guard stops the extras)."""
    if not os.path.exists(REF_CL):
        pytest.skip("oracle/_ref/libtm_ref_cl.so not built (no /root/reference in this environment)")
    import test_uvicp as tu
    L = C.CDLL(REF_CL)
    L.ref_cl_icp_projection.restype = C.c_uint32
    total_hits = 0
    for seed, maxd in ((1, 0.02), (2, 0.05), (3, 1e9), (4, 0.0)):
        pnts, image, sz, mg, ma, mu, mp, mn = tu._setup(seed, n=4000)
        if seed == 3:  # a transform that throws many points out of the image, plus NaN / huge coordinates
            pnts[:50, :3] *= 1e6
            pnts[50:60, 0] = np.nan
            pnts[60:70, 1] = np.inf
        n = pnts.shape[0]
        op_o, mi_o, si_o, c_o = po.cl_icp_projection(0, pnts, image, sz, mg, ma, mu, mp, mn, maxd)
        op = np.full((n + 7, 4), 7.0, np.float32)
        mi = np.full(n + 7, 12345, np.int32)
        si = np.full(n + 7, 12345, np.int32)
        c = L.ref_cl_icp_projection(_p(pnts), C.c_int(n), C.c_int(7), _p(image), _p(sz), _p(mg), _p(ma), _p(mu), _p(mp),
                                    _p(mn), C.c_float(maxd), _p(op), _p(mi), _p(si))
        assert c == c_o
        assert np.array_equal(mi[:n], mi_o) and np.array_equal(si[:n], si_o)
        assert np.array_equal(op[:n].view(np.uint32), op_o.view(np.uint32))
        assert np.all(mi[n:] == 12345) and np.all(op[n:] == 7.0)  # `index >= n` guard
        total_hits += c
        # icp_correlation over the correspondences just found
        sel = np.flatnonzero(mi_o >= 0)
        if sel.size < 2:
            continue
        is_, im_ = si_o[sel].astype(np.int32), mi_o[sel].astype(np.int32)
        cs = np.append(op_o[sel, :3].mean(0), 0).astype(np.float32)
        cm = np.append(image[im_, :3].mean(0), 0).astype(np.float32)
        rec_o, _ = po.cl_icp_correlation(op_o, image, is_, im_, cs, cm)
        rec = np.full((sel.size + 3, 16), 7.0, np.float32)
        L.ref_cl_icp_correlation(_p(op_o), _p(image), _p(is_), _p(im_), C.c_int(sel.size), C.c_int(3), _p(cs), _p(cm), _p(rec))
        assert np.array_equal(rec[:sel.size].view(np.uint32), rec_o.view(np.uint32))
        assert np.all(rec[sel.size:] == 7.0)
    assert total_hits > 3000


def test_reference_traits_init_from_samples(ref, tmp_path):
    """init_from_samples of the cylinder / plane / plane2 traits (the frames `project` works in): drop-in
    headers vs the reference's own code — frame, radius, origin bit for bit, and the same nullptr outcomes."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "triplet_match_b200")
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_dropin.cpp"), "-o", exe, "-L" + libdir,
                           "-ltriplet_match_b200", "-Wl,-rpath," + libdir])
    rng = np.random.default_rng(21)
    n = 600
    for kind in (0, 1, 2):
        pos = rng.standard_normal((n, 3, 3))
        nrm = rng.standard_normal((n, 3, 3))
        nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
        if kind == 0:  # samples on random cylinders (normals radial, + noise), plus parallel normals (denominator ~ 0)
            for c in range(n):
                q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
                rad, o = 0.2 + rng.random(), rng.standard_normal(3)
                for k in range(2):
                    th, z = rng.random() * 2 * np.pi, rng.standard_normal()
                    loc = np.array([np.cos(th), np.sin(th), 0.0])
                    pos[c, k] = o + q @ (rad * loc + np.array([0, 0, z]))
                    nrm[c, k] = q @ loc + 0.02 * rng.standard_normal(3)
            nrm[:20, 1] = nrm[:20, 0]
        elif kind == 2:  # three points of a plane whose normals agree with it (or not: nullptr)
            for c in range(n):
                q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
                o = rng.standard_normal(3)
                for k in range(3):
                    pos[c, k] = o + q @ np.array([rng.standard_normal(), rng.standard_normal(), 0.0])
                    nrm[c, k] = (q[:, 2] if c % 3 else rng.standard_normal(3)) + 0.05 * rng.standard_normal(3)
            nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
        elif kind == 1:  # axis-aligned normals hit both unitOrthogonal branches
            nrm[:6, 0] = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [1e-7, 0, 1], [0, 1e-7, 1]], dtype=np.float64)
        samples = np.ascontiguousarray(np.concatenate([pos, nrm], axis=2).reshape(n, 18), dtype=np.float32)
        want = np.zeros((n, 21), np.float32)
        ref.ref_traits_init(C.c_int(kind), _p(samples), C.c_uint32(n), C.c_float(0.1), _p(want))
        inp, outp = str(tmp_path / f"i{kind}.in"), str(tmp_path / f"i{kind}.out")
        with open(inp, "wb") as f:
            f.write(np.int32(kind).tobytes() + np.float32(0.1).tobytes() + np.uint32(n).tobytes() + samples.tobytes())
        subprocess.check_call([exe, "traits_init", inp, outp])
        got = np.fromfile(outp, dtype=np.float32).reshape(n, 21)
        assert np.array_equal(got[:, 0], want[:, 0]), kind
        if kind == 2:
            assert 0 < want[:, 0].sum() < n  # both outcomes
        bad = np.nonzero((got.view(np.uint32) != want.view(np.uint32)).any(axis=1))[0]
        assert bad.size == 0, (kind, bad[:5], got[bad[:2]], want[bad[:2]])


def test_reference_free_functions_dropin_headers(ref, tmp_path):
    """include/triplet_match/{feature,discretize} (the drop-in's host copies of rows a1-a4) against the
    reference's own functions on random and degenerate pairs, bit for bit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "triplet_match_b200")
    exe = str(tmp_path / "test_dropin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_dropin.cpp"), "-o", exe, "-L" + libdir,
                           "-ltriplet_match_b200", "-Wl,-rpath," + libdir])
    rng = np.random.default_rng(31)
    n = 4000
    pairs = rng.standard_normal((n, 12)).astype(np.float32)
    pairs[:, 3:6] /= np.linalg.norm(pairs[:, 3:6], axis=1, keepdims=True)
    pairs[:, 9:12] /= np.linalg.norm(pairs[:, 9:12], axis=1, keepdims=True)
    pairs[:50, 6:9] = pairs[:50, 0:3]            # coincident points: d = 0
    pairs[50:100, 3:6] = 0                       # zero tangent
    pairs[100:150, 6:9] = pairs[100:150, 0:3] + 0.3 * pairs[100:150, 3:6]  # d parallel to t0
    pairs[150:200] *= 1e-3
    mn = np.array([0.2, 0.0, 0.0, 0.2], np.float32)
    mx = np.array([3.7, np.pi, np.pi, 3.7], np.float32)
    dist_steps, angle_step, min_rel, max_rel = np.float32(20.0), np.float32(0.17453292), np.float32(0.2), np.float32(0.9)
    inp, outp = str(tmp_path / "free.in"), str(tmp_path / "free.out")
    with open(inp, "wb") as f:
        f.write(mn.tobytes() + mx.tobytes() + dist_steps.tobytes() + angle_step.tobytes() + min_rel.tobytes() + max_rel.tobytes())
        f.write(np.uint32(n).tobytes() + np.ascontiguousarray(pairs).tobytes())
    subprocess.check_call([exe, "free", inp, outp])
    got = np.fromfile(outp, dtype=np.uint32).reshape(n, 20)
    ref.ref_valid.restype = C.c_int
    vmn, vmx = np.zeros(4, np.float32), np.zeros(4, np.float32)
    ref.ref_valid_bounds(_p(mn), _p(mx), C.c_float(min_rel), C.c_float(max_rel), _p(vmn), _p(vmx))
    assert np.array_equal(got[0, 12:16], vmn.view(np.uint32)) and np.array_equal(got[0, 16:20], vmx.view(np.uint32))
    n_valid = 0
    for c in range(n):
        ft, key = np.zeros(4, np.float32), np.zeros(4, np.uint32)
        ref.ref_feature(_p(pairs[c]), _p(ft))
        same = (got[c, :4] == ft.view(np.uint32)) | (np.isnan(ft) & np.isnan(got[c, :4].view(np.float32)))
        assert same.all(), (c, got[c, :4].view(np.float32), ft)
        v = ref.ref_valid(_p(ft), _p(vmn), _p(vmx))
        assert got[c, 4] == v, c
        n_valid += v
        if np.isnan(ft).any():
            continue  # float -> uint32 of NaN is undefined in both
        ref.ref_discretize_feature(_p(ft), _p(mn), _p(mx), C.c_float(dist_steps), C.c_float(angle_step), _p(key))
        assert np.array_equal(got[c, 5:9], key), (c, got[c, 5:9], key)
        assert got[c, 9] == ref.ref_murmur4(_p(key))
        assert (int(got[c, 11]) << 32 | int(got[c, 10])) == ref.ref_std_hash4(_p(key))
    assert 0 < n_valid < n
