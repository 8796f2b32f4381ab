"""Adversarial parity cases (-m gpu): the CUDA path against the oracle on inputs chosen to sit on
every decision boundary of the path — voxel cell edges and the (-1, 0) truncation band of
voxel_query (model.hpp:182-189), points exactly at the distance threshold, non-finite points and
transforms, duplicate points, zero tangents, and a table with thousands of keys (open-addressing
collisions)."""
import os

import numpy as np
import pytest

import common
from oracle import pyoracle as po
from triplet_match_b200 import synth

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def ctx(built):
    from triplet_match_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _boundary_scene(m, om, seed):
    """Scene points that map (under identity-like transforms) onto voxel-coordinate integers, into the
    (-1, 0) band, just outside the grid, and onto model points +- the distance threshold."""
    rng = np.random.default_rng(seed)
    sc, tr, ex = om.scale.astype(np.float64), om.trans.astype(np.float64), om.extents
    pts = []
    # voxel coordinates exactly on / next to integers, incl. the truncation band and both grid faces
    for _ in range(3000):
        v = np.array([rng.integers(-2, ex[0] + 2), rng.integers(-2, ex[1] + 2), rng.integers(-2, ex[2] + 2)], np.float64)
        v += rng.choice([0.0, 1e-7, -1e-7, 0.5, -0.5, 0.999999, -0.999999], size=3)
        pts.append((v - tr) / sc)
    # model points displaced by almost exactly the threshold
    thres = om.resolution
    for i in rng.integers(0, m.n, 2000):
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        pts.append(m.pos[i].astype(np.float64) + d * thres * rng.choice([0.999999, 1.0, 1.000001, 0.5, 0.0]))
    pts = np.array(pts, F)
    # non-finite and huge coordinates, duplicates
    bad = np.array([[np.nan, 0, 0], [0, np.inf, 0], [0, 0, -np.inf], [1e30, 1e30, 1e30], [-1e30, 0, 0], [0, 0, 0]], F)
    pts = np.concatenate([pts, bad, pts[:50]])
    n = pts.shape[0]
    nrm = rng.standard_normal((n, 3)).astype(F)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tgt = np.zeros((n, 3), F)
    tm = (rng.random(n) < 0.3).astype(np.uint8)
    t = rng.standard_normal((n, 3)).astype(F)
    tgt[tm == 1] = (t / np.linalg.norm(t, axis=1, keepdims=True))[tm == 1]
    return synth.Cloud(pts, nrm, tgt, tm)


@pytest.mark.parametrize("name", ["plane_small", "cylinder_small", "freeform_small"])
def test_scoring_on_decision_boundaries(ctx, name):
    from triplet_match_b200 import capi
    m, s0, om, osc0, rec = common.config(name)
    s = _boundary_scene(m, om, 1)
    osc = po.OScene(s)
    gm = common.upload_model(ctx, m, om)
    gs = capi.Scene(ctx, s.pos, s.nrm, s.tgt, s.tangent_mask)
    rng = np.random.default_rng(2)
    Ts = [np.eye(4)]
    for _ in range(40):  # tiny perturbations of identity keep the points on the boundaries "almost"
        a = rng.standard_normal(3) * 1e-6
        T = np.eye(4)
        T[:3, :3] += np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        T[:3, 3] = rng.standard_normal(3) * 1e-7
        Ts.append(T)
    for k in range(3):  # axis permutations / reflections: exact arithmetic, other cells
        P = np.eye(4)
        P[:3, :3] = np.roll(np.eye(3), k, axis=0) * (-1 if k == 2 else 1)
        Ts.append(P)
    nanT = np.eye(4); nanT[0, 0] = np.nan
    infT = np.eye(4); infT[1, 3] = np.inf
    hugeT = np.eye(4) * 1e20; hugeT[3, 3] = 1
    zeroT = np.zeros((4, 4)); zeroT[3, 3] = 1
    Ts += [nanT, infT, hugeT, zeroT]
    T16 = np.stack([T.T.reshape(-1) for T in Ts]).astype(F)  # column-major
    for eo in (False, True):
        cg, sg, dg = gs.score(gm, T16, early_out=eo)
        co, so, do = osc.score_batch(om, T16, early_out=eo, nthreads=4)
        assert np.array_equal(cg, co), (name, eo, np.flatnonzero(cg != co))
        assert np.array_equal(dg, do)
        # scores are 2^-36 fixed point: exact for |ref . ref_n| terms below 2^27, i.e. for any rigid
        # transform (terms <= 1); the 1e20-scaled matrix is outside that domain (counts still agree)
        rigid = np.array([abs(np.linalg.det(T[:3, :3])) < 10 if np.isfinite(T).all() else True for T in Ts])
        assert np.allclose(sg[rigid], so[rigid], rtol=1e-9, atol=1e-9)
    assert co[0] > 0
    for h in (0, 5, len(Ts) - 8):
        a = osc.project(om, np.arange(s.n, dtype=np.int32), T16[h])
        scn, mdl, score = gs.correspondences(gm, T16[h], 1.0)
        assert np.array_equal(scn, a["scene_corrs"]) and np.array_equal(mdl, a["model_corrs"])
    # ICP from these transforms (incl. the degenerate ones) follows the oracle's control flow
    To, cnt, scr, it = gs.icp(gm, T16[:6], 3, 1.0)
    for h in range(6):
        oT, on, osx, oit = osc.icp(om, T16[h], 3, 1.0)
        assert cnt[h] == on and it[h] == oit and np.abs(To[h] - oT).max() < 1e-4
    gs.close(); gm.close()


def test_features_on_degenerate_pairs(ctx):
    """Zero / non-finite tangents, coincident points, tangent parallel to the pair direction, distance
    exactly at the window ends."""
    from triplet_match_b200 import capi
    m, s0, om, osc0, rec = common.config("plane_small")
    rng = np.random.default_rng(3)
    n = 600
    pos = (rng.random((n, 3)) * om.diameter).astype(F)
    tgt = rng.standard_normal((n, 3)).astype(F)
    tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
    pos[1] = pos[0]                                   # coincident pair (0, 1)
    tgt[2] = 0                                        # zero tangent
    tgt[3] = [np.nan, 0, 0]
    pos[5] = pos[4] + tgt[4] * F(0.5 * om.diameter)   # tangent exactly along the pair direction
    lower, upper = F(0.2) * F(om.diameter), F(1.0) * F(om.diameter)
    pos[7] = pos[6] + np.array([lower, 0, 0], F)      # window ends
    pos[9] = pos[8] + np.array([upper, 0, 0], F)
    pos[10] = [np.inf, 0, 0]
    nrm = np.tile(np.array([0, 0, 1], F), (n, 1))
    tm = np.ones(n, np.uint8)
    tm[11] = 0                                        # not a tangent point: pair filtered (scene.hpp:290)
    s = synth.Cloud(pos, nrm, tgt, tm)
    osc = po.OScene(s)
    gm = common.upload_model(ctx, m, om)
    gs = capi.Scene(ctx, s.pos, s.nrm, s.tgt, s.tangent_mask)
    pi = np.concatenate([[0, 2, 3, 4, 6, 8, 10, 11, 12, 12], rng.integers(0, n, 3000)]).astype(np.uint32)
    pj = np.concatenate([[1, 12, 12, 5, 7, 9, 12, 12, 11, 12], rng.integers(0, n, 3000)]).astype(np.uint32)
    f, k, v = gs.features(gm, pi, pj, 0.2, 1.0)
    fo, ko, vo = osc.pair_features(om, pi, pj)
    assert np.array_equal(v, vo) and np.array_equal(k[v.astype(bool)], ko[vo.astype(bool)])
    ok = v.astype(bool)
    assert np.array_equal(f[ok].view(np.uint32), fo[ok].view(np.uint32)) and ok.sum() > 100
    off, hits = gm.probe(k, v, 200)
    Tg, vg = gs.hypotheses(gm, pi, pj, off, hits)
    T, hp, mi, mj, va = osc.hypotheses(om, pi, pj)
    assert np.array_equal(vg, va) and np.array_equal(Tg[va.astype(bool)].view(np.uint32), T[va.astype(bool)].view(np.uint32))
    gs.close(); gm.close()


def test_table_with_thousands_of_keys(ctx):
    """distance_step_count = 400 and 2-degree angle bins: ~10^4 distinct keys -> long probe sequences
    in the open-addressing table; hits and their order must still equal equal_range's."""
    from triplet_match_b200 import capi
    m = synth.freeform_model(seed=3, n_points=1500, radius=0.12, n_bumps=6, n_curves=5)
    s = synth.make_scene(seed=7, model=m, n_points=24000, n_copies=4, extent=1.2, flat_copies=False)
    s = s.take(synth.morton_order(s.pos))
    dp = dict(distance_step_count=400.0, angle_step=float(np.deg2rad(2.0)))
    om = po.OModel(m, **dp, **common.SP)
    assert om.n_keys > 3000
    osc = po.OScene(s)
    rec = synth.record_pairs(11, s, om.diameter, 10, 48)
    keys, offsets, pairs = om.table(200)
    gm = capi.Model(ctx, m.pos, m.nrm, m.tgt, voxel=om.voxel, extents=om.extents, to_voxel16=om.to_voxel16,
                    resolution=om.resolution, diameter=om.diameter, keys=keys, offsets=offsets, pairs=pairs,
                    feat_min=om.feat_min, feat_max=om.feat_max, **dp)
    gs = common.upload_scene(ctx, s)
    f, k, v = gs.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
    fo, ko, vo = osc.pair_features(om, rec.pair_i, rec.pair_j)
    assert np.array_equal(v, vo) and np.array_equal(k, ko)
    off, hits = gm.probe(k, v, 200)
    T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    assert hits.shape[0] == T.shape[0] > 50 and np.array_equal(hits, np.stack([mi, mj], 1))
    # keys that are absent from the table (probe runs into an empty slot) give no hits
    absent = np.array([[399, 89, 89, 399], [0, 0, 0, 1], [123456, 1, 1, 123456]], np.uint32)
    o2, h2 = gm.probe(absent, None, 200)
    assert o2[-1] == 0 and h2.shape[0] == 0
    # the product's own host build produces the same table
    hm = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **dp, **common.SP)
    assert hm.n_keys == om.n_keys
    order = lambda kk: np.lexsort(kk.T[::-1])
    a, b = order(hm.keys), order(keys)
    assert np.array_equal(hm.keys[a], keys[b])
    for x, y in zip(a[:200], b[:200]):
        assert np.array_equal(hm.pairs[hm.offsets[x]:hm.offsets[x + 1]], pairs[offsets[y]:offsets[y + 1]])
    hm.close(); gs.close(); gm.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_irregular_clouds(ctx, seed):
    """Unstructured random clouds (the generator of tests/test_oracle_vs_ref_fuzz.py, on which the
    oracle equals the reference's own code): product-built model, every stage vs the oracle."""
    from triplet_match_b200 import capi
    from test_oracle_vs_ref_fuzz import _cloud
    rng = np.random.default_rng(seed)
    m = _cloud(rng, 260 + 40 * seed, 0.25, 60 + 5 * seed, dup=3)
    s = _cloud(rng, 2500, 0.6, 300)
    s.pos[:m.n] = m.pos + F(0.17)
    om = po.OModel(m, **common.DP, **common.SP)
    osc = po.OScene(s)
    hm = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **common.DP, **common.SP)
    assert np.array_equal(hm.voxel, om.voxel) and np.array_equal(hm.feat_min.view(np.uint32), om.feat_min.view(np.uint32))
    gm = hm.upload(ctx)
    gs = capi.Scene(ctx, s.pos, s.nrm, s.tgt, s.tangent_mask)
    tidx = np.flatnonzero(s.tangent_mask)
    pi = rng.choice(tidx, 600).astype(np.uint32)
    pj = rng.choice(tidx, 600).astype(np.uint32)
    f, k, v = gs.features(gm, pi, pj, 0.2, 1.0)
    fo, ko, vo = osc.pair_features(om, pi, pj)
    assert np.array_equal(v, vo) and np.array_equal(k[v.astype(bool)], ko[vo.astype(bool)])
    off, hits = gm.probe(k, v, 200)
    Tg, vg = gs.hypotheses(gm, pi, pj, off, hits)
    T, hp, mi, mj, va = osc.hypotheses(om, pi, pj)
    assert np.array_equal(hits, np.stack([mi, mj], 1)) and np.array_equal(vg, va)
    ok = va.astype(bool)
    assert np.array_equal(Tg[ok].view(np.uint32), T[ok].view(np.uint32))
    Ts = [T[h] for h in np.flatnonzero(ok)[:40]]
    for _ in range(12):  # near the planted pose: many inliers
        M = np.eye(4)
        a = rng.standard_normal(3) * 0.02
        M[:3, :3] += np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        M[:3, 3] = -0.17 + rng.standard_normal(3) * 0.003
        Ts.append(M.T.reshape(-1).astype(F))
    T16 = np.stack(Ts).astype(F)
    for eo in (False, True):
        cg, sg, dg = gs.score(gm, T16, early_out=eo)
        co, so, do = osc.score_batch(om, T16, early_out=eo, nthreads=4)
        assert np.array_equal(cg, co) and np.array_equal(dg, do) and np.allclose(sg, so, rtol=1e-9, atol=1e-9)
    assert co.max() > 50
    # the resident query's early drop over the evenly sampling walk (level-by-level evaluation, k_early2.cu) on the
    # same clouds: recorded list from the random pairs, the oracle walks the permuted balls of the query's own hypotheses
    order = np.argsort(pi, kind="stable")
    outer, pair_outer = np.unique(pi[order], return_inverse=True)
    subs = [osc.ball_subset(int(o), om.diameter) for o in outer]
    offs = np.zeros(len(subs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([x.size for x in subs])
    walk = np.concatenate([x[capi.walk_order(x.size)] for x in subs]).astype(np.int32)
    for accept in (0.1 * seed, 0.6):
        q = capi.Query(gs, gm, early_out=2, accept_prob=accept)
        q.set_pairs(outer.astype(np.uint32), pair_outer.astype(np.uint32), pj[order])
        q.run()
        d = q.download()
        cw, sw, dw = osc.score_batch(om, d["T"], pair_outer.astype(np.uint32)[d["hyp_pair"]], offs, walk, early_out=True,
                                     accept_prob=accept, nthreads=4)
        assert np.array_equal(d["counts"], cw) and np.array_equal(d["dropped"], dw), (seed, accept)
        assert np.allclose(d["scores"], sw, rtol=1e-9, atol=1e-9)
        q.close()
    gs.close(); gm.close(); hm.close()


@pytest.mark.gpu
def test_chained_scan_matches_cumsum(ctx):
    """The multi-CTA chained scan behind the hypothesis offsets of long recorded lists: every length around the tile
    (4 096) and switch-over (8 192) boundaries, ragged tails, zeros, and sums beyond 2^32."""
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 4095, 4096, 4097, 8191, 8192, 8193, 12288, 12289, 100_003, 262_144, 1_000_001):
        v = rng.integers(0, 201, size=n, dtype=np.uint32)
        if n > 10:
            v[rng.integers(0, n, n // 7)] = 0
        exp = np.concatenate([[0], np.cumsum(v.astype(np.uint64))]).astype(np.uint64)
        assert np.array_equal(ctx.scan_u64(v), exp), n
    big = np.full(70_000, 0xFFFFFFF0, dtype=np.uint32)  # total far beyond 32 bits
    assert np.array_equal(ctx.scan_u64(big), np.concatenate([[0], np.cumsum(big.astype(np.uint64))]).astype(np.uint64))


@pytest.mark.gpu
def test_sharding_edge_cases(ctx):
    """Shards of equal counts and of equal tests at the edges: more ranks than outer samples or hypotheses, a hyp_limit
    inside the list, an empty list, capacity clipping — the shards always tile the (clipped) global list in rank order."""
    from triplet_match_b200 import capi
    m, s, om, osc, rec = common.config("plane_small")
    gm = common.upload_model(ctx, m, om)
    gs = common.upload_scene(ctx, s)
    q0 = capi.Query(gs, gm)
    q0.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q0.run()
    full = q0.download_counts()[0]
    H = full.size
    assert H > 50
    for by_tests in (False, True):
        for world, limit in ((2, 0), (5, 0), (16, 0), (64, 0), (3, H // 3 + 1), (4, 7), (2, 1)):
            parts, keys = [], []
            for rank in range(world):
                q = capi.Query(gs, gm, hyp_limit=limit)
                q.set_shard(rank, world)
                q.set_balance(by_tests)
                q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
                q.run()
                c, r = q.download_counts()
                parts.append(c)
                keys.append(int(r.best_key))
                assert int(r.n_hypotheses) == (min(H, limit) if limit else H)
                q.close()
            n = min(H, limit) if limit else H
            assert np.array_equal(np.concatenate(parts), full[:n]), (by_tests, world, limit)
            exp = max((capi.pack_key(int(c), i) for i, c in enumerate(full[:n]) if c), default=0)
            assert max(keys) == exp
    # the level-by-level early drop on shards: outer samples outside a rank's range have empty rows and no work items
    qe = capi.Query(gs, gm, early_out=2, max_hypotheses=H + 64)
    qe.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    qe.run()
    de = qe.download()
    qe.close()
    for world in (2, 5):
        cs, ds = [], []
        for r in range(world):
            q = capi.Query(gs, gm, early_out=2, max_hypotheses=H + 64)
            q.set_shard(r, world)
            q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
            q.run()
            d = q.download()
            cs.append(d["counts"]); ds.append(d["dropped"])
            q.close()
        assert np.array_equal(np.concatenate(cs), de["counts"]) and np.array_equal(np.concatenate(ds), de["dropped"])
    # empty list, any sharding
    e = np.zeros(0, np.uint32)
    q = capi.Query(gs, gm)
    q.set_shard(1, 3)
    q.set_balance(True)
    q.set_pairs(e, e, e)
    q.run()
    assert q.result().n_hypotheses == 0 and q.download_counts()[0].size == 0
    q.close()
    # capacity smaller than the shard: reported, never silently truncated
    q = capi.Query(gs, gm, max_hypotheses=H // 2)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    with pytest.raises(capi.TmError):
        q.result()
    q.close()
    q0.close(); gm.close(); gs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("keep", [400, 2500])
def test_early_drop_levels_on_tiny_and_sparse_subsets(ctx, keep):
    """tm_query_run(early_out = 2), level scheme (k_early2.cu), where its regular case does not hold: subsets of a
    handful of points (checkpoint ranges that are empty — several checkpoints share an element) and hypotheses whose
    ranges reach no grid cell are walked one by one; the outcome still equals the oracle's walk, hypothesis by
    hypothesis.  A scene with non-finite points rides along."""
    from triplet_match_b200 import capi
    m, s0, om, osc0, rec0 = common.config("cylinder_small")
    rng = np.random.default_rng(keep)
    pick = np.sort(rng.choice(s0.n, keep, replace=False))
    s = s0.take(pick)
    s.pos[::97] = np.nan  # never reaching, never inliers
    s.pos[5::131, 1] = np.inf
    osc = po.OScene(s)
    rec = synth.record_pairs(3, s, om.diameter, 12, 16)
    gm = common.upload_model(ctx, m, om)
    gs = common.upload_scene(ctx, s)
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    subs = [osc.ball_subset(int(o), om.diameter) for o in rec.outer]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in subs])
    walk = np.concatenate([x[capi.walk_order(x.size)] for x in subs]).astype(np.int32)
    hyp_sub = rec.pair_outer[hp]
    walked = []
    for accept in (0.02, 0.5):
        co, so, do = osc.score_batch(om, T, hyp_sub, off, walk, early_out=True, accept_prob=accept, nthreads=4)
        q = capi.Query(gs, gm, early_out=2, accept_prob=accept)
        q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        q.run()
        d = q.download()
        assert d["counts"].size == co.size
        assert np.array_equal(d["counts"], co), (keep, accept, np.flatnonzero(d["counts"] != co)[:8])
        assert np.array_equal(d["dropped"], do)
        assert np.allclose(d["scores"], so, rtol=1e-9, atol=1e-9)
        walked.append(q.early_walked())
        q.close()
    sizes = np.diff(off.astype(np.int64))
    if keep == 400:
        assert sizes.min() < 20
        if os.environ.get("TM_EARLY_LEVELS", "1") != "0":  # (the knob routes everything through the walker)
            assert max(walked) > 0  # the irregular path is exercised
    gs.close(); gm.close()


@pytest.mark.gpu
def test_topk_selection_paths(ctx):
    """The top-k of the ICP stage (count descending, index ascending): threshold + compaction + ranking where few keys
    compete, the slice kernels where the short list overflows (many equal counts); both equal a stable host sort.
    Zero counts and excluded entries never appear; fewer than k candidates leave 0xFFFFFFFF."""
    rng = np.random.default_rng(3)

    def expect(counts, k, excluded=None):
        ok = counts > 0
        if excluded is not None:
            ok &= excluded == 0
        idx = np.flatnonzero(ok)
        order = idx[np.argsort(-counts[idx].astype(np.int64), kind="stable")][:k]
        out = np.full(k, 0xFFFFFFFF, np.uint32)
        out[: order.size] = order
        return out

    cases = []
    cases.append((rng.integers(0, 50000, 1 << 20).astype(np.uint32), 64, None))          # the usual case: short list
    cases.append((np.full(300000, 7, np.uint32), 64, None))                                # all equal: the list overflows
    c = rng.integers(0, 3, 200000).astype(np.uint32)
    cases.append((c, 100, (rng.random(c.size) < 0.5).astype(np.uint8)))                    # few distinct values, exclusions
    cases.append((np.zeros(5000, np.uint32), 16, None))                                    # nothing qualifies
    c = np.zeros(70000, np.uint32); c[[5, 69999, 4096, 4095]] = [9, 9, 3, 12]
    cases.append((c, 8, None))                                                             # fewer than k candidates
    cases.append((rng.integers(1, 1000, 37).astype(np.uint32), 64, None))                  # n < k
    c = rng.integers(0, 1 << 31, 4096 * 70 + 13).astype(np.uint32)
    cases.append((c, 64, None))                                                            # more slices than k, huge counts
    c = np.arange(1, 4096 * 80 + 1, dtype=np.uint32)[::-1].copy()
    cases.append((c, 4096, None))                                                          # k = 4096, descending
    for counts, k, ex in cases:
        got = ctx.select_topk(counts, k, ex)
        assert np.array_equal(got, expect(counts, k, ex)), (counts.size, k)
