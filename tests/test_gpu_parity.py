"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle
on the same seeded inputs, against the committed golden fixtures, and — at the full C2 size —
through size-independent properties.  Bit-exact for keys, hits, transforms, subsets and
inlier counts; scores to 1e-9 (fixed-point 2^-36 accumulation vs double sum); ICP poses to
1e-4 (north_star tolerance)."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["plane_small", "cylinder_small", "freeform_small"]
SHUFFLED = ["plane_small_shuffled", "cylinder_small_shuffled"]


@pytest.fixture(scope="module")
def ctx(built):
    from triplet_match_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def _subsets(osc, om, rec):
    subs = [osc.ball_subset(int(o), om.diameter) for o in rec.outer]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in subs])
    return off, np.concatenate(subs).astype(np.int32)


@pytest.fixture(scope="module", params=CONFIGS + SHUFFLED)
def setup(request, ctx):
    m, s, om, osc, rec = common.config(request.param)
    gm = common.upload_model(ctx, m, om)
    gs = common.upload_scene(ctx, s)
    yield request.param, m, s, om, osc, rec, gm, gs
    gm.close()
    gs.close()


def test_features_keys_valid(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    f, k, v = gs.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
    fo, ko, vo = osc.pair_features(om, rec.pair_i, rec.pair_j)
    assert np.array_equal(v, vo) and v.sum() > 10
    assert np.array_equal(k, ko)
    ok = v.astype(bool)
    assert np.array_equal(_bits(f[ok]), _bits(fo[ok]))


def test_probe_hits_order_and_limit(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    f, k, v = gs.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
    for limit in (200, 3):
        off, hits = gm.probe(k, v, limit)
        T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j, limit=limit)
        assert hits.shape[0] == T.shape[0]
        assert np.array_equal(hits, np.stack([mi, mj], 1))
        assert np.array_equal(np.repeat(np.arange(v.size), np.diff(off.astype(np.int64))), hp)


def test_hypotheses_bit_exact(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    f, k, v = gs.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
    off, hits = gm.probe(k, v, 200)
    for force_up in (False, True):
        Tg, vg = gs.hypotheses(gm, rec.pair_i, rec.pair_j, off, hits, force_up=force_up)
        T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j, force_up=force_up)
        assert np.array_equal(vg, va)
        ok = va.astype(bool)
        assert np.array_equal(_bits(Tg[ok]), _bits(T[ok]))
        if force_up:
            rows = [0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14]  # rows 0..2 of the column-major 4x4
            assert np.isnan(Tg[~ok][:, rows]).all()  # rejected hypotheses can never score


def test_ball_subsets(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    off, idx = gs.ball_subsets(rec.outer, om.diameter)
    eo, ei = _subsets(osc, om, rec)
    assert np.array_equal(off, eo) and np.array_equal(idx, ei)
    o2, i2 = gs.ball_subsets(rec.outer[:1], 0.0)  # empty ball (strict '<')
    assert o2[-1] == 0 and i2.size == 0


def test_scoring_full_and_early_drop(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    off, idx = _subsets(osc, om, rec)
    hyp_sub = rec.pair_outer[hp]
    for eo in (False, True):
        cg, sg, dg = gs.score(gm, T, hyp_sub, off, idx, early_out=eo)
        co, so, do = osc.score_batch(om, T, hyp_sub, off, idx, early_out=eo, nthreads=4)
        assert np.array_equal(cg, co), (name, eo)
        assert np.array_equal(dg, do)
        assert np.allclose(sg, so, rtol=1e-9, atol=1e-9)
    if name.endswith("_shuffled"):
        assert 0 < do.sum() < do.size  # both outcomes of the early-drop rule are exercised


def test_early_drop_over_the_evenly_sampling_walk(setup):
    """early_out = 2 == the reference's drop test on the subsets rewritten in walk order (the oracle gets the
    permuted rows explicitly), and, on the resident query, dropped hypotheses are no ICP candidates."""
    from triplet_match_b200 import capi
    name, m, s, om, osc, rec, gm, gs = setup
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    off, idx = _subsets(osc, om, rec)
    hyp_sub = rec.pair_outer[hp]
    walk = idx.copy()
    for g in range(off.size - 1):
        b, e = int(off[g]), int(off[g + 1])
        walk[b:e] = idx[b:e][capi.walk_order(e - b)]
    cg, sg, dg = gs.score(gm, T, hyp_sub, off, idx, early_out=2)
    co, so, do = osc.score_batch(om, T, hyp_sub, off, walk, early_out=True, nthreads=4)
    assert np.array_equal(cg, co) and np.array_equal(dg, do), name
    assert np.allclose(sg, so, rtol=1e-9, atol=1e-9)
    # whole scene as the subset (hyp_sub = None): identity rows in walk order
    sel = np.linspace(0, T.shape[0] - 1, 24).astype(np.int64)
    c2, s2, d2 = gs.score(gm, T[sel], early_out=2)
    o_off = np.array([0, s.n], dtype=np.uint64)
    c3, s3, d3 = osc.score_batch(om, T[sel], np.zeros(sel.size, np.uint32), o_off, capi.walk_order(s.n).astype(np.int32),
                                 early_out=True, nthreads=4)
    assert np.array_equal(c2, c3) and np.array_equal(d2, d3) and np.allclose(s2, s3, rtol=1e-9, atol=1e-9)
    # resident query in the same mode: same outcomes, and the ICP candidates are the best non-dropped hypotheses
    q = capi.Query(gs, gm, early_out=2, icp_top_k=4, max_icp_iterations=1)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    d = q.download()
    assert np.array_equal(d["counts"], co) and np.array_equal(d["dropped"], do)
    assert np.allclose(d["scores"], so, rtol=1e-9, atol=1e-9)  # on request, through the walker
    ids, *_ = q.icp_results()
    keep = np.flatnonzero((do == 0) & (co > 0))
    want = keep[np.argsort(-co[keep].astype(np.int64), kind="stable")][:4]
    assert np.array_equal(ids[: want.size], want.astype(np.uint32)) and np.all(ids[want.size:] == 0xFFFFFFFF)
    if keep.size:  # the winner is never a dropped hypothesis; its score is the full sum
        r = d["result"]
        best = keep[np.argmax(co[keep])]
        assert r.best_hypothesis == best and r.best_inliers == co[best] and abs(r.best_score - so[best]) < 1e-9
    q.close()


@pytest.mark.parametrize("accept_prob", [0.05, 0.3, 0.9])
def test_early_drop_level_by_level(setup, accept_prob):
    """tm_query_run(early_out = 2) evaluates the drop test checkpoint range by checkpoint range with the tiled scorer
    (k_early2.cu).  Counts and drop flags equal the oracle's walk of the permuted subsets for acceptance bounds that
    make hypotheses fail at early, middle and late checkpoints, with and without a scene mask."""
    from triplet_match_b200 import capi
    name, m, s, om, osc, rec, gm, gs = setup
    off, idx = _subsets(osc, om, rec)
    walk = idx.copy()
    for g in range(off.size - 1):
        b, e = int(off[g]), int(off[g + 1])
        walk[b:e] = idx[b:e][capi.walk_order(e - b)]
    rng = np.random.default_rng(5)
    for mask in (None, (rng.random(s.n) < 0.4).astype(np.uint8)):
        if mask is not None:
            gs.set_mask(mask)
            osc.set_mask(mask)
        try:
            q = capi.Query(gs, gm, early_out=2, accept_prob=accept_prob)
            q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
            q.run()
            cg, r = q.download_counts()
            d = q.download()
            # the oracle walks the query's own hypothesis list (its transforms are compared bit for bit elsewhere)
            co, so, do = osc.score_batch(om, d["T"], rec.pair_outer[d["hyp_pair"]], off, walk, early_out=True,
                                         accept_prob=accept_prob, nthreads=4)
            assert np.array_equal(cg, co), (name, accept_prob, mask is not None)
            assert np.array_equal(d["dropped"], do)
            assert np.allclose(d["scores"], so, rtol=1e-9, atol=1e-9)
            assert q.early_walked() <= cg.size
            # a second run on the same query (buffers reused, accumulators reset) gives the same answer
            q.run()
            c2, r2 = q.download_counts()
            assert np.array_equal(c2, co) and r2.n_tests == r.n_tests and r2.best_key == r.best_key
            q.close()
        finally:
            if mask is not None:
                gs.set_mask(None)
                osc.set_mask(np.zeros(s.n, np.uint8))


def test_scoring_with_mask_and_all_scene(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    T = T[:: max(1, T.shape[0] // 96)]
    rng = np.random.default_rng(1)
    mask = (rng.random(s.n) < 0.3).astype(np.uint8)
    c0, _, _ = gs.score(gm, T)
    gs.set_mask(mask)
    osc.set_mask(mask)
    try:
        cg, sg, _ = gs.score(gm, T)
        co, so, _ = osc.score_batch(om, T, nthreads=4)
        assert np.array_equal(cg, co) and np.allclose(sg, so, rtol=1e-9, atol=1e-9)
        assert np.all(cg <= c0)  # masking can only remove inliers
    finally:
        gs.set_mask(None)
        osc.set_mask(np.zeros(s.n, dtype=np.uint8))
    c1, _, _ = gs.score(gm, T)
    assert np.array_equal(c0, c1)  # idempotent / mask cleared


def test_correspondences_and_icp(setup):
    name, m, s, om, osc, rec, gm, gs = setup
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    co, _, _ = osc.score_batch(om, T, nthreads=4)
    top = np.argsort(-co.astype(np.int64), kind="stable")[:6]
    for h in top[:2]:
        sc, mc, score = gs.correspondences(gm, T[h], 1.0)
        pr = osc.project(om, np.arange(s.n, dtype=np.int32), T[h])
        assert np.array_equal(sc, pr["scene_corrs"]) and np.array_equal(mc, pr["model_corrs"])
        assert abs(score - pr["score"]) < 1e-9
    # the batch call (candidate lists of a find_parallel round in one pass) == the single calls == the oracle, incl. a
    # transform without correspondences and a non-finite one
    far = np.eye(4, dtype=np.float32); far[:3, 3] = 1e3
    nanT = np.eye(4, dtype=np.float32); nanT[0, 0] = np.nan
    Tb = np.concatenate([T[top[:3]], far.T.reshape(1, 16), T[top[3:5]], nanT.T.reshape(1, 16)])
    off, bsc, bmc, bscore = gs.correspondences_batch(gm, Tb, 2.0)
    assert off[0] == 0 and off[4] == off[3] and off[-1] == off[-2] == bsc.size
    for t in range(Tb.shape[0]):
        sc, mc, score = gs.correspondences(gm, Tb[t], 2.0)
        assert np.array_equal(bsc[int(off[t]):int(off[t + 1])], sc) and np.array_equal(bmc[int(off[t]):int(off[t + 1])], mc)
        assert bscore[t] == score
    pr = osc.project(om, np.arange(s.n, dtype=np.int32), Tb[1], dist_thres=2.0)
    assert np.array_equal(bsc[int(off[1]):int(off[2])], pr["scene_corrs"])
    import ctypes as C
    from triplet_match_b200 import capi
    o2 = np.zeros(Tb.shape[0] + 1, np.uint64)
    small = np.zeros(4, np.uint32)
    rc = gs.lib.tm_correspondences_batch(gs.h, gm.h, Tb.ctypes.data_as(C.c_void_p), C.c_uint32(Tb.shape[0]), C.c_float(2.0),
                                         o2.ctypes.data_as(C.c_void_p), small.ctypes.data_as(C.c_void_p),
                                         small.ctypes.data_as(C.c_void_p), C.c_uint64(4), None)
    assert rc != 0 and np.array_equal(o2, off)  # too small a buffer is reported, nothing is written past it
    for iters in (0, 1, 5):
        To, cnt, scr, it = gs.icp(gm, T[top], iters, 1.0)
        for r, h in enumerate(top):
            oT, on, osx, oit = osc.icp(om, T[h], iters, 1.0)
            # pose tolerance from north_star (1e-4); counts are exact whenever the float
            # transforms agree bit-for-bit, else within the few points on the threshold
            assert np.allclose(To[r], oT, atol=1e-4), (name, iters, r)
            assert int(it[r]) == oit
            if np.array_equal(_bits(To[r]), _bits(oT)):
                assert int(cnt[r]) == on
            else:
                assert abs(int(cnt[r]) - on) <= max(3, on // 100)


def test_scene_sharded_icp_is_split_invariant(setup):
    """tm_icp_sharded: any split of the scene points gives tm_icp's result bit for bit (64-bit
    fixed-point sums); the NCCL leg of the same call is exercised by tools/mgpu_check.py."""
    name, m, s, om, osc, rec, gm, gs = setup
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T, nthreads=4)
    top = np.argsort(-cnt.astype(np.int64), kind="stable")[:5]
    To, co, so, io = gs.icp(gm, T[top], 4, 1.0)
    for parts in (1, 2, 7):
        Ts, cs, ss, is_ = gs.icp_sharded(gm, T[top], 4, 1.0, 0, s.n, s.n, emulate_parts=parts)
        assert np.array_equal(_bits(Ts), _bits(To)) and np.array_equal(cs, co) and np.array_equal(is_, io)
        assert np.array_equal(ss, so)
    # a proper sub-range sees fewer correspondences
    Th, ch, *_ = gs.icp_sharded(gm, T[top], 1, 1.0, 0, s.n // 2, s.n)
    assert (ch <= co).all() and ch.sum() < co.sum()


def test_pose_sharded_icp_equals_icp(setup):
    """tm_icp_pose_sharded (SURVEY 8e first option): the slices of any pose split, refined rank by rank with no
    collective, concatenate to tm_icp's result bit for bit; repeated calls replay the cached CUDA graph."""
    name, m, s, om, osc, rec, gm, gs = setup
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T, nthreads=4)
    top = np.argsort(-cnt.astype(np.int64), kind="stable")[:7]
    To, co, so, io = gs.icp(gm, T[top], 4, 1.0)
    for world in (1, 2, 3):
        Tc, cc = np.zeros_like(To), np.zeros_like(co)
        sc, ic = np.zeros_like(so), np.zeros_like(io)
        for rank in range(world):
            Tr, cr, sr, ir = gs.icp_pose_sharded(gm, T[top], 4, 1.0, rank=rank, world=world)
            b, e = (7 * rank) // world, (7 * (rank + 1)) // world
            Tc[b:e], cc[b:e], sc[b:e], ic[b:e] = Tr[b:e], cr[b:e], sr[b:e], ir[b:e]
            assert not Tr[:b].any() and not Tr[e:].any()  # only the slice is written
        assert np.array_equal(_bits(Tc), _bits(To)) and np.array_equal(cc, co) and np.array_equal(ic, io)
        assert np.array_equal(sc, so)
    # graph replay: same inputs, same bits, again and again
    for _ in range(3):
        T2, c2, s2, i2 = gs.icp(gm, T[top], 4, 1.0)
        assert np.array_equal(_bits(T2), _bits(To)) and np.array_equal(c2, co) and np.array_equal(s2, so)


def test_resident_query_and_golden(setup, ctx):
    from triplet_match_b200 import capi
    name, m, s, om, osc, rec, gm, gs = setup
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(rec.outer, g["outer"]) and np.array_equal(rec.pair_j, g["pair_j"])
    q = capi.Query(gs, gm, icp_top_k=4, max_icp_iterations=2)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    n0 = ctx.kernel_launches()
    q.run()
    assert ctx.kernel_launches() - n0 >= 12
    d = q.download()
    r = d["result"]
    assert r.n_hypotheses == g["T"].shape[0] == r.n_scored
    assert r.n_pairs_valid == int(g["valid"].sum())
    assert np.array_equal(_bits(d["T"]), _bits(g["T"]))
    assert np.array_equal(d["hyp_pair"], g["hyp_pair"])
    assert np.array_equal(d["counts"], g["counts"])
    assert np.allclose(d["scores"], g["scores"], rtol=1e-9, atol=1e-9)
    sizes = np.diff(g["sub_off"].astype(np.int64))
    assert r.n_tests == int(sizes[rec.pair_outer[g["hyp_pair"]]].sum())
    assert r.best_inliers == int(g["counts"].max())
    assert r.best_hypothesis == int(np.argmax(g["counts"]))
    assert np.array_equal(_bits(np.array(list(r.best_T), dtype=np.float32)), _bits(g["T"][r.best_hypothesis]))
    ids, Ti, ci, si, it = q.icp_results()
    order = np.argsort(-g["counts"].astype(np.int64), kind="stable")[:4]
    assert np.array_equal(ids, order.astype(np.uint32))
    # early-drop mode of the resident query == golden early-drop outcome
    q.close()
    q = capi.Query(gs, gm, early_out=True)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    d = q.download()
    assert np.array_equal(d["counts"], g["counts_eo"]) and np.array_equal(d["dropped"], g["dropped_eo"])
    q.close()


def test_sharded_query(setup):
    from triplet_match_b200 import capi
    name, m, s, om, osc, rec, gm, gs = setup
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    for world in (2, 3):
        parts, keys = [], []
        for rank in range(world):
            q = capi.Query(gs, gm)
            q.set_shard(rank, world)
            q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
            q.run()
            d = q.download()
            hb, he = capi.shard_range(g["T"].shape[0], rank, world)
            assert d["counts"].shape[0] == he - hb
            parts.append(d["counts"])
            keys.append(int(d["result"].best_key))
            q.close()
        assert np.array_equal(np.concatenate(parts), g["counts"])
        inl, gid = capi.unpack_key(max(keys))  # what the NCCL max all-reduce computes
        assert inl == int(g["counts"].max()) and gid == int(np.argmax(g["counts"]))


def test_hyp_limit_and_capacity(setup):
    from triplet_match_b200 import capi
    name, m, s, om, osc, rec, gm, gs = setup
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    lim = max(1, g["T"].shape[0] // 3)
    q = capi.Query(gs, gm, hyp_limit=lim, max_hypotheses=lim)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    d = q.download()
    assert np.array_equal(d["counts"], g["counts"][:lim])
    q.close()
    q = capi.Query(gs, gm, max_hypotheses=max(1, lim // 2))
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    with pytest.raises(capi.TmError) as e:
        q.result()
    assert e.value.code == capi.TM_ERR_CAPACITY
    q.close()


def test_model_built_by_product_equals_oracle_model(ctx):
    """HostModel (host C++ + GPU voxel fill) == oracle model, then the full query on it."""
    from triplet_match_b200 import capi
    for name in CONFIGS:
        m, s, om, osc, rec = common.config(name)
        hm = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **common.DP, **common.SP)
        assert np.array_equal(hm.voxel, om.voxel)
        k, o, p = om.table(200)
        assert np.array_equal(hm.keys, k) and np.array_equal(hm.pairs, p)
        gm = hm.upload(ctx)
        gs = common.upload_scene(ctx, s)
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        q = capi.Query(gs, gm)
        q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        q.run()
        assert np.array_equal(q.download()["counts"], g["counts"])
        q.close(); gm.close(); gs.close(); hm.close()


def test_device_pair_enumeration_equals_host(ctx):
    """model::init's O(T^2) passes on the device (k_model.cu) vs the host loops: identical bounds,
    insertion-order entries and capped table."""
    import os
    from triplet_match_b200 import capi, synth
    clouds = [common.config(n)[0] for n in ("plane_small", "cylinder_small", "freeform_small")]
    clouds.append(synth.freeform_model(seed=5, n_points=20000, radius=0.01 * np.sqrt(20000 / (4 * np.pi)), n_bumps=9, n_curves=6))
    for m in clouds:
        for dp in (common.DP, dict(distance_step_count=400.0, angle_step=float(np.deg2rad(2.0)))):
            os.environ["TM_MODEL_PAIRS_HOST"] = "1"
            try:
                h = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **dp, **common.SP)
            finally:
                os.environ.pop("TM_MODEL_PAIRS_HOST")
            d = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **dp, **common.SP)
            assert (d.n_subset, d.n_entries, d.n_keys, d.n_kept) == (h.n_subset, h.n_entries, h.n_keys, h.n_kept)
            assert d.n_entries > 1000
            for a in ("feat_min", "feat_max", "keys", "offsets", "pairs"):
                assert np.array_equal(getattr(d, a), getattr(h, a)), a
            ek = lambda x: x._arr(x.lib.tm_hostmodel_entry_keys(x.h), 4 * x.n_entries)
            ep = lambda x: x._arr(x.lib.tm_hostmodel_entry_pairs(x.h), 2 * x.n_entries)
            assert np.array_equal(ek(d), ek(h)) and np.array_equal(ep(d), ep(h))
            h.close(); d.close()


def test_pruned_voxel_fill_equals_brute_force(ctx):
    """The block-pruned two-pass grid fill picks exactly the brute-force winners (incl. ties ->
    lowest index), on grids with partial border blocks, duplicate points and a thin (planar) model."""
    import os
    from triplet_match_b200 import synth
    models = [common.config(n)[0] for n in ("plane_small", "cylinder_small", "freeform_small")]
    models.append(synth.freeform_model(seed=5, n_points=20000, radius=0.01 * np.sqrt(20000 / (4 * np.pi)), n_bumps=9, n_curves=6))
    dup = models[1]
    models.append(synth.Cloud(np.concatenate([dup.pos, dup.pos[::3]]), np.concatenate([dup.nrm, dup.nrm[::3]]),
                              np.concatenate([dup.tgt, dup.tgt[::3]]), np.concatenate([dup.tangent_mask, dup.tangent_mask[::3]])))
    for m in models:
        from triplet_match_b200 import capi
        res = capi.host_resolution(m.pos)
        lo, hi = m.pos.min(0), m.pos.max(0)
        rng_ = (hi - lo).astype(np.float32)
        ext_f = np.maximum(rng_ / np.float32(0.5 * res), 1).astype(np.float32)
        extents = (ext_f + 10).astype(np.int32)
        scale = np.where(rng_ < 1e-5, 1, ext_f / np.where(rng_ < 1e-5, 1, rng_)).astype(np.float32)
        trans = ((scale * (-lo) + np.float32(5)) - np.float32(0.5)).astype(np.float32)
        tv = np.zeros(16, np.float32)
        tv[0], tv[5], tv[10], tv[15] = scale[0], scale[1], scale[2], 1
        tv[12:15] = trans
        os.environ["TM_VOXEL_FILL_BRUTE"] = "1"
        try:
            brute = ctx.voxel_fill(m.pos, m.nrm, m.tgt, extents, tv)
        finally:
            os.environ.pop("TM_VOXEL_FILL_BRUTE")
        pruned = ctx.voxel_fill(m.pos, m.nrm, m.tgt, extents, tv)
        assert np.array_equal(brute, pruned), (m.n, extents, int((brute != pruned).sum()))


def test_surfel_aos_upload_equals_packed(ctx):
    from triplet_match_b200 import capi
    m, s, om, osc, rec = common.config("cylinder_small")
    rec_s = np.zeros((s.n, 12), dtype=np.float32)  # pcl::PointSurfel: xyz1 | nxyz0 | rgba, tangent
    rec_s[:, 0:3], rec_s[:, 3] = s.pos, 1.0
    rec_s[:, 4:7] = s.nrm
    rec_s[:, 9:12] = s.tgt
    gs_a = capi.Scene(ctx, rec_s, None, None, s.tangent_mask, view=capi.surfel_view(rec_s))
    gs_p = common.upload_scene(ctx, s)
    gm = common.upload_model(ctx, m, om)
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    ca, sa, _ = gs_a.score(gm, T[:200])
    cp, sp, _ = gs_p.score(gm, T[:200])
    assert np.array_equal(ca, cp) and np.array_equal(sa, sp)
    gs_a.close(); gs_p.close(); gm.close()


def test_traits_project(ctx):
    from oracle import pyoracle as po
    rng = np.random.default_rng(4)
    xyz = (rng.standard_normal((4000, 3)) * 0.3).astype(np.float32)
    ax = rng.standard_normal(3); ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(0.7) * K + (1 - np.cos(0.7)) * K @ K
    g2l = np.eye(4); g2l[:3, :3] = R; g2l[:3, 3] = [0.1, -0.2, 0.05]
    g16 = g2l.T.astype(np.float32).ravel()  # column-major
    for kind, radius, thr in ((0, 0.3, 0.15), (1, 0.0, 0.2), (2, 0.0, 0.0), (3, 0.0, 0.0)):
        uvw, ok = ctx.traits_project(kind, g16, radius, thr, xyz)
        uo, oo = po.traits_project(kind, g16, radius, thr, xyz)
        assert np.array_equal(ok, oo)
        sel = ok.astype(bool)
        assert 0 < sel.sum()
        assert np.array_equal(_bits(uvw[sel]), _bits(uo[sel])), kind


def test_empty_inputs(ctx):
    from triplet_match_b200 import capi
    m, s, om, osc, rec = common.config("plane_small")
    gm = common.upload_model(ctx, m, om)
    gs = common.upload_scene(ctx, s)
    e = np.zeros(0, dtype=np.uint32)
    f, k, v = gs.features(gm, e, e, 0.2, 1.0)
    assert v.size == 0
    off, hits = gm.probe(np.zeros((0, 4), np.uint32), None, 200)
    assert off.tolist() == [0] and hits.shape[0] == 0
    c, sc, d = gs.score(gm, np.zeros((0, 16), np.float32))
    assert c.size == 0
    q = capi.Query(gs, gm)
    q.set_pairs(e, e, e)
    q.run()
    r = q.result()
    assert r.n_hypotheses == 0 and r.best_key == 0
    q.close()
    # a pair list whose pairs all fail the filter -> zero hypotheses, no crash
    q = capi.Query(gs, gm)
    nt = np.nonzero(s.tangent_mask == 0)[0][:4].astype(np.uint32)
    q.set_pairs(nt[:1], np.zeros(3, np.uint32), nt[1:4])
    q.run()
    assert q.result().n_hypotheses == 0
    q.close(); gm.close(); gs.close()


def test_full_size_properties(ctx):
    """C2-sized run (1M-point scene): properties that do not need the CPU oracle at full
    size, plus an oracle spot check on a sample of hypotheses."""
    from oracle import pyoracle as po
    from triplet_match_b200 import capi, synth, workloads
    model, scene = workloads.c2_clouds()
    hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **common.DP, **common.SP)
    gm = hm.upload(ctx)
    gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
    rec = synth.record_pairs(2, scene, hm.diameter, n_outer=24, pairs_per_outer=64)
    q = capi.Query(gs, gm)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    d = q.download()
    n = d["counts"].shape[0]
    assert n > 20000
    # (1) two independent kernels agree: register-tiled full scorer vs warp-per-hypothesis scorer
    off, idx = gs.ball_subsets(rec.outer, hm.diameter)
    sel = np.linspace(0, n - 1, 3000).astype(np.int64)
    hyp_sub = rec.pair_outer[d["hyp_pair"][sel]]
    c2, s2, dr = gs.score(gm, d["T"][sel], hyp_sub, off, idx, early_out=True, accept_prob=0.0)
    assert not dr.any()  # accept_prob 0 => the drop rule can never fire
    assert np.array_equal(c2, d["counts"][sel]) and np.allclose(s2, d["scores"][sel], rtol=1e-9, atol=1e-12)
    # (2) idempotence and (3) sharding invariance
    q.run()
    assert np.array_equal(q.download_counts()[0], d["counts"])
    parts = []
    for rank in range(2):
        q2 = capi.Query(gs, gm)
        q2.set_shard(rank, 2)
        q2.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        q2.run()
        parts.append(q2.download_counts()[0])
        q2.close()
    assert np.array_equal(np.concatenate(parts), d["counts"])
    # (3b) shards of equal TESTS (tm_query_set_balance): still a partition of the list in rank order, and the
    # per-rank test counts are within 2 % of each other (the count-based split is not); changing the shard
    # after set_pairs re-sizes on the next run
    for world in (2, 3, 8):
        parts, tests, tests_cnt = [], [], []
        for rank in range(world):
            q2 = capi.Query(gs, gm)
            q2.set_shard(rank, world)
            q2.set_balance(True)
            q2.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
            q2.run()
            parts.append(q2.download_counts()[0])
            tests.append(int(q2.result().n_tests))
            q2.set_balance(False)  # re-sized lazily by the next run
            q2.run()
            tests_cnt.append(int(q2.result().n_tests))
            q2.close()
        assert np.array_equal(np.concatenate(parts), d["counts"])
        assert sum(tests) == sum(tests_cnt) == int(q.result().n_tests)
        assert max(tests) <= 1.02 * (sum(tests) / world)
        # the device's cut points are the host mirror's (capi.balanced_bounds)
        nh = np.bincount(rec.pair_outer[d["hyp_pair"]], minlength=rec.outer.size)
        sizes = np.diff(off.astype(np.int64))
        assert np.concatenate([[0], np.cumsum([p.size for p in parts])]).tolist() == capi.balanced_bounds(nh, sizes, world)
    q2 = capi.Query(gs, gm)
    q2.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)  # sized for (0, 1)
    q2.set_shard(1, 2)                                    # then re-sharded
    q2.run()
    assert np.array_equal(q2.download_counts()[0], d["counts"][(n + 1) // 2:])
    q2.close()
    # (4) every hypothesis sees its own pair: p1 -> m_i exactly => at least one inlier
    assert (d["counts"][d["valid"].astype(bool)] >= 1).all()
    # (5) oracle spot check at full size
    om = po.OModel(model, **common.DP, **common.SP, resolution=hm.resolution)
    assert np.array_equal(om.voxel, hm.voxel)
    osc = po.OScene(scene)
    pick = np.concatenate([np.argsort(-d["counts"].astype(np.int64))[:8], sel[::100]])
    co, so, _ = osc.score_batch(om, d["T"][pick], rec.pair_outer[d["hyp_pair"][pick]], off, idx, nthreads=os.cpu_count())
    assert np.array_equal(co, d["counts"][pick])
    q.close(); gm.close(); gs.close(); hm.close()
