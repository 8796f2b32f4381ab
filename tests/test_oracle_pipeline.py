"""CPU tests: the oracle against (a) an independent numpy float32 restatement of project_,
(b) the committed golden vectors, (c) structural properties of the model table."""
import os

import numpy as np
import pytest

import common
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["plane_small", "cylinder_small", "freeform_small"]
ALL_CONFIGS = CONFIGS + ["plane_small_shuffled", "cylinder_small_shuffled"]


def numpy_project_counts(m, s, om, T16, subset, dist_thres=1.0, mask=None):
    """Independent float32 restatement of scene.hpp:411-510 (early_out = false), vectorised:
    every + and * is a separate float32 rounding, in the Eigen order stated in DESIGN.md."""
    f = np.float32
    T = np.asarray(T16, dtype=np.float32).reshape(4, 4).T  # column-major -> T[r][c]
    P = s.pos[subset]
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    tp = [((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(3)]
    sc, tr = om.scale.astype(f), om.trans.astype(f)
    v = [sc[k] * tp[k] + tr[k] for k in range(3)]
    with np.errstate(invalid="ignore"):
        ijk = [np.trunc(np.where(np.isfinite(a), a, -5.0)).astype(np.int64) for a in v]
    ex = om.extents.astype(np.int64)
    inb = np.ones(P.shape[0], dtype=bool)
    for k in range(3):
        inb &= (ijk[k] >= 0) & (ijk[k] < ex[k]) & np.isfinite(v[k])
    if mask is not None:
        inb &= mask[subset] == 0
    lin = (ijk[2] * ex[0] * ex[1] + ijk[1] * ex[0] + ijk[0])[inb]
    mi = om.voxel[lin]
    mp = m.pos[mi]
    dx, dy, dz = tp[0][inb] - mp[:, 0], tp[1][inb] - mp[:, 1], tp[2][inb] - mp[:, 2]
    dist = np.sqrt(dx * dx + (dy * dy + dz * dz))
    thres = f(dist_thres) * f(om.resolution)
    near = ~(dist > thres)
    mt = m.tgt[mi]
    is_t = np.sqrt(mt[:, 0] * mt[:, 0] + (mt[:, 1] * mt[:, 1] + mt[:, 2] * mt[:, 2])) > f(0.7)
    use_t = s.tangent_mask[subset][inb] != 0
    inl = near & (is_t == use_t)
    return int(inl.sum()), subset[inb][inl], mi[inl]


@pytest.mark.parametrize("name", CONFIGS)
def test_oracle_vs_numpy_restatement(name):
    m, s, om, osc, rec = common.config(name)
    T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    rng = np.random.default_rng(0)
    pick = rng.choice(T.shape[0], size=min(60, T.shape[0]), replace=False)
    # always include the best hypotheses (many inliers) besides random ones
    cnt_all, _, _ = osc.score_batch(om, T, nthreads=4)
    pick = np.unique(np.concatenate([pick, np.argsort(-cnt_all.astype(np.int64))[:10]]))
    mask = (rng.random(s.n) < 0.1).astype(np.uint8)
    for h in pick:
        sub = osc.ball_subset(int(rec.outer[rec.pair_outer[hp[h]]]), om.diameter)
        r = osc.project(om, sub, T[h])
        n, sc, mc = numpy_project_counts(m, s, om, T[h], sub)
        assert r["count"] == n
        assert np.array_equal(r["scene_corrs"], sc) and np.array_equal(r["model_corrs"], mc)
    osc.set_mask(mask)
    try:
        for h in pick[:15]:
            sub = osc.ball_subset(int(rec.outer[rec.pair_outer[hp[h]]]), om.diameter)
            assert osc.project(om, sub, T[h])["count"] == numpy_project_counts(m, s, om, T[h], sub, mask=mask)[0]
    finally:
        osc.set_mask(np.zeros(s.n, dtype=np.uint8))


@pytest.mark.parametrize("name", ALL_CONFIGS)
def test_oracle_reproduces_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = mg.golden_for(name)
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert set(ref.files) == set(g.keys())
    for k in ref.files:
        a, b = np.asarray(g[k]), ref[k]
        if a.dtype.kind == "f":
            assert np.array_equal(a.view(np.uint32 if a.dtype == np.float32 else np.uint64),
                                  b.view(np.uint32 if b.dtype == np.float32 else np.uint64)), k
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("name", CONFIGS)
def test_table_order_and_query_limit(name):
    """equal_range order is reverse insertion order in libstdc++ and query() is
    capped by query_limit (scene.hpp:310)."""
    m, s, om, osc, rec = common.config(name)
    keys, off, pairs = om.table(0)
    assert off[-1] == om.n_entries and keys.shape[0] == om.n_keys
    sub = om.subset
    rank = {int(v): i for i, v in enumerate(sub)}
    for k in range(min(keys.shape[0], 50)):
        seg = pairs[off[k]:off[k + 1]]
        order = [rank[int(a)] * len(sub) + rank[int(b)] for a, b in seg]
        assert order == sorted(order, reverse=True)  # LIFO of the (i outer, j inner) double loop
    keys_c, off_c, pairs_c = om.table(5)
    assert np.all(np.diff(off_c.astype(np.int64)) <= 5)
    for k in range(min(keys.shape[0], 50)):
        assert np.array_equal(pairs_c[off_c[k]:off_c[k + 1]], pairs[off[k]:off[k] + min(5, off[k + 1] - off[k])])


def test_early_drop_semantics_small_subsets():
    """Checkpoint chaining for tiny subsets (tests[] repeats / zeros) and empty subsets."""
    m, s, om, osc, rec = common.config("plane_small_shuffled")
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    sub_full = osc.ball_subset(int(rec.outer[0]), om.diameter)
    for n in (0, 1, 2, 5, 19, 20, 21, 40, 333):
        sub = sub_full[:n]
        for h in range(0, min(T.shape[0], 40), 7):
            a = osc.project(om, sub, T[h], early_out=True)
            b = osc.project(om, sub, T[h], early_out=False)
            assert a["count"] <= b["count"]
            if not a["dropped"]:
                assert a["count"] == b["count"] and abs(a["score"] - b["score"]) < 1e-12
            else:
                assert a["saved"] >= 0 and a["saved"] < max(n, 1)
