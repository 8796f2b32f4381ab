"""Randomised pin of the oracle against the reference's own sources (oracle/_ref): beyond the three
structured configurations of test_oracle_vs_ref.py, small random clouds (irregular density, random
unit tangents on a random subset, duplicate and near-duplicate points, masked points) and random
near-rigid as well as arbitrary transforms.  Everything is compared bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from triplet_match_b200 import synth
import test_oracle_vs_ref as tr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libtm_ref.so")
F = np.float32


def _cloud(rng, n, extent, n_tan, dup=0):
    pos = (rng.random((n, 3)) * extent).astype(F)
    pos[:, 2] *= F(0.3)  # flattened blob: more pairs inside the distance window
    if dup:
        pos[-dup:] = pos[:dup]                      # exact duplicates
        pos[-2 * dup:-dup] = pos[dup:2 * dup] + F(1e-7)  # near duplicates
    nrm = rng.standard_normal((n, 3)).astype(F)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tgt = np.zeros((n, 3), F)
    tm = np.zeros(n, np.uint8)
    sel = rng.choice(n, n_tan, replace=False)
    t = rng.standard_normal((n_tan, 3)).astype(F)
    tgt[sel] = t / np.linalg.norm(t, axis=1, keepdims=True)
    tgt[sel[: n_tan // 8]] *= F(0.5)                # tangents below the 0.7 norm threshold
    tm[sel[n_tan // 8:]] = 1
    return synth.Cloud(pos, nrm, tgt, tm)


@pytest.fixture(scope="module")
def ref(built):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libtm_ref.so not built (no /root/reference in this environment)")
    return C.CDLL(REF)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_clouds(ref, seed):
    rng = np.random.default_rng(seed)
    m = _cloud(rng, 260 + 40 * seed, 0.25, 60 + 5 * seed, dup=3)
    s = _cloud(rng, 2500, 0.6, 300)
    s.pos[:m.n] = m.pos + F(0.17)  # a translated copy of the model inside the scene
    om = po.OModel(m, distance_step_count=20.0, angle_step=0.17453292, min_df=0.2, max_df=1.0)
    osc = po.OScene(s)
    rm, rs = tr.RefModel(ref, m), tr.RefScene(ref, s)
    # model::init
    assert np.float32(rm.resolution) == np.float32(om.resolution) and np.float32(rm.diameter) == np.float32(om.diameter)
    assert np.array_equal(rm.extents, om.extents)
    assert np.array_equal(rm.to_voxel16.view(np.uint32), om.to_voxel16.view(np.uint32))
    assert np.array_equal(rm.feat_min.view(np.uint32), om.feat_min.view(np.uint32))
    assert np.array_equal(rm.feat_max.view(np.uint32), om.feat_max.view(np.uint32))
    # features / keys / hits (equal_range order, two limits) on random scene pairs
    tidx = np.flatnonzero(s.tangent_mask)
    pi = rng.choice(tidx, 400).astype(np.uint32)
    pj = rng.choice(tidx, 400).astype(np.uint32)
    feats, keys, valid = osc.pair_features(om, pi, pj)
    n_hits = 0
    for k in np.flatnonzero(valid)[:80]:
        for limit in (200, 5):
            a, b = rm.query(feats[k], limit), om.query(feats[k], limit)
            assert np.array_equal(a, b)
            n_hits += a.shape[0]
    assert n_hits > 0
    # base_transform_ on arbitrary inputs
    for _ in range(100):
        inp = rng.standard_normal(18).astype(F)
        out = np.zeros(16, F)
        rs.L.ref_base_transform(rs.h, tr._p(inp), tr._p(out))
        exp = po.base_transform(inp[0:3], inp[3:6], inp[6:9], inp[9:12], inp[12:15], inp[15:18])
        assert np.array_equal(out.view(np.uint32), exp.view(np.uint32))
    # project_: near the true pose (many inliers), random rigid, and arbitrary affine transforms
    Ts = []
    for k in range(24):
        T = np.eye(4)
        if k < 12:
            a = rng.standard_normal(3) * 0.02
            T[:3, :3] += np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            T[:3, 3] = -0.17 + rng.standard_normal(3) * 0.003
        elif k < 18:
            q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            T[:3, :3] = q
            T[:3, 3] = rng.standard_normal(3) * 0.2
        else:
            T[:3, :4] = rng.standard_normal((3, 4))
        Ts.append(T.T.reshape(-1).astype(F))
    sub_all = np.arange(s.n, dtype=np.int32)
    some = False
    for T16 in Ts:
        for sub in (sub_all, rng.permutation(sub_all)[:777]):
            for eo in (False, True):
                a = rs.project(rm, sub, T16, early_out=eo)
                b = osc.project(om, sub, T16, early_out=eo)
                assert a["count"] == b["count"] and a["saved"] == b["saved"] and a["score"] == b["score"]
                assert np.array_equal(a["scene_corrs"], b["scene_corrs"]) and np.array_equal(a["model_corrs"], b["model_corrs"])
                some = some or a["count"] > 50
    assert some
    # icp_ control flow from the good start poses
    for T16 in Ts[:3]:
        out, score = np.zeros(16, F), C.c_double()
        n = rs.L.ref_icp(rs.h, rm.h, tr._p(T16), C.c_uint32(4), C.c_float(1.0), C.c_float(0.5), tr._p(out), C.byref(score))
        oT, on, osx, oit = osc.icp(om, T16, 4, 1.0)
        assert n == on and np.array_equal(out.view(np.uint32), oT.view(np.uint32)) and score.value == osx
