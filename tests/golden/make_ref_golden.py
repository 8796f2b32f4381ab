"""Generates tests/golden/ref_*.npz from the REFERENCE's own sources (not from the oracle):
oracle/_ref/libtm_ref.so = /root/reference's model/scene/feature/discretize code compiled where
it lies against the header stand-ins of oracle/shim (recipe: oracle/Makefile target `ref`).
Run from the repo root in a container that has /root/reference:
    make -C oracle ref && python tests/golden/make_ref_golden.py
The inputs (clouds, recorded pairs) are the seed-fixed small configurations of tests/common.py;
only input *selection* (which pairs / hypotheses to record) uses the oracle, every recorded
output value comes from the reference code.  tests/test_ref_golden.py checks the oracle (CPU)
and the CUDA path (-m gpu) against these files; they travel to the GPU box, /root/reference
does not."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import test_oracle_vs_ref as tr  # noqa: E402  (ctypes wrappers RefModel / RefScene)

N_HYP = 96      # hypotheses recorded per configuration
N_ICP = 3


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_golden(L, name):
    m, s, om, osc, rec = common.config(name)
    rm, rs = tr.RefModel(L, m), tr.RefScene(L, s)
    n_pairs = rec.pair_i.shape[0]
    # a1-a4: feature / valid / key of every recorded pair through the reference's functions
    feats = np.zeros((n_pairs, 4), np.float32)
    keys = np.zeros((n_pairs, 4), np.uint32)
    valid = np.zeros(n_pairs, np.uint8)
    for k in range(n_pairs):
        i, j = int(rec.pair_i[k]), int(rec.pair_j[k])
        inp = np.concatenate([s.pos[i], s.tgt[i], s.pos[j], s.tgt[j]]).astype(np.float32)
        L.ref_feature(_p(inp), _p(feats[k]))
        valid[k] = L.ref_valid(_p(feats[k]), _p(rm.feat_min), _p(rm.feat_max))
        L.ref_discretize_feature(_p(feats[k]), _p(rm.feat_min), _p(rm.feat_max), C.c_float(20.0),
                                 C.c_float(0.17453292), _p(keys[k]))
    # a5: hash hits (equal_range order, query_limit 200) for every recorded pair's feature
    hit_off = np.zeros(n_pairs + 1, np.uint64)
    hits = []
    for k in range(n_pairs):
        h = rm.query(feats[k], 200)
        hits.append(h)
        hit_off[k + 1] = hit_off[k] + h.shape[0]
    hits = np.concatenate(hits).astype(np.uint32) if hits else np.zeros((0, 2), np.uint32)
    # a6: base_transform_ for a spread of (pair, hit) combinations
    rng = np.random.default_rng(17)
    pair_of_hit = np.repeat(np.arange(n_pairs), np.diff(hit_off).astype(np.int64))
    # choose hypotheses whose pair passes the scene-side filters the caller applies (scene.hpp:290-302)
    T_o, hp_o, mi_o, mj_o, _ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt_o, _, _ = osc.score_batch(om, T_o, nthreads=4)
    pick = np.unique(np.concatenate([np.argsort(-cnt_o.astype(np.int64), kind="stable")[:N_HYP // 3],
                                     rng.choice(T_o.shape[0], size=min(N_HYP, T_o.shape[0]), replace=False)]))[:N_HYP]
    T = np.zeros((pick.size, 16), np.float32)
    hyp_in = np.zeros((pick.size, 18), np.float32)
    for q, h in enumerate(pick):
        i, j = rec.pair_i[hp_o[h]], rec.pair_j[hp_o[h]]
        hyp_in[q] = np.concatenate([s.pos[i], s.pos[j], s.tgt[i], m.pos[mi_o[h]], m.pos[mj_o[h]], m.tgt[mi_o[h]]])
        L.ref_base_transform(rs.h, _p(hyp_in[q]), _p(T[q]))
    # a8-a11: project_ over the recorded ball subset, both early_out modes
    outer_of = rec.pair_outer[hp_o[pick]]
    counts = np.zeros(pick.size, np.uint32)
    scores = np.zeros(pick.size, np.float64)
    counts_eo = np.zeros(pick.size, np.uint32)
    scores_eo = np.zeros(pick.size, np.float64)
    saved_eo = np.zeros(pick.size, np.uint32)
    corr_crc = np.zeros(pick.size, np.uint64)
    for q in range(pick.size):
        sub = osc.ball_subset(int(rec.outer[outer_of[q]]), float(rm.diameter))
        a = rs.project(rm, sub, T[q], early_out=False)
        b = rs.project(rm, sub, T[q], early_out=True)
        counts[q], scores[q] = a["count"], a["score"]
        counts_eo[q], scores_eo[q], saved_eo[q] = b["count"], b["score"], b["saved"]
        corr_crc[q] = np.uint64(int((a["scene_corrs"].astype(np.uint64) * 31 + a["model_corrs"].astype(np.uint64)).sum()))
    # a12: icp_ (the rigid solve is the stand-in's; control flow / thresholds are the reference's)
    best = np.argsort(-counts.astype(np.int64), kind="stable")[:N_ICP]
    icp_T = np.zeros((best.size, 2, 16), np.float32)
    icp_n = np.zeros((best.size, 2), np.uint32)
    for q, b in enumerate(best):
        for w, iters in enumerate((1, 5)):
            sc = C.c_double()
            icp_n[q, w] = L.ref_icp(rs.h, rm.h, _p(T[b]), C.c_uint32(iters), C.c_float(1.0), C.c_float(0.5),
                                    _p(icp_T[q, w]), C.byref(sc))
    # a9: voxel_query at a spread of positions (cell interiors, so centre rounding cannot matter)
    ex = rm.extents.astype(np.int64)
    cells = rng.choice(int(ex.prod()), size=min(3000, int(ex.prod())), replace=False)
    vq_pos = np.zeros((cells.size, 4), np.float32)
    vq = np.zeros(cells.size, np.int64)
    for q, lin in enumerate(cells):
        k, r = divmod(int(lin), int(ex[0] * ex[1]))
        j, i = divmod(r, int(ex[0]))
        c = (np.array([i, j, k], np.float32) + np.float32(0.25) - om.trans) / om.scale
        vq_pos[q] = [c[0], c[1], c[2], 1.0]
        got = rm.voxel_query(vq_pos[q])
        vq[q] = -1 if got is None else got
    return dict(scene_n=np.int64(s.n), model_n=np.int64(m.n),
                scene_pos_crc=np.uint64(int(s.pos.view(np.uint32).astype(np.uint64).sum())),
                resolution=np.float32(rm.resolution), diameter=np.float32(rm.diameter),
                extents=rm.extents.astype(np.int32), margin=np.int32(rm.margin), to_voxel16=rm.to_voxel16,
                feat_min=rm.feat_min, feat_max=rm.feat_max, point_count=np.int64(rm.point_count),
                pair_i=rec.pair_i, pair_j=rec.pair_j, feats=feats, keys=keys, valid=valid,
                hit_off=hit_off, hits=hits,
                hyp_in=hyp_in, T=T, hyp_outer=outer_of.astype(np.int64), outer=rec.outer,
                counts=counts, scores=scores, counts_eo=counts_eo, scores_eo=scores_eo, saved_eo=saved_eo,
                corr_crc=corr_crc, icp_src=best.astype(np.int64), icp_T=icp_T, icp_n=icp_n,
                vq_pos=vq_pos, vq=vq)


if __name__ == "__main__":
    lib = os.path.join(ROOT, "oracle", "_ref", "libtm_ref.so")
    L = C.CDLL(lib)
    L.ref_valid.restype = C.c_int
    for name in ("plane_small", "cylinder_small", "freeform_small", "plane_small_shuffled",
                 "cylinder_small_shuffled"):
        g = ref_golden(L, name)
        np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), **g)
        print(name, "pairs", g["keys"].shape[0], "valid", int(g["valid"].sum()), "hits", g["hits"].shape[0],
              "hyps", g["T"].shape[0], "max count", int(g["counts"].max()), "eo kept",
              int((g["saved_eo"] == 0).sum()), "icp", g["icp_n"].tolist())
