"""Generates tests/golden/*.npz with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors
(SURVEY §4), so these pin the oracle's outputs on the seed-fixed small
configurations: keys, hash hits, transforms, subsets, inlier counts, scores.
The CUDA path is checked against the same files under -m gpu."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402


def golden_for(name: str) -> dict:
    m, s, om, osc, rec = common.config(name)
    feats, keys, valid = osc.pair_features(om, rec.pair_i, rec.pair_j)
    T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    subs = [osc.ball_subset(int(o), om.diameter) for o in rec.outer]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in subs])
    idx = np.concatenate(subs).astype(np.int32)
    hyp_sub = rec.pair_outer[hp]
    counts, scores, _ = osc.score_batch(om, T, hyp_sub, off, idx, nthreads=4)
    counts_eo, scores_eo, dropped_eo = osc.score_batch(om, T, hyp_sub, off, idx, early_out=True,
                                                       nthreads=4)
    kk, oo, pp = om.table(200)
    return dict(scene_n=np.int64(s.n), model_n=np.int64(m.n),
                scene_pos_crc=np.uint64(int(s.pos.view(np.uint32).astype(np.uint64).sum())),
                extents=om.extents, to_voxel16=om.to_voxel16, resolution=np.float32(om.resolution),
                diameter=np.float32(om.diameter), feat_min=om.feat_min, feat_max=om.feat_max,
                voxel_crc=np.uint64(int((om.voxel.astype(np.uint64) * (np.arange(om.voxel.size, dtype=np.uint64) % 65521 + 1)).sum())),
                table_keys=kk, table_offsets=oo, table_pairs=pp,
                outer=rec.outer, pair_outer=rec.pair_outer, pair_j=rec.pair_j,
                keys=keys, valid=valid, feats=feats * valid[:, None],
                T=T, hyp_pair=hp, hyp_mi=mi, hyp_mj=mj,
                sub_off=off, sub_crc=np.uint64(int(idx.astype(np.uint64).sum())),
                counts=counts, scores=scores, counts_eo=counts_eo, dropped_eo=dropped_eo)


if __name__ == "__main__":
    for name in ("plane_small", "cylinder_small", "freeform_small", "plane_small_shuffled",
                 "cylinder_small_shuffled"):
        g = golden_for(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
        print(name, "pairs", g["keys"].shape[0], "valid", int(g["valid"].sum()), "hyps",
              g["T"].shape[0], "max count", int(g["counts"].max()), "early-out dropped",
              int(g["dropped_eo"].sum()), "kept", int((1 - g["dropped_eo"]).sum()))
