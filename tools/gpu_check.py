"""Development check on a real GPU: stage-by-stage parity against the oracle, verbose."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from triplet_match_b200 import capi

def main():
    ctx = capi.Context(0)
    print("SMs", ctx.sm_count)
    for name in ("plane_small", "cylinder_small", "freeform_small"):
        m, s, om, osc, rec = common.config(name)
        gm = common.upload_model(ctx, m, om); gs = common.upload_scene(ctx, s)
        f, k, v = gs.features(gm, rec.pair_i, rec.pair_j, 0.2, 1.0)
        fo, ko, vo = osc.pair_features(om, rec.pair_i, rec.pair_j)
        print(name, "pairs", v.size, "valid", int(v.sum()), "valid==", np.array_equal(v, vo),
              "keys==", np.array_equal(k, ko), "feats==", np.array_equal(f.view(np.uint32), fo.view(np.uint32)))
        off, hits = gm.probe(k, v, 200)
        T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
        print("  hits", hits.shape[0], "oracle hyps", T.shape[0], "hits==", np.array_equal(hits, np.stack([mi, mj], 1)))
        Tg, vg = gs.hypotheses(gm, rec.pair_i, rec.pair_j, off, hits)
        print("  T==", np.array_equal(Tg.view(np.uint32), T.view(np.uint32)),
              "maxabs", float(np.nanmax(np.abs(Tg - T))) if T.size else 0)
        boff, bidx = gs.ball_subsets(rec.outer, om.diameter)
        ok = True
        for o in range(rec.outer.size):
            ok &= np.array_equal(bidx[int(boff[o]):int(boff[o+1])], osc.ball_subset(int(rec.outer[o]), om.diameter))
        print("  subsets==", ok, "total", int(boff[-1]))
        hyp_sub = rec.pair_outer[hp]
        for eo in (False, True):
            cg, sg, dg = gs.score(gm, T, hyp_sub, boff, bidx, early_out=eo)
            co, so, do = osc.score_batch(om, T, hyp_sub, boff, bidx, early_out=eo, nthreads=8)
            print("  early_out", eo, "counts==", np.array_equal(cg, co), "dropped==", np.array_equal(dg, do),
                  "score maxerr", float(np.max(np.abs(sg - so))) if T.size else 0, "max count", int(co.max()), "ndropped", int(do.sum()))
            if not np.array_equal(cg, co):
                bad = np.nonzero(cg != co)[0]; print("   bad", bad[:10], cg[bad[:10]], co[bad[:10]])
        ca, sa, _ = gs.score(gm, T[:64])
        coa, soa, _ = osc.score_batch(om, T[:64], nthreads=8)
        print("  all-scene counts==", np.array_equal(ca, coa), float(np.max(np.abs(sa - soa))))
        best = T[int(np.argmax(co))]
        sc, mc, scv = gs.correspondences(gm, best, 1.0)
        pr = osc.project(om, np.arange(s.n, dtype=np.int32), best)
        print("  corrs==", np.array_equal(sc, pr["scene_corrs"]) and np.array_equal(mc, pr["model_corrs"]), sc.size, abs(scv - pr["score"]))
        top = np.argsort(-co.astype(np.int64), kind="stable")[:4]
        To, cnt, scr, it = gs.icp(gm, T[top], 5, 1.0)
        for r, h in enumerate(top):
            oT, on, osx, oit = osc.icp(om, T[h], 5, 1.0)
            print("  icp", r, "gpu n", int(cnt[r]), "it", int(it[r]), "| oracle n", on, "it", oit, "| dT", float(np.max(np.abs(To[r] - oT))))
        # resident query
        q = capi.Query(gs, gm, icp_top_k=4, max_icp_iterations=5)
        q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j); q.run(); d = q.download()
        print("  query counts==", np.array_equal(d["counts"], co if False else osc.score_batch(om, T, hyp_sub, boff, bidx, nthreads=8)[0]),
              "best", d["result"].best_inliers, d["result"].best_hypothesis, "tests", d["result"].n_tests, "valid pairs", d["result"].n_pairs_valid)
        vf = ctx.voxel_fill(m.pos, m.nrm, m.tgt, om.extents, om.to_voxel16)
        print("  voxel_fill==", np.array_equal(vf, om.voxel), int((vf != om.voxel).sum()))
        q.close(); gm.close(); gs.close()
    ctx.close()

if __name__ == "__main__":
    main()
