"""One pass over every kernel of the path on the C2 workload (for an ncu --metrics capture):
model::init, scene upload + tangent-mask pre-processing, full query, early-drop queries (subset order; even walk, staged), ICP,
correspondence lists."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from triplet_match_b200 import workloads as bench_wl
from triplet_match_b200 import capi

ctx = capi.Context(0)
model, scene = bench_wl.c2_clouds()
hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **bench_wl.DP, min_df=0.2, max_df=1.0, cap=200)
gm = hm.upload(ctx)
gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
gs.compute_tangent_mask(30, 0.2, apply=False)
rec = bench_wl.c2_record(scene, hm.diameter, 1)
for eo in (0, 1, 2):
    q = capi.Query(gs, gm, **bench_wl.QP, early_out=eo, hyp_limit=1 << 20, max_hypotheses=1 << 20, icp_top_k=64, max_icp_iterations=5)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    q.run()
    r = q.result()
    print("early_out", eo, "scored", r.n_scored, "tests", r.n_tests, "best", r.best_inliers, flush=True)
    q.close()
d = np.eye(4, dtype=np.float32).T.reshape(1, 16)
gs.correspondences_batch(gm, np.repeat(d, 4, axis=0), 1.0)
