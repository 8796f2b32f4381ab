"""Dev probe: cull / in-grid / inlier statistics of one C2 scoring pass (TM_SCORE_STATS=1)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TM_SCORE_STATS"] = "1"
from triplet_match_b200 import workloads as bench_wl
from triplet_match_b200 import capi

ctx = capi.Context(0)
model, scene = bench_wl.c2_clouds()
hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **bench_wl.DP, min_df=0.2, max_df=1.0, cap=200)
gm = hm.upload(ctx)
gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
rec = bench_wl.c2_record(scene, hm.diameter, 1)
q = capi.Query(gs, gm, **bench_wl.QP, hyp_limit=1 << 20, max_hypotheses=1 << 20)
q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
for _ in range(2):
    ctx.timer_start(); q.run(); ms = ctx.timer_stop()
    r = q.result()
    print("ms", ms, "tests", r.n_tests, flush=True)
c = q.download_counts()[0]
print("sum inliers", int(c.astype(np.int64).sum()), "hyps", c.size, "mean", c.mean(), "max", c.max(),
      "hyps with 0", int((c == 0).sum()), "hyps>1000", int((c > 1000).sum()), "hyps>5000", int((c > 5000).sum()))
print("inlier fraction of tests", c.astype(np.int64).sum() / r.n_tests)
