"""C3 probe (dev): free-form 50 k model vs 10 M scene, 2^20 hypotheses; prints step / kernel time and the cull statistics
(TM_SCORE_STATS=1).  Run under ncu for the kernel's profile."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
from triplet_match_b200 import capi, workloads as wl

ctx = capi.Context(0)
model, scene, poses = wl.c3_clouds()
hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **wl.DP, min_df=0.2, max_df=1.0, cap=200)
gm = hm.upload(ctx)
gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
rec = wl.c3_record(scene, hm.diameter, 1)
q = capi.Query(gs, gm, **wl.QP, hyp_limit=1 << 20, max_hypotheses=1 << 20)
q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
for it in range(int(os.environ.get("STEPS", 3))):
    ctx.flush_l2(); ctx.timer_start(); q.run(); ms = ctx.timer_stop(); r = q.result()
    print(f"step {it}: {ms:.2f} ms (kernel {q.score_kernel_ms():.2f}), hyps {r.n_scored}, tests {r.n_tests:.3e}, best {r.best_inliers}", flush=True)
