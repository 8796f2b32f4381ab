"""Dev probe: C2 with project_(early_out=true) — the mode find_in_subset actually runs."""
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
rec = bench_wl.c2_record(scene, hm.diameter, 1)
for eo in (False, True):
    q = capi.Query(gs, gm, **bench_wl.QP, early_out=eo, hyp_limit=1 << 20, max_hypotheses=1 << 20)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    for _ in range(3):
        ctx.flush_l2(); ctx.timer_start(); q.run(); ms = ctx.timer_stop()
    r = q.result()
    d = q.download()
    print(f"early_out={eo}: {ms:.2f} ms (score kernel {q.score_kernel_ms():.2f}), tests {r.n_tests:.3e}, hyps/s {r.n_scored / ms * 1e3:.3e}, "
          f"best {r.best_inliers} @ {r.best_hypothesis}, dropped {int(d['dropped'].sum())} of {d['dropped'].size}", flush=True)
    q.close()
