"""early_out = 2 on C2 (C3=1: on C3) (dev): the level-by-level tiled drop test (default) against the one-warp-per-hypothesis walker
(TM_EARLY_LEVELS=0): step / scoring time, tests, survivors, hypotheses walked one by one, and a checksum of the
outcome (counts, drop flags) that must be the same in both modes."""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
from triplet_match_b200 import capi, workloads as wl

ctx = capi.Context(0)
C3 = bool(os.environ.get("C3"))  # the free-form 50 k model / 10 M scene instead (occupancy-mask instantiations)
if C3:
    model, scene, _ = wl.c3_clouds()
else:
    model, scene = wl.c2_clouds()
hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **wl.DP, min_df=wl.QP["min_df"],
                    max_df=wl.QP["max_df"], cap=wl.QP["query_limit"])
gm = hm.upload(ctx)
gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
rec = wl.c3_record(scene, hm.diameter, 1) if C3 else wl.c2_record(scene, hm.diameter, 1)
H = wl.HYP_PER_GPU
q = capi.Query(gs, gm, **wl.QP, early_out=2, hyp_limit=H, max_hypotheses=H)
q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
for it in range(int(os.environ.get("STEPS", 4))):
    ctx.flush_l2(); ctx.timer_start(); q.run(); ms = ctx.timer_stop(); r = q.result()
    print(f"step {it}: {ms:.2f} ms (scoring {q.score_kernel_ms():.2f}), hyps {r.n_scored}, tests {r.n_tests:.4e}, "
          f"best {r.best_inliers} @ {r.best_hypothesis}, walked one by one {q.early_walked()}", flush=True)
d = q.download()
print("levels" if os.environ.get("TM_EARLY_LEVELS", "1") != "0" else "walker", "survivors", int((d["dropped"] == 0).sum()),
      "crc counts %08x dropped %08x scores %08x" % (zlib.crc32(d["counts"].tobytes()), zlib.crc32(d["dropped"].tobytes()),
                                                    zlib.crc32(d["scores"].tobytes())), flush=True)
