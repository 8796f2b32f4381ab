"""Dev probe: the 8 test-balanced shards of the 8-GPU C2 list run one after the other on ONE GPU, with the scoring
kernel's time and (TM_SCORE_STATS=1) its cull statistics per shard: what the rank skew of an 8-GPU step is made of."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
from triplet_match_b200 import capi, workloads as wl

world = int(os.environ.get("WORLD", 8))
ctx = capi.Context(0)
model, scene = wl.c2_clouds()
hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **wl.DP, min_df=0.2, max_df=1.0, cap=200)
gm = hm.upload(ctx)
gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
rec = wl.c2_record(scene, hm.diameter, world)
for rank in range(world):
    q = capi.Query(gs, gm, **wl.QP, hyp_limit=wl.HYP_PER_GPU * world, max_hypotheses=int(wl.HYP_PER_GPU * 1.25))
    q.set_shard(rank, world)
    q.set_balance(True)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    ms = []
    for it in range(4):
        ctx.flush_l2(); q.run(); r = q.result(); ms.append(q.score_kernel_ms())
    print(f"rank {rank}: kernel {np.mean(ms[1:]):.2f} ms, hyps {r.n_scored}, tests {r.n_tests:.4e}", flush=True)
    q.close()
