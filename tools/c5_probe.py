"""C5 probe (dev): wall time of tm_icp for 64 poses and for the 8-pose share of an 8-GPU pose-sharded run,
on the 10 M-point C3 scene.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os, sys, time
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
Ts = wl.c5_start_poses(poses)
reps = int(os.environ.get("REPS", 20))
for n in (64, 8):
    for _ in range(3):
        gs.icp(gm, Ts[:n], wl.C5_ITERS, 1.0)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = gs.icp(gm, Ts[:n], wl.C5_ITERS, 1.0)
    ctx.sync()
    print(f"poses {n}: {(time.perf_counter() - t0) * 1e3 / reps:.3f} ms per refinement, iters {r[3][:8].tolist()}, counts {r[1][:4].tolist()}", flush=True)
