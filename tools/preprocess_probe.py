"""Dev probe: tangent-mask pre-processing (k-NN + curvature) on a 10M-point scene."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from triplet_match_b200 import capi, synth
n_scene = int(os.environ.get("N_SCENE", 10_000_000))
m = synth.pyramid_model(seed=7, size=1.0, height=0.4, res=0.01)
s = synth.make_scene(seed=8, model=m, n_points=n_scene, n_copies=8, extent=10.0 * np.sqrt(n_scene / 1e6), flat_copies=False)
print("scene", s.n, "marked tangent", int(s.tangent_mask.sum()), flush=True)
ctx = capi.Context(0)
for order in ("morton", "shuffled"):
    sc = s.take(synth.morton_order(s.pos)) if order == "morton" else s.take(synth.shuffle_perm(1, 1, s.n))
    gs = capi.Scene(ctx, sc.pos, sc.nrm, sc.tgt, np.zeros(sc.n, np.uint8))
    for _ in range(2):
        ctx.sync(); t = time.perf_counter(); mask, cnt = gs.compute_tangent_mask(30, 0.2, apply=True); dt = time.perf_counter() - t
    cand = int((np.linalg.norm(sc.tgt, axis=1) > 0.7).sum())
    print(f"{order}: tangent mask of {sc.n} points ({cand} candidates, 30-NN each): {dt * 1e3:.1f} ms wall, {cnt} pass", flush=True)
    gs.close()
