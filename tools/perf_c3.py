"""Dev probe at the C3 size: free-form 50k-point model vs 10M-point scene (BASELINE configs[2])."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from triplet_match_b200 import capi, synth

def main():
    n_scene = int(os.environ.get("N_SCENE", 10_000_000)); n_model = int(os.environ.get("N_MODEL", 50_000))
    n_outer = int(os.environ.get("N_OUTER", 256)); ppo = int(os.environ.get("PPO", 128))
    hyp_limit = int(os.environ.get("HYP_LIMIT", 1 << 20)); steps = int(os.environ.get("STEPS", 3))
    t = time.time()
    radius = 0.01 * np.sqrt(n_model / (4 * np.pi))
    m = synth.freeform_model(seed=3, n_points=n_model, radius=radius, n_bumps=12, n_curves=8)
    ext = 10.0 * np.sqrt(n_scene / 1e6)
    s = synth.make_scene(seed=3, model=m, n_points=n_scene, n_copies=8, extent=ext, flat_copies=False)
    print("gen", round(time.time() - t, 1), "s; model", m.n, "tangent", int(m.tangent_mask.sum()), "scene", s.n,
          "tangent", int(s.tangent_mask.sum()), flush=True)
    t = time.time(); s = s.take(synth.morton_order(s.pos)); print("morton", round(time.time() - t, 1), flush=True)
    ctx = capi.Context(0)
    t = time.time()
    hm = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, min_df=0.2, max_df=1.0, cap=200)
    print("model::init", round(time.time() - t, 1), "s; extents", hm.extents, "cells", int(np.prod(hm.extents.astype(np.int64))),
          "entries", hm.n_entries, "keys", hm.n_keys, "kept", hm.n_kept, "diam", hm.diameter, "res", hm.resolution, flush=True)
    t = time.time(); gm = hm.upload(ctx); gs = capi.Scene(ctx, s.pos, s.nrm, s.tgt, s.tangent_mask)
    print("upload", round(time.time() - t, 1), flush=True)
    t = time.time(); rec = synth.record_pairs(3, s, hm.diameter, n_outer, ppo); print("record", round(time.time() - t, 1), "pairs", rec.pair_j.size, flush=True)
    q = capi.Query(gs, gm, hyp_limit=hyp_limit, max_hypotheses=hyp_limit)
    t = time.time(); q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j); ctx.sync(); print("set_pairs", round(time.time() - t, 2), flush=True)
    for it in range(steps):
        ctx.flush_l2(); ctx.timer_start(); q.run(); ms = ctx.timer_stop(); r = q.result()
        print(f"step {it}: {ms:.2f} ms (score kernel {q.score_kernel_ms():.2f}), hyps {r.n_scored}/{r.n_hypotheses}, tests {r.n_tests:.3e}, "
              f"{r.n_tests / ms * 1e3:.3e} tests/s, {r.n_scored / ms * 1e3:.3e} hyps/s, best {r.best_inliers} @ {r.best_hypothesis}", flush=True)
    c = q.download_counts()[0]
    print("inliers: sum", int(c.astype(np.int64).sum()), "mean", float(c.mean()), "max", int(c.max()))
    # ICP of the top 64 on the whole scene (C5)
    d_top = np.argsort(-c.astype(np.int64), kind="stable")[:64]
    dl = q.download()
    for iters in (1, 5):
        ctx.sync(); t = time.time(); To, co, so, io = gs.icp(gm, dl["T"][d_top], iters, 1.0); dt = time.time() - t
        print(f"icp top-64 x {iters} iters on {s.n} points: {dt * 1e3:.1f} ms wall (incl. H2D/D2H), counts max {co.max()} iters {io.tolist()[:8]}", flush=True)

if __name__ == "__main__":
    main()
