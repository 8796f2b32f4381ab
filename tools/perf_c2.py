"""Dev perf probe at the C2 size (plane model, 1M-point scene, up to 2^20 hypotheses)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from oracle import pyoracle as po
from triplet_match_b200 import capi, synth

def main():
    n_scene = int(os.environ.get("N_SCENE", 1_000_000)); n_outer = int(os.environ.get("N_OUTER", 96))
    ppo = int(os.environ.get("PPO", 128)); hyp_limit = int(os.environ.get("HYP_LIMIT", 1 << 20))
    steps = int(os.environ.get("STEPS", 3))
    t = time.time()
    m = synth.plane_model(seed=2, size=1.0, res=0.01, n_curves=6)
    s = synth.make_scene(seed=2, model=m, n_points=n_scene, n_copies=8, extent=10.0)
    if os.environ.get("MORTON", "1") == "1":
        s = s.take(synth.morton_order(s.pos))
    print("gen", round(time.time() - t, 1), "model", m.n, "tangent", int(m.tangent_mask.sum()), "scene", s.n, "tangent", int(s.tangent_mask.sum()), flush=True)
    t = time.time()
    om = po.OModel(m, resolution=float(os.environ.get("RES", -1)))
    print("oracle model", round(time.time() - t, 1), "ext", om.extents, "entries", om.n_entries, "keys", om.n_keys, "diam", om.diameter, "res", om.resolution, flush=True)
    rec = synth.record_pairs(2, s, om.diameter, n_outer, ppo)
    print("pairs", rec.pair_j.size, flush=True)
    ctx = capi.Context(0)
    gm = common.upload_model(ctx, m, om); gs = common.upload_scene(ctx, s)
    q = capi.Query(gs, gm, hyp_limit=hyp_limit, max_hypotheses=hyp_limit)
    q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
    for it in range(steps):
        ctx.flush_l2(); ctx.timer_start(); q.run(); ms = ctx.timer_stop(); r = q.result()
        print(f"step {it}: {ms:.2f} ms, hyps {r.n_scored}/{r.n_hypotheses}, tests {r.n_tests:.3e}, "
              f"{r.n_tests / ms * 1e3:.3e} tests/s, {r.n_scored / ms * 1e3:.3e} hyps/s, best {r.best_inliers} @ {r.best_hypothesis}", flush=True)
    print("launches", ctx.kernel_launches())
    # CPU oracle sample for reference
    d = q.download()
    osc = po.OScene(s)
    nh = min(512, d["T"].shape[0])
    hp = d["hyp_pair"][:nh]; hyp_sub = rec.pair_outer[hp]
    boff, bidx = gs.ball_subsets(rec.outer, om.diameter)
    t = time.time(); co, so, _ = osc.score_batch(om, d["T"][:nh], hyp_sub, boff, bidx, nthreads=os.cpu_count()); dt = time.time() - t
    tests = int(sum(int(boff[g + 1] - boff[g]) for g in hyp_sub))
    print("cpu oracle", nh, "hyps", f"{tests / dt:.3e} tests/s on", os.cpu_count(), "threads; counts==", np.array_equal(co, d["counts"][:nh]))

if __name__ == "__main__":
    main()
