#!/usr/bin/env python
"""The non-headline BASELINE.json configurations, measured the way bench.py measures C2
(CUDA events on the context stream, L2 flushed between steps, max over ranks).  One JSON line
per run on rank 0.

    python tools/bench_configs.py --config C3 [--steps K]     # free-form 50k model vs 10M scene
    python tools/bench_configs.py --config C4                 # 16 models x 5M scene, batched argmax
    python tools/bench_configs.py --config C5                 # ICP of 64 poses on a 10M scene
  multi-GPU: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port P tools/bench_configs.py --config C4
C3 / C4 shard hypotheses (weak scaling: 2^20 resp. 16 x 2^18 per GPU); C5 shards the scene
(strong scaling: 64 poses x 10M points in total)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DP = dict(distance_step_count=20.0, angle_step=0.17453292)
QP = dict(min_df=0.2, max_df=1.0, query_limit=200, dist_thres=1.0, accept_prob=0.5)


def model_of(kind, seed, synth):
    if kind == 0:
        return synth.plane_model(seed=seed, size=1.0, res=0.01, n_curves=10)
    if kind == 1:
        return synth.cylinder_model(seed=seed, radius=0.25, height=1.0, res=0.01, n_curves=4)
    return synth.freeform_model(seed=seed, n_points=12000, radius=0.01 * np.sqrt(12000 / (4 * np.pi)), n_bumps=8,
                                n_curves=6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["C3", "C4", "C5"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="scene size scale (dev)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import __graft_entry__ as ge
    ge.build()
    import torch
    from triplet_match_b200 import capi, synth
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    comm = None
    if world > 1:
        ids = [capi.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(ctx, ids[0], rank, world)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t0 = time.time()
    out = {"config": args.config, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "data": "synthetic",
           "dtype": "f32"}
    if args.config in ("C3", "C5"):
        n_scene = int(10_000_000 * args.scale)
        n_model = 50_000
        model = synth.freeform_model(seed=3, n_points=n_model, radius=0.01 * np.sqrt(n_model / (4 * np.pi)), n_bumps=12,
                                     n_curves=8)
        scene = synth.make_scene(seed=3, model=model, n_points=n_scene, n_copies=8, extent=10.0 * np.sqrt(n_scene / 1e6),
                                 flat_copies=False)
        poses = scene.poses
        scene = scene.take(synth.morton_order(scene.pos))
        hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **DP, min_df=QP["min_df"],
                            max_df=QP["max_df"], cap=QP["query_limit"])
        gm = hm.upload(ctx)
        gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
    if args.config == "C3":
        hyp_per_gpu = 1 << 20
        rec = synth.record_pairs(3, scene, hm.diameter, n_outer=256 * world, pairs_per_outer=128)
        q = capi.Query(gs, gm, **QP, hyp_limit=hyp_per_gpu * world, max_hypotheses=hyp_per_gpu)
        q.set_shard(rank, world)
        q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)

        def step():
            q.run()
            if comm is not None:
                comm.allreduce_best(q)
        queries = [q]
        out["workload"] = (f"C3: free-form {n_model}-point model (identity_traits case) vs {n_scene}-point scene, 2^20 "
                           f"hypotheses per GPU, grid {hm.extents.tolist()} = {int(np.prod(hm.extents.astype(np.int64)))} cells")
    elif args.config == "C4":
        n_scene = int(5_000_000 * args.scale)
        models = [model_of(k % 3, 40 + k, synth) for k in range(16)]
        scene = synth.make_scene(seed=4, model=models[0], n_points=n_scene, n_copies=16, extent=10.0 * np.sqrt(n_scene / 1e6),
                                 flat_copies=True, models=models)
        scene = scene.take(synth.morton_order(scene.pos))
        gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
        hyp_per_model = 1 << 18
        queries, hms, gms = [], [], []
        for k, mdl in enumerate(models):
            hmk = capi.HostModel(ctx, mdl.pos, mdl.nrm, mdl.tgt, curv_ok=mdl.tangent_mask, **DP, min_df=QP["min_df"],
                                 max_df=QP["max_df"], cap=QP["query_limit"])
            gmk = hmk.upload(ctx)
            rec = synth.record_pairs(100 + k, scene, hmk.diameter, n_outer=96 * world, pairs_per_outer=96)
            qk = capi.Query(gs, gmk, **QP, hyp_limit=hyp_per_model * world, max_hypotheses=hyp_per_model)
            qk.set_shard(rank, world)
            qk.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
            queries.append(qk); hms.append(hmk); gms.append(gmk)

        def step():
            for qk in queries:
                qk.run()
            if comm is not None:
                comm.allreduce_best_many(queries)
        out["workload"] = (f"C4: batched search, 16 models (plane / cylinder / free-form, {[m.n for m in models[:3]]}... points) x one "
                           f"{n_scene}-point scene, 2^18 hypotheses per model per GPU, one batched best-pose all-reduce")
    if args.config in ("C3", "C4"):
        for _ in range(args.warmup):
            step()
        barrier()
        ms = []
        for _ in range(args.steps):
            ctx.flush_l2()
            barrier()
            ctx.timer_start()
            step()
            ms.append(ctx.timer_stop())
        sec = maxr(float(np.sum(ms)) * 1e-3)
        rs = [qk.result() for qk in queries]
        scored = sumr(float(sum(r.n_scored for r in rs)))
        tests = sumr(float(sum(r.n_tests for r in rs)))
        out.update({"metric": "pose hypotheses scored/sec", "value": scored * args.steps / sec, "unit": "hypotheses/s",
                    "ms_per_step": sec * 1e3 / args.steps, "tests_per_sec": tests * args.steps / sec, "scaling": "weak",
                    "hypotheses_per_step": scored, "tests_per_step": tests,
                    "best_inliers": [int(r.best_inliers) for r in rs], "best_hypothesis": [int(r.best_hypothesis) for r in rs]})
    else:  # C5
        n_top, iters = 64, 5
        # 64 start poses: the scene's ground-truth model->scene poses inverted (scene->model), perturbed by <= 2 deg / <= 2 r
        rng = np.random.default_rng(5)
        Ts = np.zeros((n_top, 16), np.float32)
        for k in range(n_top):
            P = np.linalg.inv(poses[k % len(poses)])
            ax = rng.standard_normal(3); ax /= np.linalg.norm(ax)
            ang = np.deg2rad(2.0) * rng.random()
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            dR = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
            D = np.eye(4); D[:3, :3] = dR; D[:3, 3] = 0.02 * (rng.random(3) - 0.5)
            Ts[k] = (D @ P).T.reshape(-1).astype(np.float32)  # column-major
        b, e = capi.point_range(scene.n, rank, world)

        def step():
            return gs.icp_sharded(gm, Ts, iters, 1.0, b, e, scene.n, comm=comm)
        for _ in range(args.warmup):
            res = step()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            res = step()
        barrier()
        sec = maxr(time.perf_counter() - t1)
        To, co, so, io = res
        passes = float((io.astype(np.int64) + 1).clip(max=iters + 1).sum())  # accumulate passes per pose (>= 1)
        out.update({"metric": "ICP refinements/sec (top-64 poses, whole scene per pass)", "value": n_top * args.steps / sec,
                    "unit": "poses/s", "ms_per_step": sec * 1e3 / args.steps, "scaling": "strong",
                    "point_tests_per_sec": (iters + 1) * n_top * scene.n * args.steps / sec,
                    "timing": "host wall clock around tm_icp_sharded incl. the 4 KB pose H2D and result D2H",
                    "workload": f"C5: ICP (max {iters} iterations, 2 x dist_thres) of {n_top} perturbed poses against a {scene.n}-point scene, "
                                f"scene sharded x{world}; one int64 all-reduce of 64 x 17 sums per iteration",
                    "counts_max": int(co.max()), "iterations": io.tolist()[:8]})
    out["setup_s"] = round(time.time() - t0 - out.get("ms_per_step", 0) * 1e-3 * (args.steps + args.warmup), 1)
    out["gpu_launches"] = int(ctx.kernel_launches())
    if rank == 0:
        print(json.dumps(out), flush=True)
    barrier()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
