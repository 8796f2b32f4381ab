#!/usr/bin/env python
"""The non-headline BASELINE.json configurations on their own (bench.py carries the same legs in its line):

    python tools/bench_configs.py --config C3 [--steps K]     # free-form 50k model vs 10M scene
    python tools/bench_configs.py --config C4                 # 16 models x 5M scene, batched argmax
    python tools/bench_configs.py --config C5                 # ICP of 64 poses on a 10M scene, poses sharded
  multi-GPU: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port P tools/bench_configs.py --config C4

Timing and workloads are bench.py's (bench_c3_c5 / bench_c4: CUDA events on the context stream, L2 flushed
between steps, max over ranks; triplet_match_b200/workloads.py).  Unless --no-check is given, rank 0 then runs
the at-size parity test of that configuration (tests/test_parity_at_size.py: >= 2 048 hypotheses spread over all
outer samples against the CPU oracle, whole subsets; ICP counts and poses for C5) and the tool FAILS if it does."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["C3", "C4", "C5"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import __graft_entry__ as ge
    ge.build()
    import bench
    from triplet_match_b200 import capi, workloads as wl
    D = bench.Dist(world, local)
    ctx = capi.Context(local)
    comm = None
    if D.on:
        ids = [capi.Comm.unique_id() if rank == 0 else None]
        D.dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(ctx, ids[0], rank, world)
    fn = bench.bench_c4 if args.config == "C4" else bench.bench_c3_c5
    res = fn(ctx, D, comm, capi, wl, args.steps, args.warmup, world, rank)
    out = {"config": args.config, "n_gpus": world, "steps": args.steps, "data": "synthetic", "dtype": "f32", **res[args.config]}
    if comm is not None:
        comm.close()
    ctx.close()
    rc = 0
    if rank == 0:
        if not args.no_check:
            key = {"C3": "c3_counts", "C4": "c4_counts", "C5": "c5_icp"}[args.config]
            env = dict(os.environ)
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
                env.pop(k, None)
            p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_at_size.py"), "-m", "gpu",
                                "-x", "-q", "-k", key], capture_output=True, text=True, env=env)
            out["parity_check"] = {"test": f"tests/test_parity_at_size.py -k {key}", "passed": p.returncode == 0,
                                   "tail": p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""}
            rc = p.returncode
            if rc:
                sys.stderr.write(p.stdout[-4000:] + p.stderr[-2000:])
        print(json.dumps(out), flush=True)
    D.close()
    if rc:
        raise SystemExit(f"parity check of {args.config} FAILED")


if __name__ == "__main__":
    main()
