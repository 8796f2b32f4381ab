"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
       --master-port 29511 tools/mgpu_check.py
1. hypothesis-sharded query + tm_query_allreduce_best == the unsharded query's best pose;
2. scene-sharded ICP (tm_icp_sharded over NCCL) == tm_icp on the whole scene, bit for bit,
   and identical on every rank;
3. pose-sharded ICP (tm_icp_pose_sharded + one ncclAllGather) == tm_icp on every rank;
4. test-balanced shards (tm_query_set_balance with the communicator's max all-reduce of the subset
   sizes): the shards still partition the list (counts gathered in rank order == the unsharded
   counts) and the best pose is the unsharded one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import common
from triplet_match_b200 import capi


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    ids = [capi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = capi.Comm(ctx, ids[0], rank, world)
    ok = True
    for name in ("cylinder_small", "freeform_small"):
        m, s, om, osc, rec = common.config(name)
        gm = common.upload_model(ctx, m, om)
        gs = common.upload_scene(ctx, s)
        # 1. sharded search
        q = capi.Query(gs, gm)
        q.set_shard(rank, world)
        q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        q.run()
        comm.allreduce_best(q)
        r = q.result()
        q1 = capi.Query(gs, gm)
        q1.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        q1.run()
        r1 = q1.result()
        same = (r.best_inliers == r1.best_inliers and r.best_hypothesis == r1.best_hypothesis and
                np.array_equal(np.array(list(r.best_T), np.float32).view(np.uint32),
                               np.array(list(r1.best_T), np.float32).view(np.uint32)))
        # 2. scene-sharded ICP of the top hypotheses
        d = q1.download()
        top = np.argsort(-d["counts"].astype(np.int64), kind="stable")[:8]
        To, co, so, io = gs.icp(gm, d["T"][top], 5, 1.0)
        b, e = capi.point_range(s.n, rank, world)
        Ts, cs, ss, is_ = gs.icp_sharded(gm, d["T"][top], 5, 1.0, b, e, s.n, comm=comm)
        same_icp = (np.array_equal(Ts.view(np.uint32), To.view(np.uint32)) and np.array_equal(cs, co) and
                    np.array_equal(ss, so) and np.array_equal(is_, io))
        # 3. pose-sharded ICP
        for _ in range(4):  # repeated calls replay the cached graph between the gathers
            Tp, cp, sp_, ip = gs.icp_pose_sharded(gm, d["T"][top], 5, 1.0, comm=comm)
        same_pose = (np.array_equal(Tp.view(np.uint32), To.view(np.uint32)) and np.array_equal(cp, co) and
                     np.array_equal(sp_, so) and np.array_equal(ip, io))
        # 4. test-balanced shards
        qb = capi.Query(gs, gm)
        qb.set_shard(rank, world)
        qb.set_balance(True, comm)
        qb.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        qb.run()
        comm.allreduce_best(qb)
        rb = qb.result()
        mine = qb.download_counts()[0]
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        same_bal = (np.array_equal(np.concatenate(parts), d["counts"]) and rb.best_inliers == r1.best_inliers and
                    rb.best_hypothesis == r1.best_hypothesis)
        tests = [None] * world
        dist.all_gather_object(tests, int(rb.n_tests))
        flag = torch.tensor([int(same), int(same_icp), int(same_pose), int(same_bal)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{name}: world {world}: sharded best == unsharded: {bool(flag[0])} (inliers {r.best_inliers} @ "
                  f"{r.best_hypothesis}); scene-sharded ICP == tm_icp on all ranks: {bool(flag[1])} (counts {cs.tolist()}); "
                  f"pose-sharded ICP == tm_icp on all ranks: {bool(flag[2])}; test-balanced shards partition the list and find "
                  f"the same best: {bool(flag[3])} (tests per rank {tests})", flush=True)
        ok = ok and all(bool(x) for x in flag)
        q.close(); q1.close(); qb.close(); gm.close(); gs.close()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
