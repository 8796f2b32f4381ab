"""world_size-2 test (gloo, CPU) of the multi-GPU plan: hypotheses sharded contiguously,
one MAX all-reduce of the packed (inliers, ~id) key picks the global best pose, and
per-shard results concatenate to the unsharded result.  Scoring here is the oracle's;
the CUDA sharding itself is covered by tests/test_gpu_parity.py::test_sharded_query."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from triplet_match_b200 import capi


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, s, om, osc, rec = common.config("cylinder_small")
    T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    hb, he = capi.shard_range(T.shape[0], rank, world)
    subs = [osc.ball_subset(int(o), om.diameter) for o in rec.outer]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in subs])
    idx = np.concatenate(subs)
    cnt, scr, _ = osc.score_batch(om, T[hb:he], rec.pair_outer[hp[hb:he]], off, idx)
    best = 0
    for l, c in enumerate(cnt):
        if c:
            best = max(best, capi.pack_key(int(c), hb + l))
    # int64 transport: keys < 2^63 because inlier counts < 2^31
    t = torch.tensor([best], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (hb, he, cnt.tolist()))
    if rank == 0:
        q.put((int(t.item()), gathered, T.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_argmax_and_partition():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    key, gathered, n_hyp = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    m, s, om, osc, rec = common.config("cylinder_small")
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    subs = [osc.ball_subset(int(o), om.diameter) for o in rec.outer]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in subs])
    full, _, _ = osc.score_batch(om, T, rec.pair_outer[hp], off, np.concatenate(subs), nthreads=4)
    # shards tile the list exactly, in order
    assert gathered[0][0] == 0 and gathered[-1][1] == n_hyp
    assert all(gathered[r][1] == gathered[r + 1][0] for r in range(world - 1))
    assert np.array_equal(np.concatenate([np.array(g[2], dtype=np.uint32) for g in gathered]), full)
    inl, gid = capi.unpack_key(key)
    assert inl == int(full.max()) and gid == int(np.argmax(full))  # lowest id on ties


def test_shard_range_properties():
    for H in (0, 1, 7, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            ranges = [capi.shard_range(H, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= max(1, (H + world - 1) // world)
    assert capi.shard_range(100, 1, 2, hyp_limit=50) == (25, 50)
    assert capi.unpack_key(capi.pack_key(17, 5)) == (17, 5)
    assert capi.pack_key(3, 9) > capi.pack_key(3, 10) > capi.pack_key(2, 0)


# ---- scene-sharded ICP (tm_icp_sharded): integer sums all-reduce exactly ---------------------
def _icp_sums(scene_pos, model_pos, corr_s, corr_m, centre, scale):
    """64-bit fixed-point n, sum s, sum m, sum s m^T of the pairs (the quantities icp_accumulate
    reduces), quantised per term as round(v * scale)."""
    s = scene_pos[corr_s].astype(np.float64) - centre
    m = model_pos[corr_m].astype(np.float64) - centre
    out = np.zeros(16, dtype=np.int64)
    out[0] = corr_s.size
    out[1:4] = np.rint(s * scale).astype(np.int64).sum(0)
    out[4:7] = np.rint(m * scale).astype(np.int64).sum(0)
    out[7:16] = np.rint((s[:, :, None] * m[:, None, :]).reshape(-1, 9) * scale).astype(np.int64).sum(0)
    return out


def _icp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, s, om, osc, rec = common.config("cylinder_small")
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T[:64], nthreads=2)
    Tb = T[int(np.argmax(cnt))]
    b, e = capi.point_range(s.n, rank, world)
    r = osc.project(om, np.arange(b, e, dtype=np.int32), Tb, dist_thres=2.0)  # this rank's scene shard
    centre = m.pos.mean(0).astype(np.float64)
    sums = _icp_sums(s.pos, m.pos, r["scene_corrs"], r["model_corrs"], centre, 2.0 ** 30)
    t = torch.from_numpy(sums.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((t.numpy().copy(), (b, e)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scene_sharded_icp_sums():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_icp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, rng0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    m, s, om, osc, rec = common.config("cylinder_small")
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    cnt, _, _ = osc.score_batch(om, T[:64], nthreads=2)
    Tb = T[int(np.argmax(cnt))]
    full = osc.project(om, np.arange(s.n, dtype=np.int32), Tb, dist_thres=2.0)
    exp = _icp_sums(s.pos, m.pos, full["scene_corrs"], full["model_corrs"], m.pos.mean(0).astype(np.float64), 2.0 ** 30)
    assert rng0 == (0, s.n // 2)
    assert np.array_equal(total, exp) and total[0] == full["count"] > 100  # exact: integer sums commute


def test_point_range_properties():
    for n in (0, 1, 9, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            r = [capi.point_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


# ---- test-balanced shards (tm_query_set_balance) and pose-sharded ICP (tm_icp_pose_sharded) ------------
def _balance_worker(rank, world, port, q):
    """What the device path does with a communicator: every rank measures the subset sizes of the outer samples of its
    count-based share only, one MAX all-reduce completes the table, every rank derives the same bounds."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, s, om, osc, rec = common.config("freeform_small")
    T, hp, *_ = osc.hypotheses(om, rec.pair_i, rec.pair_j)
    n_outer = rec.outer.size
    nh = np.bincount(rec.pair_outer[hp], minlength=n_outer).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(nh)])
    hb, he = capi.shard_range(T.shape[0], rank, world)
    sizes = np.zeros(n_outer, dtype=np.int64)
    for o in range(n_outer):  # outer samples with hypotheses inside this rank's count-based shard
        if nh[o] and starts[o] < he and starts[o + 1] > hb:
            sizes[o] = osc.ball_subset(int(rec.outer[o]), om.diameter).size
    t = torch.from_numpy(sizes.copy())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    bounds = capi.balanced_bounds(nh, t.numpy(), world)
    # pose-sharded refinement: each rank's slice of 7 poses, gathered in rank order
    pb, pe = capi.pose_range(7, rank, world)
    mine = [(k, k * k) for k in range(pb, pe)]
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    gathered = [None] * world
    dist.all_gather_object(gathered, (bounds, int(np.count_nonzero(sizes))))
    if rank == 0:
        q.put((gathered, t.numpy().copy(), nh, [x for part in allp for x in part]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_balanced_bounds_and_pose_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_balance_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, sizes, nh, poses = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    m, s, om, osc, rec = common.config("freeform_small")
    full = np.array([osc.ball_subset(int(o), om.diameter).size if nh[k] else 0 for k, o in enumerate(rec.outer)])
    assert np.array_equal(sizes, full)                       # the MAX all-reduce completed the table
    assert gathered[0][0] == gathered[1][0]                  # every rank derives the same bounds
    assert all(g[1] < np.count_nonzero(full) for g in gathered)  # ... after sizing only its own share
    b = gathered[0][0]
    H = int(nh.sum())
    assert b[0] == 0 and b[-1] == H and all(x <= y for x, y in zip(b, b[1:]))
    tests_of = np.repeat(full, nh)                           # cost of every hypothesis of the global list
    per_rank = [int(tests_of[b[r]:b[r + 1]].sum()) for r in range(world)]
    assert sum(per_rank) == int(tests_of.sum()) and max(per_rank) <= 1.01 * sum(per_rank) / world + int(full.max())
    assert poses == [(k, k * k) for k in range(7)]           # slices concatenate in rank order


def test_balanced_bounds_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for trial in range(20):
            n = int(rng.integers(1, 40))
            nh = rng.integers(0, 500, n)
            sz = rng.integers(0, 3000, n)
            b = capi.balanced_bounds(nh, sz, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == int(nh.sum())
            assert all(x <= y for x, y in zip(b, b[1:]))
            cost = np.repeat(sz, nh)
            if cost.sum():
                per = [int(cost[b[r]:b[r + 1]].sum()) for r in range(world)]
                assert max(per) <= cost.sum() / world + sz.max() + 1
    assert capi.balanced_bounds([4, 4], [10, 30], 2) == [0, 5, 8]  # 160 tests, half = 80 = 4 * 10 + 1.33 * 30 -> 4 + 1
    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            r = [capi.pose_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b_[0] for a, b_ in zip(r, r[1:]))
