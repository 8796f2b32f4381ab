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
