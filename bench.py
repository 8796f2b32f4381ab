#!/usr/bin/env python
"""bench.py — hypotheses scored / s of the triplet_match search path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)

Headline workload "C2" (BASELINE.json configs[1]): plane_traits-style model (1 m x 1 m plane patch,
10 201 points, 12 line-segment feature curves) in a 1 M-point synthetic scene, 2^20 pose hypotheses per
GPU generated from a recorded, seed-fixed sample list.  One step = one pass of the hot path over that
batch: radius subsets -> pair features + keys -> hash probe -> base_transform_ -> inlier scoring of every
hypothesis against its subset (finish_find semantics, i.e. no early drop) -> best-pose argmax (+ one NCCL
max all-reduce when N > 1).  Hypotheses are sharded across ranks (weak scaling: 2^20 per GPU, shards of
equal hypothesis-point tests), scene + model replicated.

The same line carries, at the same N: `strong` (ONE 2^20-hypothesis C2 query split N ways), `configs`
(C3, C4, C5 of BASELINE.json, timed the same way) and `per_rank` (kernel time / tests min..max over ranks).

Prints ONE JSON line (see DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pose hypotheses scored/sec"
UNIT = "hypotheses/s"
BYTES_PER_TEST = 36  # 16 B scene float4 + 4 B voxel cell + 16 B model float4
SM_COUNT = 148
L1_WAVEFRONT_BYTES = 128  # one L1 data-stage wavefront moves at most one 128-byte line


def log(msg: str) -> None:
    """progress on stderr (the JSON line is the only thing on stdout)"""
    sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}\n")
    sys.stderr.flush()


def host_threads() -> int:
    """Worker threads of both CPU legs: hardware_concurrency() - 1, as find_parallel does (scene.hpp:146)."""
    return max(1, (os.cpu_count() or 2) - 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_profile():
    """Per-launch counters of the scoring kernel on the C2 step, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------- CPU legs
def cpu_reference_run(model, scene, rec, n_hyp_sample: int, steps: int, warmup: int, threads: int):
    """The reference's CPU path (oracle port) on a bounded sample: pair features -> query ->
    base_transform_ -> project_ (early_out = false) with `threads` std::threads.
    Returns (hyps/s, tests/s, ms/step, sample description)."""
    from oracle import pyoracle as po
    from triplet_match_b200.workloads import DP, QP
    om = po.OModel(model, **DP, min_df=QP["min_df"], max_df=QP["max_df"])
    osc = po.OScene(scene)
    # bounded sample: the first pairs of the recorded list until n_hyp_sample hypotheses
    npairs = min(rec.pair_j.size, 64)
    while True:
        T, hp, mi, mj, va = osc.hypotheses(om, rec.pair_i[:npairs], rec.pair_j[:npairs],
                                           limit=QP["query_limit"])
        if T.shape[0] >= n_hyp_sample or npairs >= rec.pair_j.size:
            break
        npairs = min(rec.pair_j.size, npairs * 2)
    T, hp = T[:n_hyp_sample], hp[:n_hyp_sample]
    outers = np.unique(rec.pair_outer[hp])
    remap = {int(o): k for k, o in enumerate(outers)}
    subs = [osc.ball_subset(int(rec.outer[o]), om.diameter) for o in outers]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([s.size for s in subs])
    idx = np.concatenate(subs) if subs else np.zeros(0, np.int32)
    hyp_sub = np.array([remap[int(o)] for o in rec.pair_outer[hp]], dtype=np.uint32)
    tests = int(sum(int(off[g + 1] - off[g]) for g in hyp_sub))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        # the whole path for the sample, as the reference runs it per hypothesis
        osc.pair_features(om, rec.pair_i[:npairs], rec.pair_j[:npairs])
        osc.hypotheses(om, rec.pair_i[:npairs], rec.pair_j[:npairs], limit=QP["query_limit"])
        osc.score_batch(om, T, hyp_sub, off, idx, accept_prob=QP["accept_prob"],
                        dist_thres=QP["dist_thres"], early_out=False, nthreads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    sample = (f"first {T.shape[0]} hypotheses of the recorded C2 list ({npairs} pairs, "
              f"{len(subs)} outer samples, {tests:.3e} hypothesis-point tests per step), "
              f"project_ early_out=false, {threads} std::threads (hardware_concurrency - 1)")
    return T.shape[0] / sec, tests / sec, sec * 1e3, sample


REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libtm_ref.so")


def cpu_reference_run_ref(model, scene, rec, n_hyp_sample: int, steps: int, warmup: int, threads: int):
    """Same bounded sample through the REFERENCE's own code: oracle/_ref/libtm_ref.so =
    /root/reference's model / scene / feature / discretize sources compiled against header
    stand-ins (oracle/Makefile target `ref`; Eigen/PCL/range-v3 are absent from the image).
    Per step: feature -> valid -> model::query -> base_transform_ for the sample's pairs, then
    project_(early_out=false) per hypothesis over its recorded radius subset, `threads` threads.
    Returns None when the library is absent."""
    import ctypes as C
    if not os.path.exists(REF_LIB):
        return None
    from oracle import pyoracle as po
    from triplet_match_b200.workloads import DP, QP
    L = C.CDLL(REF_LIB)
    L.ref_model_create.restype = C.c_void_p
    L.ref_model_create.argtypes = [C.c_void_p] * 3 + [C.c_uint32] + [C.c_float] * 4
    L.ref_scene_create.restype = C.c_void_p
    L.ref_scene_create.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p]
    L.ref_hypotheses_batch.restype = C.c_uint64
    L.ref_hypotheses_batch.argtypes = [C.c_void_p] * 4 + [C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
    L.ref_project_batch.argtypes = [C.c_void_p] * 3 + [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint32, C.c_float, C.c_float,
                                                                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    mp, mn, mt = f32(model.pos), f32(model.nrm), f32(model.tgt)
    sp, sn, st = f32(scene.pos), f32(scene.nrm), f32(scene.tgt)
    t0 = time.perf_counter()
    rm = C.c_void_p(L.ref_model_create(p(mp), p(mn), p(mt), model.n, DP["distance_step_count"], DP["angle_step"],
                                       QP["min_df"], QP["max_df"]))
    tmask = np.ascontiguousarray(scene.tangent_mask, dtype=np.uint8)
    rs = C.c_void_p(L.ref_scene_create(p(sp), p(sn), p(st), scene.n, p(tmask), None))
    t_init = time.perf_counter() - t0
    # sample selection (which pairs pass the scene-side filters, radius subsets) via the oracle
    om = po.OModel(model, **DP, min_df=QP["min_df"], max_df=QP["max_df"])
    osc = po.OScene(scene)
    npairs = min(rec.pair_j.size, 64)
    while True:
        f, k, v = osc.pair_features(om, rec.pair_i[:npairs], rec.pair_j[:npairs])
        ok = np.flatnonzero(v)
        pi = np.ascontiguousarray(rec.pair_i[:npairs][ok], dtype=np.uint32)
        pj = np.ascontiguousarray(rec.pair_j[:npairs][ok], dtype=np.uint32)
        T = np.zeros((n_hyp_sample, 16), np.float32)
        hp = np.zeros(n_hyp_sample, np.uint32)
        n = int(L.ref_hypotheses_batch(rs, rm, p(pi), p(pj), pi.size, QP["query_limit"], n_hyp_sample, p(T), p(hp)))
        if n >= n_hyp_sample or npairs >= rec.pair_j.size:
            break
        npairs = min(rec.pair_j.size, npairs * 2)
    T, hp = T[:n], hp[:n]
    pair_outer = rec.pair_outer[:npairs][ok][hp]
    outers = np.unique(pair_outer)
    remap = {int(o): q for q, o in enumerate(outers)}
    subs = [osc.ball_subset(int(rec.outer[o]), om.diameter) for o in outers]
    off = np.zeros(len(subs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in subs])
    idx = np.ascontiguousarray(np.concatenate(subs), dtype=np.int32)
    hyp_sub = np.array([remap[int(o)] for o in pair_outer], dtype=np.uint32)
    tests = int(sum(int(off[g + 1] - off[g]) for g in hyp_sub))
    counts = np.zeros(n, np.uint32)
    scores = np.zeros(n, np.float64)
    T2 = np.zeros_like(T)
    hp2 = np.zeros_like(hp)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        L.ref_hypotheses_batch(rs, rm, p(pi), p(pj), pi.size, QP["query_limit"], n, p(T2), p(hp2))
        L.ref_project_batch(rs, rm, p(T2), n, p(hyp_sub), p(off), p(idx), len(subs), QP["accept_prob"],
                            QP["dist_thres"], 0, threads, p(counts), p(scores))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    # the reference's counts on this sample against the oracle's (cross-check of the arm; the GPU is held to
    # both at this size by tests/test_parity_at_size.py)
    co, _, _ = osc.score_batch(om, T, hyp_sub, off, idx, accept_prob=QP["accept_prob"], dist_thres=QP["dist_thres"],
                               early_out=False, nthreads=threads)
    n_equal = int((co == counts).sum())
    sec = float(np.mean(times))
    sample = (f"first {n} hypotheses of the recorded C2 list ({int(pi.size)} valid pairs, {len(subs)} outer "
              f"samples, {tests:.3e} hypothesis-point tests per step): reference feature/valid/query/"
              f"base_transform_/project_(early_out=false) compiled from /root/reference against header stand-ins, "
              f"{threads} std::threads (hardware_concurrency - 1); counts equal the oracle's: {n_equal} of {n}; "
              f"reference model::init took {t_init:.1f} s (untimed)")
    return n / sec, tests / sec, sec * 1e3, sample, n_equal, n


# ------------------------------------------------------------------------------------------- GPU legs
class Dist:
    """torch.distributed plumbing (NCCL): barrier and max / sum / gather over ranks."""

    def __init__(self, world, local_rank):
        self.on = world > 1
        self.torch = None
        if self.on:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        if self.on:
            self.dist.barrier()

    def _red(self, v, op):
        if not self.on:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._red(v, self.dist.ReduceOp.MAX) if self.on else float(v)

    def min(self, v):
        return self._red(v, self.dist.ReduceOp.MIN) if self.on else float(v)

    def sum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM) if self.on else float(v)

    def close(self):
        if self.on:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_steps(ctx, D, step, steps, warmup, queries=None):
    """W warm-ups, then K steps, each between CUDA events on the context stream with L2 flushed before it
    (outside the events).  Returns (per-rank total ms, [scoring-kernel ms per step], kernel launches inside
    the timed steps)."""
    for _ in range(warmup):
        step()
    ctx.sync()
    D.barrier()
    ms, kms = [], []
    l0 = ctx.kernel_launches()
    for _ in range(steps):
        ctx.flush_l2()
        ctx.timer_start()
        step()
        ms.append(ctx.timer_stop())
        if queries:
            kms.append(float(sum(q.score_kernel_ms() for q in queries)))
    ctx.sync()
    launches = ctx.kernel_launches() - l0 - steps  # minus the untimed L2-flush launches
    D.barrier()
    return float(np.sum(ms)), kms, int(launches)


def run_query_config(ctx, D, comm, gs, gm_list, recs, hyp_per_query, steps, warmup, world, rank, balance=True):
    """One resident query per model over the same scene, shards of equal tests; one (batched) best-pose
    all-reduce per step.  Returns a dict of whole-job numbers."""
    from triplet_match_b200 import capi
    from triplet_match_b200.workloads import QP
    queries = []
    for gm, rec in zip(gm_list, recs):
        q = capi.Query(gs, gm, **QP, hyp_limit=hyp_per_query * world,
                       max_hypotheses=int(hyp_per_query * (1.25 if world > 1 else 1.0)) + 4096)
        q.set_shard(rank, world)
        if balance and world > 1:
            q.set_balance(True, comm)
        q.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        queries.append(q)

    def step():
        for q in queries:
            q.run()
        if comm is not None:
            if len(queries) == 1:
                comm.allreduce_best(queries[0])
            else:
                comm.allreduce_best_many(queries)

    total_ms, kms, launches = timed_steps(ctx, D, step, steps, warmup, queries)
    rs = [q.result() for q in queries]
    scored = float(sum(int(r.n_scored) for r in rs))
    tests = float(sum(int(r.n_tests) for r in rs))
    sec = D.max(total_ms) * 1e-3
    k_ms = float(np.mean(kms)) if kms else 0.0
    out = {"value": D.sum(scored) * steps / sec, "unit": UNIT, "ms_per_step": sec * 1e3 / steps,
           "tests_per_sec": D.sum(tests) * steps / sec, "hypotheses_per_step": D.sum(scored),
           "tests_per_step": D.sum(tests),
           "per_rank": {"score_kernel_ms": [D.min(k_ms), D.max(k_ms)], "tests": [D.min(tests), D.max(tests)],
                        "hypotheses": [D.min(scored), D.max(scored)], "step_ms": [D.min(total_ms / steps), D.max(total_ms / steps)]},
           "best_inliers": [int(r.best_inliers) for r in rs][:4], "gpu_launches": launches}
    return out, queries, step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("TM_BENCH_SCALE", 1.0)),
                    help="scene size scale (dev only; 1.0 = the named configuration)")
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("TM_CPU_SAMPLE", 16384)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 legs (dev)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    n_gpus = max(args.gpus, world)
    threads = host_threads()

    from triplet_match_b200 import workloads as wl
    DP, QP, HYP_PER_GPU = wl.DP, wl.QP, wl.HYP_PER_GPU
    config = {"workload": "C2: plane model 10201 pts / 1M-point synthetic scene / 2^20 hypotheses "
                          "per GPU from a recorded seed-fixed sample list (BASELINE.json configs[1])",
              "scene_points": int(1_000_000 * args.scale), "hypotheses_per_gpu": HYP_PER_GPU,
              "scoring": "finish_find semantics (project_ early_out=false)",
              "parallelism": f"hypotheses sharded x{n_gpus} (shards of equal hypothesis-point tests), scene+model replicated",
              "l2": "flushed between timed steps (256 MiB write, outside the per-step events)"}

    if args.impl == "reference":
        if rank != 0:
            return
        # test infrastructure only: the oracle and the shim-compiled reference; this arm never loads the CUDA library
        from oracle import build as oracle_build
        oracle_build.build_oracle()
        model, scene = wl.c2_clouds(args.scale)
        # recorded list needs the model diameter only (no GPU): bbox diagonal in float32
        lo, hi = model.pos.min(axis=0), model.pos.max(axis=0)
        d = (hi - lo).astype(np.float32)
        diam = float(np.sqrt(np.float32(d[0] * d[0]) + (np.float32(d[1] * d[1]) + np.float32(d[2] * d[2]))))
        rec = wl.c2_record(scene, diam, n_gpus)
        port = cpu_reference_run(model, scene, rec, args.cpu_sample, max(1, min(args.steps, 3)), 1, threads)
        ref = cpu_reference_run_ref(model, scene, rec, args.cpu_sample, args.steps, args.warmup, threads)
        kind = "reference" if ref is not None else "port"
        hps, tps, ms, sample = (ref if ref is not None else port)[:4]
        line = {"impl": "reference", "metric": METRIC, "value": hps, "unit": UNIT, "n_gpus": n_gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "tests_per_sec": tps,
                "cpu_baseline": {"value": hps, "unit": UNIT, "cores": threads, "kind": kind,
                                 "sample": sample},
                "e2e": {"value": hps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "oracle_port": {"value": port[0], "unit": UNIT, "tests_per_sec": port[1], "ms_per_step": port[2],
                                "cores": threads},
                "note": "the reference's own build (cmake + PCL/FLANN/Eigen/range-v3/boost/fmt) is impossible here; "
                        "kind=reference runs its sources compiled against header stand-ins (oracle/_ref), "
                        "kind=port the dependency-free oracle restatement"}
        if ref is not None:
            line["counts_equal_oracle"] = {"equal": ref[4], "of": ref[5]}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ ours
    import __graft_entry__ as ge
    ge.build()
    import torch
    from triplet_match_b200 import capi

    D = Dist(world, local_rank)
    ctx = capi.Context(local_rank)
    comm = None
    if D.on:
        ids = [capi.Comm.unique_id() if rank == 0 else None]
        D.dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(ctx, ids[0], rank, world)

    model, scene = wl.c2_clouds(args.scale)
    hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **DP,
                        min_df=QP["min_df"], max_df=QP["max_df"], cap=QP["query_limit"])
    gm = hm.upload(ctx)
    gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
    rec = wl.c2_record(scene, hm.diameter, n_gpus)

    # ---- headline: weak scaling, 2^20 hypotheses per GPU ------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    wall0 = time.perf_counter()
    head, queries, step = run_query_config(ctx, D, comm, gs, [gm], [rec], HYP_PER_GPU, args.steps, args.warmup, world, rank)
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    q = queries[0]
    launches = head["gpu_launches"]
    r = q.result()
    n_scored, n_tests = int(r.n_scored), int(r.n_tests)
    front_ms = q.frontend_ms()

    log(f"headline done: {head['ms_per_step']:.2f} ms/step")
    # ---- e2e: host buffers in, host results out, through the C-ABI ----------
    # inputs (the recorded list) and outputs (result record + per-hypothesis counts) live in pinned host memory
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    p_outer, p_pair_outer, p_pair_j = pinned(rec.outer.astype(np.uint32)), pinned(rec.pair_outer.astype(np.uint32)), pinned(rec.pair_j.astype(np.uint32))
    p_counts = torch.empty(int(q.params.max_hypotheses), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    e2e_ms = []
    for it in range(2 + args.steps):
        D.barrier()
        t0 = time.perf_counter()
        q.set_pairs(p_outer, p_pair_outer, p_pair_j)         # H2D: recorded list; front end + sizing of this rank's shard
        q.run()
        if comm is not None:
            comm.allreduce_best(q)
        q.download_counts(out=p_counts)                      # D2H: result + per-hypothesis counts
        dt = (time.perf_counter() - t0) * 1e3
        if it >= 2:
            e2e_ms.append(dt)
    e2e_sec = D.max(float(np.mean(e2e_ms))) * 1e-3
    h2d = 4 * (rec.outer.size + 2 * rec.pair_j.size)
    d2h = 4 * n_scored + 152 + 8 * (rec.outer.size + 1)
    e2e = {"value": head["hypotheses_per_step"] / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_sec * 1e3,
           "call": "tm_query_set_pairs + tm_query_run + tm_query_result_get + tm_query_download "
                   "(scene + model resident, as in the reference where they are built before find)"}

    log(f"e2e done: {e2e_sec * 1e3:.2f} ms/step")
    # ---- roofline of the dominant kernel (score_count_x2_kernel) ------------
    # What binds it is on-chip: the L1 data stage (ncu l1tex__data_pipe_lsu_wavefronts 96 % of peak), fed by the
    # scattered 16-byte cell gathers.  achieved = data-stage wavefronts of one launch (from the committed ncu
    # capture of this exact workload: they depend on the data, not on the run) x 128 B / this run's kernel time;
    # peak = one wavefront per SM per cycle at the SM clock sampled during the timed region.
    prof = kernel_profile()
    peak_hbm, peak_src = measured_peak_gbs()
    k_ms = head["per_rank"]["score_kernel_ms"][1]
    sm_mhz = float(clocks.get("sm_mhz") or 1965.0)
    wavefronts = float(prof.get("l1_data_pipe_wavefronts_per_launch") or 0.0)
    l1_peak = SM_COUNT * L1_WAVEFRONT_BYTES * sm_mhz * 1e6 / 1e9
    l1_ach = wavefronts * L1_WAVEFRONT_BYTES / (k_ms * 1e-3) / 1e9 if (k_ms > 0 and args.scale == 1.0) else 0.0
    hbm_ach = n_tests * BYTES_PER_TEST / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    executed = prof.get("executed_tests_per_launch")
    roofline = {"bound": "l1tex", "achieved": l1_ach, "peak": l1_peak, "unit": "GB/s", "frac": l1_ach / l1_peak if l1_peak else None,
                "traffic": prof.get("dram_bytes_per_launch"),
                "kernel": prof.get("kernel", "score_count_x2_kernel"), "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms / head["ms_per_step"] if head["ms_per_step"] else None,
                "definition": "L1 data-stage wavefronts per launch (ncu l1tex__data_pipe_lsu_wavefronts.sum, committed "
                              "capture of this workload) x 128 B / live kernel time, against 148 SMs x 1 wavefront/clk x "
                              "the SM clock sampled in this run; ncu read 96 % for the same quantity",
                "ncu": {k: prof.get(k) for k in ("l1tex_data_pipe_pct", "lsu_writeback_pct", "issue_active_pct", "lts_throughput_pct",
                                                  "dram_throughput_pct", "l2_hit_pct", "warp_instructions_per_launch", "source")},
                "nominal_tests_per_step": n_tests, "executed_tests_per_step": executed,
                "executed_tests_note": "tests that survive the exact box cull (TM_SCORE_STATS); the rest are proven misses",
                "warp_instructions_per_32_executed_tests": (prof["warp_instructions_per_launch"] * 32.0 / executed)
                if executed and prof.get("warp_instructions_per_launch") else None,
                "hbm_form": {"achieved": hbm_ach, "peak": peak_hbm, "unit": "GB/s", "frac": hbm_ach / peak_hbm, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": n_tests * BYTES_PER_TEST,
                             "note": "SURVEY §8d's 36 B per nominal test against measured HBM copy bandwidth: far above 1 because "
                                     "78 % of the nominal tests are culled and the rest is served from registers and L2 "
                                     "(DRAM moves ~0.1 GB per launch) — kept for the contract, it is not what bounds the kernel"}}

    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "tests_per_sec": head["tests_per_sec"],
            "hypotheses_per_step": head["hypotheses_per_step"], "tests_per_step": head["tests_per_step"],
            "best_inliers": int(r.best_inliers), "best_hypothesis": int(r.best_hypothesis),
            "wall_ms_per_step_incl_flush": wall * 1e3 / (args.steps + args.warmup),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "per_rank": head["per_rank"],
            "front_end_ms": {"value": D.max(front_ms), "note": "replicated on every rank: pair features + probe + scan over the "
                             f"whole {rec.pair_j.size}-pair list, shard and per-outer ranges (CUDA events, last step)"}}

    # ---- strong scaling: ONE 2^20-hypothesis C2 query split N ways ----------------------------------
    if world > 1:
        try:
            rec1 = wl.c2_record(scene, hm.diameter, 1)  # the 1-GPU headline's own list
            sq, squeries, _ = run_query_config(ctx, D, comm, gs, [gm], [rec1], HYP_PER_GPU // world, args.steps, args.warmup,
                                               world, rank)
            line["strong"] = {"value": sq["value"], "unit": UNIT, "ms_per_step": sq["ms_per_step"],
                              "hypotheses_per_step": sq["hypotheses_per_step"], "per_rank": sq["per_rank"],
                              "workload": f"one 2^20-hypothesis C2 query split over {world} GPUs"}
            for x in squeries:
                x.close()
        except Exception as e:  # noqa: BLE001
            log("strong leg failed: " + repr(e))
            line["strong"] = {"error": repr(e)}
        log("strong done")
    else:
        line["strong"] = {"value": head["value"], "unit": UNIT, "ms_per_step": head["ms_per_step"],
                          "hypotheses_per_step": head["hypotheses_per_step"],
                          "workload": "one 2^20-hypothesis C2 query on 1 GPU (= the headline)"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        hps, tps, ms, sample = cpu_reference_run(model, scene, rec, args.cpu_sample, 3, 1, threads)
        line["cpu_baseline"] = {"value": hps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": sample, "tests_per_sec": tps, "ms_per_step": ms}
        # p50 full-query latency incl. ICP of the top 64 (SURVEY §8d metric ii)
        q2 = capi.Query(gs, gm, **QP, hyp_limit=HYP_PER_GPU, max_hypotheses=HYP_PER_GPU,
                        icp_top_k=64, max_icp_iterations=5)
        q2.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        lat = []
        for it in range(3 + 20):
            t0 = time.perf_counter()
            q2.run()
            q2.result()
            if it >= 3:
                lat.append((time.perf_counter() - t0) * 1e3)
        line["p50_query_ms"] = float(np.median(lat))
        line["p50_query_def"] = ("resident scene+model; subsets->features->probe->hypotheses(2^20)->"
                                 "score->argmax->ICP(top 64, 5 iterations); 20 repeats after 3 warm-ups")
        q2.close()
        # the reference's own operating mode, project_(early_out = true) (scene.hpp:326, 492-506), beside the headline
        try:
            line["early_drop_modes"] = early_drop_legs(ctx, capi, gs, gm, rec, r, QP, HYP_PER_GPU)
        except Exception as e:  # noqa: BLE001
            line["early_drop_modes"] = {"error": repr(e)}
    q.close()
    gm.close(); gs.close(); hm.close()

    # ---- the other named configurations at the same N -------------------------------------------------
    if not args.no_configs and args.scale == 1.0:
        cfgs = {}
        for name, fn in (("C3", bench_c3_c5), ("C4", bench_c4)):
            try:
                cfgs.update(fn(ctx, D, comm, capi, wl, args.steps, args.warmup, world, rank))
                log(f"{name} leg done")
            except Exception as e:  # noqa: BLE001
                log(f"{name} leg failed: " + repr(e))
                cfgs[name] = {"error": repr(e)}
        line["configs"] = cfgs
    if comm is not None:
        comm.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    D.close()


def early_drop_legs(ctx, capi, gs, gm, rec, r_full, QP, hyp):
    out = {}
    for mode, key in ((1, "subset_order"), (2, "even_walk")):
        qe = capi.Query(gs, gm, **QP, early_out=mode, hyp_limit=hyp, max_hypotheses=hyp)
        qe.set_pairs(rec.outer, rec.pair_outer, rec.pair_j)
        ms = []
        for it in range(2 + 3):
            ctx.flush_l2()
            ctx.timer_start()
            qe.run()
            t = ctx.timer_stop()
            if it >= 2:
                ms.append(t)
        re_ = qe.result()
        d = qe.download()
        alive = d["dropped"] == 0
        out[key] = {"value": re_.n_scored / (float(np.mean(ms)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ms)),
                    "scoring_ms": float(qe.score_kernel_ms()), "tests_per_step": int(re_.n_tests), "survivors": int(alive.sum()),
                    "best_pose_survives": bool(alive.any() and int(d["counts"][alive].max()) == int(r_full.best_inliers))}
        if mode == 2:
            out[key]["kernel"] = "score_level_kernel: 19 checkpoint ranges in 6 launches + el_eval_kernel (k_early2.cu)"
            out[key]["walked_one_by_one"] = qe.early_walked()
        qe.close()
    out["note"] = ("project_(early_out=true) semantics, bit-exact with the reference incl. drop points: in the subset's own order "
                   "(early_out=1, one warp walks one hypothesis) and over the evenly sampling walk p -> (p*s) mod n (early_out=2, "
                   "evaluated per checkpoint range with the tiled, box-culled scorer); not the headline, which scores every "
                   "hypothesis over its whole subset")
    return out


def bench_c3_c5(ctx, D, comm, capi, wl, steps, warmup, world, rank):
    """C3 (free-form 50 k model vs 10 M scene, 2^20 hypotheses per GPU) and C5 (ICP of 64 poses on that scene,
    poses sharded over the ranks: strong scaling)."""
    model, scene, poses = wl.c3_clouds()
    hm = capi.HostModel(ctx, model.pos, model.nrm, model.tgt, curv_ok=model.tangent_mask, **wl.DP,
                        min_df=wl.QP["min_df"], max_df=wl.QP["max_df"], cap=wl.QP["query_limit"])
    gm = hm.upload(ctx)
    gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
    rec = wl.c3_record(scene, hm.diameter, world)
    c3, queries, _ = run_query_config(ctx, D, comm, gs, [gm], [rec], wl.HYP_PER_GPU, max(2, steps // 2), 3, world, rank)
    c3["scaling"] = "weak"
    c3["workload"] = (f"free-form {model.n}-point model vs {scene.n}-point scene, 2^20 hypotheses per GPU, grid "
                      f"{hm.extents.tolist()}")
    if world == 1:  # the reference's operating mode on this workload, beside full scoring (as for C2)
        try:
            c3["early_drop_modes"] = early_drop_legs(ctx, capi, gs, gm, rec, queries[0].result(), wl.QP, wl.HYP_PER_GPU)
        except Exception as e:  # noqa: BLE001 - a failing side leg must not take the line with it
            c3["early_drop_modes"] = {"error": repr(e)}
    for q in queries:
        q.close()
    # C5
    Ts = wl.c5_start_poses(poses)
    n_top, iters = Ts.shape[0], wl.C5_ITERS

    def icp_step():
        return gs.icp_pose_sharded(gm, Ts, iters, wl.QP["dist_thres"], rank=rank, world=world, comm=comm)
    for _ in range(3):
        res = icp_step()
    ctx.sync()
    D.barrier()
    reps = max(5, steps)
    t1 = time.perf_counter()
    for _ in range(reps):
        res = icp_step()
    ctx.sync()
    sec = D.max(time.perf_counter() - t1)
    To, co, so, io = res
    c5 = {"value": n_top * reps / sec, "unit": "poses/s", "ms_per_step": sec * 1e3 / reps, "scaling": "strong",
          "metric": "ICP refinements/sec (64 poses, whole scene per pass)",
          "timing": "host wall clock around tm_icp_pose_sharded incl. pose H2D, result D2H and the all-gather; max over ranks",
          "workload": f"ICP (max {iters} iterations, 2 x dist_thres) of {n_top} perturbed poses against a {scene.n}-point scene; "
                      f"poses sharded x{world}, whole scene on every GPU, one all-gather of 80-byte records at the end",
          "counts_min_max": [int(co.min()), int(co.max())]}
    gm.close(); gs.close(); hm.close()
    return {"C3": c3, "C5": c5}


def bench_c4(ctx, D, comm, capi, wl, steps, warmup, world, rank):
    """C4: 16 models x one 5 M-point scene, 2^18 hypotheses per model per GPU, one batched best-pose all-reduce."""
    models, scene = wl.c4_clouds()
    gs = capi.Scene(ctx, scene.pos, scene.nrm, scene.tgt, scene.tangent_mask)
    hms, gms, recs = [], [], []
    for k, mdl in enumerate(models):
        hmk = capi.HostModel(ctx, mdl.pos, mdl.nrm, mdl.tgt, curv_ok=mdl.tangent_mask, **wl.DP, min_df=wl.QP["min_df"],
                             max_df=wl.QP["max_df"], cap=wl.QP["query_limit"])
        hms.append(hmk)
        gms.append(hmk.upload(ctx))
        recs.append(wl.c4_record(scene, k, hmk.diameter, world))
    c4, queries, _ = run_query_config(ctx, D, comm, gs, gms, recs, wl.C4_HYP_PER_MODEL, max(2, steps // 2), 3, world, rank)
    c4["scaling"] = "weak"
    c4["workload"] = (f"16 models (plane / cylinder / free-form) x one {scene.n}-point scene, 2^18 hypotheses per model per GPU, "
                      "one batched best-pose all-reduce")
    for q in queries:
        q.close()
    for g in gms:
        g.close()
    for h in hms:
        h.close()
    gs.close()
    return {"C4": c4}


if __name__ == "__main__":
    main()
