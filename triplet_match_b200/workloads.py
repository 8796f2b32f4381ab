"""The named configurations of BASELINE.json (C2 .. C5) as seed-fixed synthetic workloads.

Harness code (like synth.py): bench.py, tools/bench_configs.py and tests/test_parity_at_size.py build
the same clouds and recorded sample lists from here, so a number in the bench line and a parity
assertion at that size are statements about the same input.

  C2  plane model (10 201 points) in a 1 M-point scene, 2^20 hypotheses per GPU      (configs[1])
  C3  free-form 50 k-point model vs a 10 M-point scene, 2^20 hypotheses per GPU      (configs[2])
  C4  16 models (plane / cylinder / free-form) x one 5 M-point scene, 2^18 per model (configs[3])
  C5  ICP of 64 perturbed poses against the C3 scene                                  (configs[4])
"""
from __future__ import annotations

import numpy as np

from . import synth

DP = dict(distance_step_count=20.0, angle_step=0.17453292)
QP = dict(min_df=0.2, max_df=1.0, query_limit=200, dist_thres=1.0, accept_prob=0.5)
HYP_PER_GPU = 1 << 20
C4_MODELS = 16
C4_HYP_PER_MODEL = 1 << 18
C5_TOP = 64
C5_ITERS = 5


# ---------------------------------------------------------------- C2
def c2_clouds(scale: float = 1.0):
    n_scene = int(1_000_000 * scale)
    model = synth.plane_model(seed=2, size=1.0, res=0.01, n_curves=12)
    scene = synth.make_scene(seed=2, model=model, n_points=n_scene, n_copies=8, extent=10.0 * np.sqrt(scale))
    scene = scene.take(synth.morton_order(scene.pos))
    return model, scene


def c2_record(scene, diameter: float, n_gpus: int):
    # ~ 9 k hypotheses per outer sample on this workload; oversample, the query clips the global list to
    # exactly n_gpus * 2^20 hypotheses (hyp_limit)
    return synth.record_pairs(2, scene, diameter, n_outer=256 * n_gpus, pairs_per_outer=128)


# ---------------------------------------------------------------- C3 / C5
def c3_clouds(scale: float = 1.0):
    """(model, scene in Z-curve order, ground-truth model->scene poses)."""
    n_scene = int(10_000_000 * scale)
    n_model = 50_000
    model = synth.freeform_model(seed=3, n_points=n_model, radius=0.01 * np.sqrt(n_model / (4 * np.pi)), n_bumps=12,
                                 n_curves=8)
    scene = synth.make_scene(seed=3, model=model, n_points=n_scene, n_copies=8, extent=10.0 * np.sqrt(n_scene / 1e6),
                             flat_copies=False)
    poses = scene.poses
    scene = scene.take(synth.morton_order(scene.pos))
    return model, scene, poses


def c3_record(scene, diameter: float, n_gpus: int):
    return synth.record_pairs(3, scene, diameter, n_outer=256 * n_gpus, pairs_per_outer=128)


def c5_start_poses(poses, n_top: int = C5_TOP, seed: int = 5) -> np.ndarray:
    """n_top start poses (column-major 4x4, scene -> model): the ground-truth poses inverted and perturbed by
    <= 2 degrees / <= 2 resolutions."""
    rng = np.random.default_rng(seed)
    Ts = np.zeros((n_top, 16), np.float32)
    for k in range(n_top):
        P = np.linalg.inv(poses[k % len(poses)])
        ax = rng.standard_normal(3)
        ax /= np.linalg.norm(ax)
        ang = np.deg2rad(2.0) * rng.random()
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        dR = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        D = np.eye(4)
        D[:3, :3] = dR
        D[:3, 3] = 0.02 * (rng.random(3) - 0.5)
        Ts[k] = (D @ P).T.reshape(-1).astype(np.float32)
    return Ts


# ---------------------------------------------------------------- C4
def c4_model(k: int):
    kind, seed = k % 3, 40 + k
    if kind == 0:
        return synth.plane_model(seed=seed, size=1.0, res=0.01, n_curves=10)
    if kind == 1:
        return synth.cylinder_model(seed=seed, radius=0.25, height=1.0, res=0.01, n_curves=4)
    return synth.freeform_model(seed=seed, n_points=12000, radius=0.01 * np.sqrt(12000 / (4 * np.pi)), n_bumps=8,
                                n_curves=6)


def c4_clouds(scale: float = 1.0, n_models: int = C4_MODELS):
    n_scene = int(5_000_000 * scale)
    models = [c4_model(k) for k in range(n_models)]
    scene = synth.make_scene(seed=4, model=models[0], n_points=n_scene, n_copies=16, extent=10.0 * np.sqrt(n_scene / 1e6),
                             flat_copies=True, models=models)
    scene = scene.take(synth.morton_order(scene.pos))
    return models, scene


def c4_record(scene, k: int, diameter: float, n_gpus: int):
    return synth.record_pairs(100 + k, scene, diameter, n_outer=96 * n_gpus, pairs_per_outer=96)
