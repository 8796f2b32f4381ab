"""Shared test helpers: small seed-fixed configurations, oracle <-> product glue."""
from __future__ import annotations

import functools

import numpy as np

from oracle import pyoracle as po
from triplet_match_b200 import synth

DP = dict(distance_step_count=20.0, angle_step=0.17453292)
SP = dict(min_df=0.2, max_df=1.0)


@functools.lru_cache(maxsize=None)
def config(name: str):
    """(model cloud, scene cloud, oracle model, oracle scene, recorded pairs)."""
    shuffled = name.endswith("_shuffled")
    if shuffled:  # scene in seeded random order: the order the early-drop test assumes
        name = name[: -len("_shuffled")]
    if name == "plane_small":
        m = synth.plane_model(seed=2, size=0.3, res=0.01, n_curves=4)
        s = synth.make_scene(seed=5, model=m, n_points=20000, n_copies=3, extent=1.2)
        n_outer, ppo = 8, 24
    elif name == "cylinder_small":
        m = synth.cylinder_model(seed=1, radius=0.06, height=0.25, res=0.01, n_curves=3)
        s = synth.make_scene(seed=6, model=m, n_points=16000, n_copies=3, extent=1.0,
                             flat_copies=False)
        n_outer, ppo = 8, 24
    elif name == "freeform_small":
        m = synth.freeform_model(seed=3, n_points=1500, radius=0.12, n_bumps=6, n_curves=5)
        s = synth.make_scene(seed=7, model=m, n_points=24000, n_copies=4, extent=1.2,
                             flat_copies=False)
        n_outer, ppo = 10, 24
    else:
        raise KeyError(name)
    order = synth.shuffle_perm(13, 3, s.n) if shuffled else synth.morton_order(s.pos)
    s = s.take(order)
    om = po.OModel(m, **DP, **SP)
    osc = po.OScene(s)
    rec = synth.record_pairs(11, s, om.diameter, n_outer, ppo)
    return m, s, om, osc, rec


def upload_model(ctx, m, om, cap=200):
    from triplet_match_b200 import capi
    keys, offsets, pairs = om.table(cap)
    return capi.Model(ctx, m.pos, m.nrm, m.tgt, voxel=om.voxel, extents=om.extents,
                      to_voxel16=om.to_voxel16, resolution=om.resolution, diameter=om.diameter,
                      keys=keys, offsets=offsets, pairs=pairs, feat_min=om.feat_min,
                      feat_max=om.feat_max, distance_step_count=DP["distance_step_count"],
                      angle_step=DP["angle_step"])


def upload_scene(ctx, s):
    from triplet_match_b200 import capi
    return capi.Scene(ctx, s.pos, s.nrm, s.tgt, s.tangent_mask)
