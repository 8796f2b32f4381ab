"""Seed-fixed synthetic clouds and recorded sample lists.

The reference ships no data; every configuration is generated here from a
counter-based splitmix64 stream, so the same (seed, stream, index) always
yields the same value on any machine.  Points carry position, unit normal and
a tangent that is a unit vector on "feature curves" and zero elsewhere; the
tangent / curvature masks are supplied by the generator instead of being
derived with PCL (which is absent, and whose curvature ratio is NaN on exact
planes — SURVEY §7 "Degenerate curvature").

Sampling randomness of the reference (range-v3 sample/shuffle seeded from the
wall clock, include/impl/scene.hpp:122-128,143-144,285) is replaced by a
RECORDED list: outer samples p1, and for each an ordered list of second
samples p2, exactly what find_in_subset would draw.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed: int, stream: int, n: int, start: int = 0) -> np.ndarray:
    """n doubles in [0,1) from counter (seed, stream, start..start+n)."""
    with np.errstate(over="ignore"):
        base = splitmix64(np.array([seed], dtype=np.uint64))[0] ^ splitmix64(
            np.array([stream + 0x1234567], dtype=np.uint64))[0]
        ctr = np.arange(start, start + n, dtype=np.uint64) + base * np.uint64(0x2545F4914F6CDD1D)
    z = splitmix64(ctr)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed: int, stream: int, n: int) -> np.ndarray:
    u1 = uniform(seed, stream, n)
    u2 = uniform(seed, stream + 7919, n)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def shuffle_perm(seed: int, stream: int, n: int) -> np.ndarray:
    return np.argsort(uniform(seed, stream, n), kind="stable")


@dataclass
class Cloud:
    pos: np.ndarray            # (n,3) float32
    nrm: np.ndarray            # (n,3) float32
    tgt: np.ndarray            # (n,3) float32
    tangent_mask: np.ndarray   # (n,) uint8 — curvature criterion + ||tangent|| > 0.7
    poses: list = field(default_factory=list)  # ground-truth model->scene 4x4 (float64)

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    def take(self, idx: np.ndarray) -> "Cloud":
        return Cloud(np.ascontiguousarray(self.pos[idx]), np.ascontiguousarray(self.nrm[idx]),
                     np.ascontiguousarray(self.tgt[idx]),
                     np.ascontiguousarray(self.tangent_mask[idx]), list(self.poses))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))


def _unit(v: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return v / np.maximum(n, 1e-30)


def _mark_curve_points(uv: np.ndarray, segs: np.ndarray, half_width: float):
    """uv: (n,2) intrinsic coords; segs: (k,4) = (u0,v0,u1,v1).  Returns (on, dir2)."""
    on = np.zeros(uv.shape[0], dtype=bool)
    d2 = np.zeros((uv.shape[0], 2))
    for s in segs:
        a, b = s[:2], s[2:]
        ab = b - a
        L = np.linalg.norm(ab)
        if L <= 0:
            continue
        t = ((uv - a) @ ab) / (L * L)
        proj = a + np.clip(t, 0, 1)[:, None] * ab
        dist = np.linalg.norm(uv - proj, axis=1)
        hit = (dist <= half_width) & (t >= 0) & (t <= 1) & ~on
        on |= hit
        d2[hit] = ab / L
    return on, d2


def plane_model(seed: int = 2, size: float = 1.0, res: float = 0.01, n_curves: int = 6,
                jitter: float = 0.1) -> Cloud:
    """size x size plane patch on an r-spaced lattice with line-segment feature curves."""
    m = int(round(size / res)) + 1
    g = np.arange(m) * res
    u, v = np.meshgrid(g, g, indexing="xy")
    uv = np.stack([u.ravel(), v.ravel()], axis=1)
    n = uv.shape[0]
    r = uniform(seed, 11, 4 * n_curves).reshape(n_curves, 4)
    segs = np.empty((n_curves, 4))
    segs[:, 0] = 0.1 * size + 0.8 * size * r[:, 0]
    segs[:, 1] = 0.1 * size + 0.8 * size * r[:, 1]
    ang = 2 * np.pi * r[:, 2]
    ln = (0.4 + 0.5 * r[:, 3]) * size
    segs[:, 2] = np.clip(segs[:, 0] + ln * np.cos(ang), 0.02 * size, 0.98 * size)
    segs[:, 3] = np.clip(segs[:, 1] + ln * np.sin(ang), 0.02 * size, 0.98 * size)
    on, d2 = _mark_curve_points(uv, segs, 0.5 * res)
    pos = np.zeros((n, 3))
    pos[:, :2] = uv
    pos += jitter * res * np.stack([normal(seed, 21, n), normal(seed, 22, n), normal(seed, 23, n)], 1)
    nrm = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    tgt = np.zeros((n, 3))
    tgt[on, :2] = d2[on]
    return Cloud(_f32(pos), _f32(nrm), _f32(tgt), on.astype(np.uint8))


def cylinder_model(seed: int = 1, radius: float = 0.25, height: float = 1.0, res: float = 0.01,
                   n_curves: int = 4, jitter: float = 0.1) -> Cloud:
    """Cylinder on an r-spaced (theta, z) lattice with helical feature curves."""
    nth = int(round(2 * np.pi * radius / res))
    nz = int(round(height / res)) + 1
    th, z = np.meshgrid(np.arange(nth) * (2 * np.pi / nth), np.arange(nz) * res, indexing="xy")
    th, z = th.ravel(), z.ravel()
    n = th.size
    arc = th * radius
    uv = np.stack([arc, z], axis=1)
    r = uniform(seed, 11, 3 * n_curves).reshape(n_curves, 3)
    segs = np.empty((n_curves, 4))
    circ = 2 * np.pi * radius
    segs[:, 0] = 0.05 * circ + 0.3 * circ * r[:, 0]
    segs[:, 1] = 0.05 * height + 0.2 * height * r[:, 1]
    segs[:, 2] = segs[:, 0] + (0.3 + 0.3 * r[:, 2]) * circ
    segs[:, 3] = 0.95 * height - 0.2 * height * r[:, 1]
    on, d2 = _mark_curve_points(uv, segs, 0.5 * res)
    er = np.stack([np.cos(th), np.sin(th), np.zeros(n)], 1)
    et = np.stack([-np.sin(th), np.cos(th), np.zeros(n)], 1)
    ez = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    pos = radius * er + z[:, None] * ez
    pos += jitter * res * np.stack([normal(seed, 21, n), normal(seed, 22, n), normal(seed, 23, n)], 1)
    tgt = np.zeros((n, 3))
    tgt[on] = _unit(d2[on, 0:1] * et[on] + d2[on, 1:2] * ez[on])
    return Cloud(_f32(pos), _f32(er), _f32(tgt), on.astype(np.uint8))


def freeform_model(seed: int = 3, n_points: int = 50000, radius: float = 0.5, n_bumps: int = 12,
                   n_curves: int = 8) -> Cloud:
    """Bumpy sphere (sum of Gaussian bumps on a Fibonacci sphere) with great-circle arcs."""
    i = np.arange(n_points) + 0.5
    phi = np.arccos(1 - 2 * i / n_points)
    theta = np.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], 1)
    rb = uniform(seed, 31, 5 * n_bumps).reshape(n_bumps, 5)
    centres = _unit(np.stack([normal(seed, 32, n_bumps), normal(seed, 33, n_bumps),
                              normal(seed, 34, n_bumps)], 1))
    h = np.zeros(n_points)
    for b in range(n_bumps):
        ang = np.arccos(np.clip(d @ centres[b], -1, 1))
        h += (0.05 + 0.1 * rb[b, 0]) * radius * np.exp(-0.5 * (ang / (0.25 + 0.3 * rb[b, 1])) ** 2)
    pos = (radius + h)[:, None] * d
    res = np.sqrt(4 * np.pi * radius * radius / n_points)
    on = np.zeros(n_points, dtype=bool)
    tgt = np.zeros((n_points, 3))
    axes = _unit(np.stack([normal(seed, 41, n_curves), normal(seed, 42, n_curves),
                           normal(seed, 43, n_curves)], 1))
    ra = uniform(seed, 44, 2 * n_curves).reshape(n_curves, 2)
    for c in range(n_curves):
        a = axes[c]
        off = d @ a
        e1 = _unit(np.cross(a, [0.3, 0.5, 0.8]))
        e2 = np.cross(a, e1)
        az = np.arctan2(d @ e2, d @ e1)
        lo = -np.pi + 2 * np.pi * ra[c, 0]
        span = (0.4 + 0.8 * ra[c, 1]) * np.pi
        inarc = ((az - lo) % (2 * np.pi)) < span
        hit = (np.abs(off) * radius <= 0.5 * res) & inarc & ~on
        on |= hit
        tgt[hit] = _unit(np.cross(a, d[hit]))
    return Cloud(_f32(pos), _f32(d), _f32(tgt), on.astype(np.uint8))


def pyramid_model(seed: int = 7, size: float = 0.3, height: float = 0.12, res: float = 0.01,
                  normal_noise: float = 0.02, blend: float = 1.5) -> Cloud:
    """Square pyramid whose four slanted edges are genuine creases: the normal field turns across
    an edge within ~`blend` * res (and carries a little isotropic noise), so the reference's
    principal-curvature criterion pc_min / pc_max < 0.2 (scene.hpp:50, model.hpp:98) selects the
    edge points — unlike the analytic clouds above, whose exact normals give 0/0."""
    half = 0.5 * size
    n1 = int(round(size / res)) + 1
    g = np.linspace(-half, half, n1)
    x, y = np.meshgrid(g, g, indexing="xy")
    x, y = x.ravel(), y.ravel()
    n = x.size
    x = x + 0.1 * res * normal(seed, 1, n)
    y = y + 0.1 * res * normal(seed, 2, n)
    ax, ay = np.abs(x), np.abs(y)
    z = height * (1.0 - np.maximum(ax, ay) / half) + 0.05 * res * normal(seed, 3, n)
    # smooth max for the normal field only
    w = blend * res
    root = np.sqrt((ax - ay) ** 2 + w * w)
    dmx = (0.5 + 0.5 * (ax - ay) / root) * np.sign(x)
    dmy = (0.5 - 0.5 * (ax - ay) / root) * np.sign(y)
    slope = height / half
    nrm = _unit(np.stack([slope * dmx, slope * dmy, np.ones(n)], 1))
    nrm = _unit(nrm + normal_noise * np.stack([normal(seed, 4, n), normal(seed, 5, n), normal(seed, 6, n)], 1))
    d_edge = np.abs(ax - ay) / np.sqrt(2.0)
    m = np.maximum(ax, ay)
    on = (d_edge <= 0.5 * res) & (m > 0.1 * half) & (m < 0.95 * half)
    tgt = np.zeros((n, 3))
    e = _unit(np.stack([np.sign(x) * half, np.sign(y) * half, -height * np.ones(n)], 1))
    tgt[on] = e[on]
    return Cloud(_f32(np.stack([x, y, z], 1)), _f32(nrm), _f32(tgt), on.astype(np.uint8))


def _rot(axis: np.ndarray, angle: float) -> np.ndarray:
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def make_scene(seed: int, model: Cloud, n_points: int, n_copies: int = 3, extent: float = 3.0,
               res: float = 0.01, clutter_frac: float = 0.1, flat_copies: bool = True,
               noise_tangent_frac: float = 0.002, walls: bool = True, models=None) -> Cloud:
    """Ground plane (+ walls) + posed model copies + uniform clutter, exactly n_points.

    Roughly half of the copies lie flat (rotation about z, small lift), the rest are tilted.
    A fraction of floor points get random in-plane tangents ("false" feature points)."""
    parts_pos, parts_nrm, parts_tgt, parts_tm = [], [], [], []
    poses = []
    rp = uniform(seed, 51, 6 * n_copies).reshape(n_copies, 6)
    for c in range(n_copies):
        if models is not None:  # batched search: copy c is an instance of models[c % len(models)]
            model = models[c % len(models)]
        mc = model.pos.astype(np.float64).mean(axis=0)
        yaw = 2 * np.pi * rp[c, 0]
        R = _rot(np.array([0, 0, 1.0]), yaw)
        if not (flat_copies and c % 2 == 0):
            ax = _unit(np.array([np.cos(2 * np.pi * rp[c, 1]), np.sin(2 * np.pi * rp[c, 1]), 0.3]))
            R = _rot(ax, 0.2 + 0.6 * rp[c, 2]) @ R
        flat = flat_copies and c % 2 == 0
        # flat copies sit on a "table" well above the floor so that the floor itself never
        # falls inside the model's voxel grid when the planes align
        centre = np.array([0.15 * extent + 0.7 * extent * rp[c, 3],
                           0.15 * extent + 0.7 * extent * rp[c, 4],
                           (0.08 + 0.1 * rp[c, 5]) if flat else (0.15 + 0.3 * rp[c, 5])])
        t = centre - R @ mc
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = t
        poses.append(T)
        parts_pos.append(model.pos.astype(np.float64) @ R.T + t)
        parts_nrm.append(model.nrm.astype(np.float64) @ R.T)
        parts_tgt.append(model.tgt.astype(np.float64) @ R.T)
        parts_tm.append(model.tangent_mask.copy())
    n_model = sum(p.shape[0] for p in parts_pos)
    n_clutter = int(clutter_frac * n_points)
    n_struct = n_points - n_model - n_clutter
    if n_struct < 0:
        raise ValueError("n_points too small for the requested copies + clutter")
    # floor (+ two walls): lattice at spacing chosen so that the structure has >= n_struct points
    n_wall = int(0.2 * n_struct) if walls else 0
    n_floor = n_struct - 2 * (n_wall // 2)
    n_wall = n_wall // 2

    def lattice(count, w, h):
        if count <= 0:
            return np.zeros((0, 2))
        s = np.sqrt(w * h / count)
        nx, ny = int(np.ceil(w / s)) + 1, int(np.ceil(h / s)) + 1
        while nx * ny < count:
            nx += 1
        gx, gy = np.meshgrid(np.arange(nx) * (w / max(nx - 1, 1)), np.arange(ny) * (h / max(ny - 1, 1)), indexing="xy")
        return np.stack([gx.ravel(), gy.ravel()], 1)[:count]

    fl = lattice(n_floor, extent, extent)
    floor = np.zeros((fl.shape[0], 3))
    floor[:, :2] = fl
    floor += 0.1 * res * np.stack([normal(seed, 61, fl.shape[0]), normal(seed, 62, fl.shape[0]),
                                   normal(seed, 63, fl.shape[0])], 1)
    fn = np.tile(np.array([0.0, 0.0, 1.0]), (fl.shape[0], 1))
    ft = np.zeros((fl.shape[0], 3))
    ftm = np.zeros(fl.shape[0], dtype=np.uint8)
    k = int(noise_tangent_frac * fl.shape[0])
    if k > 0:
        sel = shuffle_perm(seed, 64, fl.shape[0])[:k]
        a = 2 * np.pi * uniform(seed, 65, k)
        ft[sel, 0], ft[sel, 1] = np.cos(a), np.sin(a)
        ftm[sel] = 1
    parts_pos.append(floor); parts_nrm.append(fn); parts_tgt.append(ft); parts_tm.append(ftm)
    wall_h = 0.4 * extent
    for w in range(2):
        wl = lattice(n_wall, extent, wall_h)
        p = np.zeros((wl.shape[0], 3))
        if w == 0:
            p[:, 0], p[:, 2] = wl[:, 0], wl[:, 1]
            nn = np.tile(np.array([0.0, 1.0, 0.0]), (wl.shape[0], 1))
        else:
            p[:, 1], p[:, 2] = wl[:, 0], wl[:, 1]
            nn = np.tile(np.array([1.0, 0.0, 0.0]), (wl.shape[0], 1))
        p += 0.1 * res * np.stack([normal(seed, 71 + w, wl.shape[0]), normal(seed, 73 + w, wl.shape[0]),
                                   normal(seed, 75 + w, wl.shape[0])], 1)
        parts_pos.append(p); parts_nrm.append(nn)
        parts_tgt.append(np.zeros_like(p)); parts_tm.append(np.zeros(wl.shape[0], dtype=np.uint8))
    if n_clutter:
        cp = np.stack([extent * uniform(seed, 81, n_clutter), extent * uniform(seed, 82, n_clutter),
                       0.4 * extent * uniform(seed, 83, n_clutter)], 1)
        cn = _unit(np.stack([normal(seed, 84, n_clutter), normal(seed, 85, n_clutter),
                             normal(seed, 86, n_clutter)], 1))
        parts_pos.append(cp); parts_nrm.append(cn)
        parts_tgt.append(np.zeros_like(cp)); parts_tm.append(np.zeros(n_clutter, dtype=np.uint8))
    pos = np.concatenate(parts_pos); nrm = np.concatenate(parts_nrm)
    tgt = np.concatenate(parts_tgt); tm = np.concatenate(parts_tm)
    assert pos.shape[0] == n_points, (pos.shape[0], n_points)
    return Cloud(_f32(pos), _f32(nrm), _f32(tgt), tm.astype(np.uint8), poses)


def morton_order(pos: np.ndarray, bits: int = 10) -> np.ndarray:
    """Permutation sorting points along a 3-D Morton curve (spatial locality of subset tiles)."""
    p = pos.astype(np.float64)
    lo, hi = p.min(axis=0), p.max(axis=0)
    q = ((p - lo) / np.maximum(hi - lo, 1e-30) * ((1 << bits) - 1)).astype(np.uint64)

    def spread(v):
        v = v & np.uint64(0x3FF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x030000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x0300F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x030C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x09249249)
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    return np.argsort(code, kind="stable")


@dataclass
class Recorded:
    outer: np.ndarray       # (n_outer,) uint32 scene index of p1
    pair_outer: np.ndarray  # (n_pairs,) uint32 index into outer (sorted)
    pair_j: np.ndarray      # (n_pairs,) uint32 scene index of p2

    @property
    def pair_i(self) -> np.ndarray:
        return self.outer[self.pair_outer]


def record_pairs(seed: int, scene: Cloud, radius: float, n_outer: int, pairs_per_outer: int,
                 min_dist: float = 0.0) -> Recorded:
    """Recorded sample list: outer samples drawn from the tangent points, and for each the
    first `pairs_per_outer` entries of a seeded shuffle of the tangent points inside its ball
    (what find_in_subset's shuffled inner loop would visit; scene.hpp:284-290)."""
    tidx = np.nonzero(scene.tangent_mask)[0].astype(np.uint32)
    if tidx.size == 0:
        return Recorded(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    perm = shuffle_perm(seed, 91, tidx.size)
    outer = tidx[perm[:min(n_outer, tidx.size)]]
    tp = scene.pos[tidx].astype(np.float64)
    po, pj = [], []
    for o, i in enumerate(outer):
        d2 = ((tp - scene.pos[i].astype(np.float64)) ** 2).sum(axis=1)
        cand = tidx[(d2 < radius * radius) & (d2 >= min_dist * min_dist) & (tidx != i)]
        if cand.size == 0:
            continue
        sh = shuffle_perm(seed, 1000 + o, cand.size)[:pairs_per_outer]
        pj.append(cand[sh])
        po.append(np.full(sh.size, o, dtype=np.uint32))
    if not pj:
        return Recorded(outer.astype(np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    return Recorded(outer.astype(np.uint32), np.concatenate(po).astype(np.uint32),
                    np.concatenate(pj).astype(np.uint32))
