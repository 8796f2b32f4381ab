"""ctypes wrapper of oracle/liboracle.so — the CPU ORACLE (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module; the product (triplet_match_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s", LIB_PATH])
    L = C.CDLL(LIB_PATH)
    L.orc_murmur4.restype = C.c_uint32
    L.orc_discretize_range.restype = C.c_uint32
    L.orc_discretize_range.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32]
    L.orc_discretize_step.restype = C.c_uint32
    L.orc_discretize_step.argtypes = [C.c_float, C.c_float]
    for f in ("orc_atan2f_q1", "orc_atan2f_full", "orc_libm_atan2f"):
        getattr(L, f).restype = C.c_float
        getattr(L, f).argtypes = [C.c_float, C.c_float]
    L.orc_early_drop_upper.restype = C.c_uint32
    L.orc_early_drop_upper.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    L.orc_resolution.restype = C.c_float
    L.orc_model_create.restype = C.c_void_p
    L.orc_model_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    L.orc_model_destroy.argtypes = [C.c_void_p]
    L.orc_model_destroy.restype = None
    L.orc_model_table.restype = C.c_uint64
    L.orc_model_query.restype = C.c_uint32
    L.orc_scene_create.restype = C.c_void_p
    L.orc_scene_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_void_p]
    L.orc_scene_destroy.argtypes = [C.c_void_p]
    L.orc_scene_destroy.restype = None
    L.orc_ball_subset.restype = C.c_uint64
    L.orc_project.restype = C.c_uint32
    L.orc_hypotheses.restype = C.c_uint64
    L.orc_icp.restype = C.c_uint32
    L.orc_get_octant.restype = C.c_uint32
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def murmur4(key) -> int:
    k = np.ascontiguousarray(key, dtype=np.uint32)
    return int(load().orc_murmur4(_p(k)))


def feature(p0, t0, p1, t1, use_libm=False) -> np.ndarray:
    inp = _f32(np.concatenate([p0, t0, p1, t1]))
    f = np.zeros(4, dtype=np.float32)
    load().orc_feature(_p(inp), C.c_int(1 if use_libm else 0), _p(f))
    return f


def base_transform(src_i, src_j, src_t, tgt_i, tgt_j, tgt_t) -> np.ndarray:
    inp = _f32(np.concatenate([src_i, src_j, src_t, tgt_i, tgt_j, tgt_t]))
    out = np.zeros(16, dtype=np.float32)
    load().orc_base_transform(_p(inp), _p(out))
    return out


def atan2f_q1_batch(y, x):
    y, x = _f32(y), _f32(x)
    o = np.empty_like(y)
    l = np.empty_like(y)
    load().orc_atan2f_q1_batch(_p(y), _p(x), C.c_uint64(y.size), _p(o), _p(l))
    return o, l


def early_drop_tests(nsub: int) -> np.ndarray:
    out = np.zeros(18, dtype=np.uint32)
    load().orc_early_drop_tests(C.c_uint64(nsub), _p(out))
    return out


def umeyama(src, dst) -> np.ndarray:
    s, d = _f32(src, (-1, 3)), _f32(dst, (-1, 3))
    out = np.zeros(16, dtype=np.float32)
    load().orc_umeyama(_p(s), _p(d), C.c_uint32(s.shape[0]), _p(out))
    return out


def traits_project(kind, g2l16, radius, threshold, xyz):
    g = _f32(g2l16, (16,))
    xyz = _f32(xyz, (-1, 3))
    uvw = np.zeros_like(xyz)
    ok = np.zeros(xyz.shape[0], dtype=np.uint8)
    L = load()
    for i in range(xyz.shape[0]):
        o = np.zeros(3, dtype=np.float32)
        ok[i] = L.orc_traits_project(C.c_int(kind), _p(g), C.c_float(radius), C.c_float(threshold),
                                     _p(xyz[i]), _p(o))
        uvw[i] = o
    return uvw, ok


def knn(pos, query, k):
    """pointcloud::knn_inclusive by brute force: ascending (d^2, index)."""
    pos = _f32(pos, (-1, 3))
    q = np.ascontiguousarray(query, dtype=np.uint32)
    idx = np.zeros((q.size, k), dtype=np.int32)
    d2 = np.zeros((q.size, k), dtype=np.float32)
    load().orc_knn(_p(pos), C.c_uint32(pos.shape[0]), _p(q), C.c_uint32(q.size), C.c_uint32(k), _p(idx), _p(d2))
    return idx, d2


def curvature(pos, nrm, query, k, nbr=None):
    """pointcloud::curvature(k, idx): (pc_min, pc_max, cov 3x3) per query point."""
    pos, nrm = _f32(pos, (-1, 3)), _f32(nrm, (-1, 3))
    q = np.ascontiguousarray(query, dtype=np.uint32)
    if nbr is None:
        nbr, _ = knn(pos, q, k)
    nbr = np.ascontiguousarray(nbr, dtype=np.int32)
    mn = np.zeros(q.size, dtype=np.float32)
    mx = np.zeros(q.size, dtype=np.float32)
    cov = np.zeros((q.size, 9), dtype=np.float32)
    load().orc_curvature(_p(pos), _p(nrm), C.c_uint32(pos.shape[0]), _p(q), C.c_uint32(q.size), C.c_uint32(k),
                         _p(nbr), _p(mn), _p(mx), _p(cov))
    return mn, mx, cov.reshape(-1, 3, 3)


def eigen33(cov):
    ev = np.zeros(3, dtype=np.float32)
    load().orc_eigen33(_p(_f32(cov, (9,))), _p(ev))
    return ev


def tangent_mask(cloud, k=30, ratio=0.2):
    """scene.hpp:50 / model.hpp:98: ||tangent|| > 0.7 and pc_min / pc_max < ratio."""
    t = _f32(cloud.tgt, (-1, 3))
    nrm = np.sqrt((t[:, 0] * t[:, 0] + (t[:, 1] * t[:, 1] + t[:, 2] * t[:, 2])).astype(np.float32)).astype(np.float32)
    cand = np.flatnonzero(nrm > np.float32(0.7)).astype(np.uint32)
    mn, mx, _ = curvature(cloud.pos, cloud.nrm, cand, k)
    mask = np.zeros(t.shape[0], dtype=np.uint8)
    with np.errstate(divide="ignore", invalid="ignore"):
        mask[cand] = ((mn / mx) < np.float32(ratio)).astype(np.uint8)
    return mask, cand, mn, mx


def cl_icp_projection(projector, pnts4, image4, img_size, img_margin, mat_align, mat_uvw, mat_proj, mat_norm,
                      max_corr_dist):
    """opencl/icp.cl:1-53 (+ cylinder.cl / util.cl) over all work-items."""
    pn = _f32(pnts4, (-1, 4))
    im = _f32(image4, (-1, 4))
    n = pn.shape[0]
    sz = np.ascontiguousarray(img_size, dtype=np.int32)
    mg = np.ascontiguousarray(img_margin, dtype=np.int32)
    op = np.zeros((n, 4), dtype=np.float32)
    mi = np.zeros(n, dtype=np.int32)
    si = np.zeros(n, dtype=np.int32)
    L = load()
    L.orc_cl_icp_projection.restype = C.c_uint32
    c = L.orc_cl_icp_projection(C.c_int(projector), _p(pn), C.c_int(n), _p(im), _p(sz), _p(mg), _p(_f32(mat_align, (16,))),
                                _p(_f32(mat_uvw, (16,))), _p(_f32(mat_proj, (16,))), _p(_f32(mat_norm, (16,))),
                                C.c_float(max_corr_dist), _p(op), _p(mi), _p(si))
    return op, mi, si, int(c)


def cl_icp_correlation(scene4, model4, indices_scene, indices_model, centroid_scene, centroid_model):
    """opencl/icp.cl:55-86 over all work-items + the sum of the records in double."""
    sc, md = _f32(scene4, (-1, 4)), _f32(model4, (-1, 4))
    is_ = np.ascontiguousarray(indices_scene, dtype=np.int32)
    im_ = np.ascontiguousarray(indices_model, dtype=np.int32)
    n = is_.shape[0]
    rec = np.zeros((max(n, 1), 16), dtype=np.float32)
    cov = np.zeros(9, dtype=np.float64)
    load().orc_cl_icp_correlation(_p(sc), _p(md), _p(is_), _p(im_), C.c_int(n), _p(_f32(centroid_scene, (4,))),
                                  _p(_f32(centroid_model, (4,))), _p(rec), _p(cov))
    return rec[:n], cov


class OModel:
    def __init__(self, cloud, distance_step_count=20.0, angle_step=0.17453292, min_df=0.2,
                 max_df=1.0, resolution=-1.0, curv_ok=None, voxel=None, subset=None):
        """voxel: a grid supplied instead of filled (the brute-force fill is O(cells x points)); spot-check it
        with cell_nearest()."""
        L = load()
        self.pos, self.nrm, self.tgt = _f32(cloud.pos), _f32(cloud.nrm), _f32(cloud.tgt)
        self.n = self.pos.shape[0]
        co = cloud.tangent_mask if curv_ok is None else curv_ok
        co = None if co is None else np.ascontiguousarray(co, dtype=np.uint8)
        if subset is not None:  # model::init(subset, params)
            ins = np.zeros(self.n, dtype=np.uint8)
            ins[np.asarray(subset, dtype=np.int64)] = 1
            L.orc_model_create_subset.restype = C.c_void_p
            self.h = C.c_void_p(L.orc_model_create_subset(_p(self.pos), _p(self.nrm), _p(self.tgt), C.c_uint32(self.n),
                                                          _p(ins), _p(co), C.c_float(distance_step_count),
                                                          C.c_float(angle_step), C.c_float(min_df), C.c_float(max_df),
                                                          C.c_float(resolution)))
        elif voxel is None:
            self.h = C.c_void_p(L.orc_model_create(_p(self.pos), _p(self.nrm), _p(self.tgt),
                                                   C.c_uint32(self.n), _p(co),
                                                   C.c_float(distance_step_count), C.c_float(angle_step),
                                                   C.c_float(min_df), C.c_float(max_df),
                                                   C.c_float(resolution)))
        else:
            vin = np.ascontiguousarray(voxel, dtype=np.uint32)
            L.orc_model_create_with_grid.restype = C.c_void_p
            self.h = C.c_void_p(L.orc_model_create_with_grid(_p(self.pos), _p(self.nrm), _p(self.tgt),
                                                             C.c_uint32(self.n), _p(co),
                                                             C.c_float(distance_step_count), C.c_float(angle_step),
                                                             C.c_float(min_df), C.c_float(max_df),
                                                             C.c_float(resolution), _p(vin)))
        f16 = np.zeros(16, dtype=np.float32)
        i4 = np.zeros(4, dtype=np.int32)
        c3 = np.zeros(3, dtype=np.uint64)
        L.orc_model_info(self.h, _p(f16), _p(i4), _p(c3))
        self.resolution, self.diameter = float(f16[0]), float(f16[1])
        self.scale, self.trans = f16[2:5].copy(), f16[5:8].copy()
        self.feat_min, self.feat_max = f16[8:12].copy(), f16[12:16].copy()
        self.extents, self.margin = i4[:3].copy(), int(i4[3])
        self.n_subset, self.n_entries, self.n_keys = int(c3[0]), int(c3[1]), int(c3[2])
        self.distance_step_count, self.angle_step = distance_step_count, angle_step
        self.voxel = np.zeros(int(np.prod(self.extents.astype(np.int64))), dtype=np.uint32)
        L.orc_model_voxels(self.h, _p(self.voxel))
        self.subset = np.zeros(self.n_subset, dtype=np.uint32)
        if self.n_subset:
            L.orc_model_subset(self.h, _p(self.subset))

    @property
    def to_voxel16(self) -> np.ndarray:
        t = np.zeros((4, 4), dtype=np.float32)  # t[col][row] (column-major)
        for k in range(3):
            t[k, k] = self.scale[k]
            t[3, k] = self.trans[k]
        t[3, 3] = 1.0
        return t.ravel()

    def table(self, cap=200):
        L = load()
        total = int(L.orc_model_table(self.h, C.c_uint32(cap), None, None, None))
        keys = np.zeros((self.n_keys, 4), dtype=np.uint32)
        offsets = np.zeros(self.n_keys + 1, dtype=np.uint32)
        pairs = np.zeros((max(total, 1), 2), dtype=np.uint32)
        L.orc_model_table(self.h, C.c_uint32(cap), _p(keys), _p(offsets), _p(pairs))
        return keys, offsets, pairs[:total]

    def query(self, f, limit=200):
        f = _f32(f, (4,))
        out = np.zeros((max(limit, 1) if limit else self.n_entries, 2), dtype=np.uint32)
        n = load().orc_model_query(self.h, _p(f), C.c_uint32(limit), _p(out))
        return out[:n]

    def cell_nearest(self, ijk) -> np.ndarray:
        """model.hpp:87-88 for the given cells (n x 3 int32): exact brute-force nearest model point."""
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        out = np.zeros(ijk.shape[0], dtype=np.uint32)
        load().orc_model_cell_nearest(self.h, _p(ijk), C.c_uint64(ijk.shape[0]), _p(out))
        return out

    def voxel_query(self, pos4):
        p = _f32(pos4, (4,))
        o = C.c_uint32()
        ok = load().orc_model_voxel_query(self.h, _p(p), C.byref(o))
        return int(o.value) if ok else None

    def close(self):
        if self.h:
            load().orc_model_destroy(self.h)
            self.h = C.c_void_p()


class OScene:
    def __init__(self, cloud, mask=None):
        L = load()
        self.pos, self.nrm, self.tgt = _f32(cloud.pos), _f32(cloud.nrm), _f32(cloud.tgt)
        self.n = self.pos.shape[0]
        tm = np.ascontiguousarray(cloud.tangent_mask, dtype=np.uint8)
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.h = C.c_void_p(L.orc_scene_create(_p(self.pos), _p(self.nrm), _p(self.tgt),
                                               C.c_uint32(self.n), _p(tm), _p(mk)))

    def set_mask(self, mask):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        load().orc_scene_set_mask(self.h, _p(m))

    def ball_subset(self, idx: int, radius: float) -> np.ndarray:
        L = load()
        n = int(L.orc_ball_subset(self.h, C.c_uint32(idx), C.c_float(radius), None))
        out = np.zeros(max(n, 1), dtype=np.int32)
        L.orc_ball_subset(self.h, C.c_uint32(idx), C.c_float(radius), _p(out))
        return out[:n]

    def pair_features(self, model: OModel, pi, pj, min_df=0.2, max_df=1.0):
        pi = np.ascontiguousarray(pi, dtype=np.uint32)
        pj = np.ascontiguousarray(pj, dtype=np.uint32)
        n = pi.shape[0]
        feats = np.zeros((n, 4), dtype=np.float32)
        keys = np.zeros((n, 4), dtype=np.uint32)
        valid = np.zeros(n, dtype=np.uint8)
        load().orc_pair_features(self.h, model.h, _p(pi), _p(pj), C.c_uint64(n), C.c_float(min_df),
                                 C.c_float(max_df), _p(feats), _p(keys), _p(valid))
        return feats, keys, valid

    def hypotheses(self, model: OModel, pi, pj, min_df=0.2, max_df=1.0, limit=200, force_up=False):
        L = load()
        pi = np.ascontiguousarray(pi, dtype=np.uint32)
        pj = np.ascontiguousarray(pj, dtype=np.uint32)
        n = pi.shape[0]
        args = (self.h, model.h, _p(pi), _p(pj), C.c_uint64(n), C.c_float(min_df), C.c_float(max_df),
                C.c_uint32(limit), C.c_int(1 if force_up else 0))
        nh = int(L.orc_hypotheses(*args, None, None, None, None, None))
        T = np.zeros((max(nh, 1), 16), dtype=np.float32)
        hp = np.zeros(max(nh, 1), dtype=np.uint32)
        mi = np.zeros(max(nh, 1), dtype=np.uint32)
        mj = np.zeros(max(nh, 1), dtype=np.uint32)
        va = np.zeros(max(nh, 1), dtype=np.uint8)
        L.orc_hypotheses(*args, _p(T), _p(hp), _p(mi), _p(mj), _p(va))
        return T[:nh], hp[:nh], mi[:nh], mj[:nh], va[:nh]

    def project(self, model: OModel, subset, T16, accept_prob=0.5, dist_thres=1.0, early_out=False):
        sub = np.ascontiguousarray(subset, dtype=np.int32)
        T = _f32(T16, (16,))
        sc = np.zeros(max(sub.size, 1), dtype=np.uint32)
        mc = np.zeros(max(sub.size, 1), dtype=np.uint32)
        score = C.c_double()
        saved = C.c_uint32()
        dropped = C.c_int()
        n = load().orc_project(self.h, model.h, _p(sub), C.c_uint64(sub.size), _p(T),
                               C.c_float(accept_prob), C.c_float(dist_thres),
                               C.c_int(1 if early_out else 0), _p(sc), _p(mc), C.byref(score),
                               C.byref(saved), C.byref(dropped))
        return dict(count=int(n), scene_corrs=sc[:n].copy(), model_corrs=mc[:n].copy(),
                    score=float(score.value), saved=int(saved.value), dropped=bool(dropped.value))

    def score_batch(self, model: OModel, T16s, hyp_sub=None, sub_off=None, sub_idx=None,
                    accept_prob=0.5, dist_thres=1.0, early_out=False, nthreads=1):
        T = _f32(T16s, (-1, 16))
        n = T.shape[0]
        counts = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.float64)
        dropped = np.zeros(n, dtype=np.uint8)
        hs = so = si = None
        if hyp_sub is not None:
            hs = np.ascontiguousarray(hyp_sub, dtype=np.uint32)
            so = np.ascontiguousarray(sub_off, dtype=np.uint64)
            si = np.ascontiguousarray(sub_idx, dtype=np.int32)
        load().orc_score_batch(self.h, model.h, _p(T), C.c_uint64(n), _p(hs), _p(so), _p(si),
                               C.c_float(accept_prob), C.c_float(dist_thres),
                               C.c_int(1 if early_out else 0), C.c_int(nthreads), _p(counts),
                               _p(scores), _p(dropped))
        return counts, scores, dropped

    def icp(self, model: OModel, T16, max_iterations=5, dist_thres=1.0, accept_prob=0.5):
        T = _f32(T16, (16,))
        out = np.zeros(16, dtype=np.float32)
        score = C.c_double()
        iters = C.c_uint32()
        n = load().orc_icp(self.h, model.h, _p(T), C.c_uint32(max_iterations), C.c_float(dist_thres),
                           C.c_float(accept_prob), _p(out), C.byref(score), C.byref(iters))
        return out, int(n), float(score.value), int(iters.value)

    def close(self):
        if self.h:
            load().orc_scene_destroy(self.h)
            self.h = C.c_void_p()
