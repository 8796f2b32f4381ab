"""ctypes binding of the C-ABI in include/tm_b200.h (harness side: tests, bench, smoke).

This is a thin mirror: every method maps to one exported function and takes /
returns numpy arrays living on the HOST.  There is no fallback — if the shared
library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtriplet_match_b200" + os.environ.get("TM_LIB_SUFFIX", "") + ".so")

TM_OK = 0
TM_ERR_INVALID, TM_ERR_CUDA, TM_ERR_CAPACITY, TM_ERR_UNINITIALIZED, TM_ERR_NCCL = 1, 2, 3, 4, 5


class TmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tm_b200 error {code}: {msg}")
        self.code = code


class CloudView(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("nrm", C.c_void_p), ("tgt", C.c_void_p),
                ("stride", C.c_uint32), ("n", C.c_uint32)]


class ModelDesc(C.Structure):
    _fields_ = [("voxel", C.c_void_p), ("extents", C.c_int32 * 3), ("to_voxel", C.c_float * 16),
                ("resolution", C.c_float), ("diameter", C.c_float),
                ("keys", C.c_void_p), ("offsets", C.c_void_p), ("pairs", C.c_void_p),
                ("n_keys", C.c_uint32),
                ("feat_min", C.c_float * 4), ("feat_max", C.c_float * 4),
                ("distance_step_count", C.c_float), ("angle_step", C.c_float)]


class QueryParams(C.Structure):
    _fields_ = [("min_diameter_factor", C.c_float), ("max_diameter_factor", C.c_float),
                ("force_up", C.c_int32), ("query_limit", C.c_uint32),
                ("dist_thres", C.c_float), ("accept_prob", C.c_float),
                ("early_out", C.c_int32), ("icp_top_k", C.c_uint32),
                ("max_icp_iterations", C.c_uint32),
                ("max_hypotheses", C.c_uint64), ("hyp_limit", C.c_uint64)]


class QueryResult(C.Structure):
    _fields_ = [("n_pairs_valid", C.c_uint64), ("n_hypotheses", C.c_uint64),
                ("n_scored", C.c_uint64), ("n_tests", C.c_uint64), ("best_key", C.c_uint64),
                ("best_hypothesis", C.c_uint32), ("best_inliers", C.c_uint32),
                ("best_score", C.c_double), ("best_T", C.c_float * 16)]


EXPORTS = [
    "tm_last_error", "tm_version", "tm_ctx_create", "tm_ctx_destroy", "tm_ctx_sync",
    "tm_ctx_stream", "tm_ctx_sm_count", "tm_timer_start", "tm_timer_stop", "tm_ctx_flush_l2",
    "tm_ctx_kernel_launches", "tm_ctx_scan_u64", "tm_ctx_measure_l2_gather", "tm_model_upload", "tm_model_destroy", "tm_voxel_fill",
    "tm_scene_upload", "tm_scene_upload_sorted", "tm_scene_set_mask", "tm_scene_destroy", "tm_features", "tm_probe",
    "tm_hypotheses", "tm_ball_subsets", "tm_score", "tm_walk_stride", "tm_correspondences", "tm_correspondences_batch", "tm_ctx_select_topk", "tm_icp", "tm_icp_pose_sharded", "tm_query_set_balance", "tm_query_frontend_ms", "tm_query_early_walked", "tm_early_level_begin",
    "tm_traits_project", "tm_scene_knn", "tm_scene_curvature", "tm_scene_tangent_mask", "tm_uvicp_projection", "tm_uvicp_correlation", "tm_query_create", "tm_query_destroy", "tm_query_set_pairs",
    "tm_query_set_shard", "tm_query_run", "tm_query_result_get", "tm_query_best_key_device",
    "tm_query_score_kernel_ms",
    "tm_query_set_global_best", "tm_query_download", "tm_query_icp_results",
    "tm_nccl_unique_id", "tm_comm_create", "tm_comm_destroy", "tm_query_allreduce_best", "tm_queries_allreduce_best", "tm_icp_sharded",
]
HOST_EXPORTS = [
    "tm_host_last_error", "tm_host_resolution", "tm_hostmodel_build", "tm_hostmodel_build_subset", "tm_hostmodel_destroy",
    "tm_hostmodel_desc", "tm_hostmodel_counts", "tm_hostmodel_subset", "tm_hostmodel_entry_keys",
    "tm_hostmodel_entry_pairs", "tm_model_create", "tm_hostmodel_save", "tm_hostmodel_load",
]

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.tm_last_error.restype = C.c_char_p
    lib.tm_version.restype = C.c_char_p
    lib.tm_ctx_stream.restype = C.c_void_p
    lib.tm_ctx_stream.argtypes = [C.c_void_p]
    lib.tm_ctx_kernel_launches.restype = C.c_uint64
    lib.tm_ctx_kernel_launches.argtypes = [C.c_void_p]
    lib.tm_query_best_key_device.restype = C.c_void_p
    lib.tm_query_best_key_device.argtypes = [C.c_void_p]
    lib.tm_host_last_error.restype = C.c_char_p
    lib.tm_host_resolution.restype = C.c_float
    for name in ("tm_hostmodel_subset", "tm_hostmodel_entry_keys", "tm_hostmodel_entry_pairs"):
        getattr(lib, name).restype = C.c_void_p
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.tm_hostmodel_desc.restype = None
    lib.tm_hostmodel_counts.restype = None
    for name in ("tm_ctx_destroy", "tm_model_destroy", "tm_scene_destroy", "tm_query_destroy",
                 "tm_comm_destroy", "tm_hostmodel_destroy"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _chk(rc: int) -> None:
    if rc != TM_OK:
        raise TmError(rc, load().tm_last_error().decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _view(pos, nrm, tgt):
    pos, nrm, tgt = _f32(pos), _f32(nrm), _f32(tgt)
    v = CloudView(pos.ctypes.data, nrm.ctypes.data, tgt.ctypes.data, 3, pos.shape[0])
    return v, (pos, nrm, tgt)


def surfel_view(records: np.ndarray):
    """View over an (n,12) float32 array laid out like pcl::PointSurfel (48 B)."""
    assert records.dtype == np.float32 and records.ndim == 2 and records.shape[1] == 12
    base = records.ctypes.data
    return CloudView(base, base + 16, base + 36, 12, records.shape[0])


class Context:
    def __init__(self, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        _chk(self.lib.tm_ctx_create(C.c_int(device), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.tm_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def sync(self):
        _chk(self.lib.tm_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.tm_ctx_stream(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.lib.tm_ctx_sm_count(self.h))

    def timer_start(self):
        _chk(self.lib.tm_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _chk(self.lib.tm_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        _chk(self.lib.tm_ctx_flush_l2(self.h))

    def select_topk(self, counts, k: int, excluded=None) -> np.ndarray:
        """tm_ctx_select_topk (self-test hook): ids of the k largest counts (ties: lower index first)."""
        cnt = np.ascontiguousarray(counts, dtype=np.uint32)
        ex = None if excluded is None else np.ascontiguousarray(excluded, dtype=np.uint8)
        ids = np.zeros(k, dtype=np.uint32)
        _chk(self.lib.tm_ctx_select_topk(self.h, _p(cnt), _p(ex), C.c_uint32(cnt.size), C.c_uint32(k), _p(ids)))
        return ids

    def scan_u64(self, values):
        """tm_ctx_scan_u64: exclusive prefix sum (n + 1 entries) on the device."""
        v = np.ascontiguousarray(values, dtype=np.uint32)
        out = np.zeros(v.size + 1, dtype=np.uint64)
        _chk(self.lib.tm_ctx_scan_u64(self.h, _p(v), C.c_uint64(v.size), _p(out)))
        return out

    def measure_l2_gather(self, working_set_bytes: int = 32 << 20) -> float:
        g = C.c_double()
        _chk(self.lib.tm_ctx_measure_l2_gather(self.h, C.c_uint64(working_set_bytes), C.byref(g)))
        return float(g.value)

    def kernel_launches(self) -> int:
        return int(self.lib.tm_ctx_kernel_launches(self.h))

    def voxel_fill(self, pos, nrm, tgt, extents, to_voxel16) -> np.ndarray:
        v, keep = _view(pos, nrm, tgt)
        ext = (C.c_int32 * 3)(*[int(e) for e in extents])
        tv = _f32(to_voxel16, (16,))
        out = np.empty(int(extents[0]) * int(extents[1]) * int(extents[2]), dtype=np.uint32)
        _chk(self.lib.tm_voxel_fill(self.h, C.byref(v), ext, _p(tv), _p(out)))
        return out

    def traits_project(self, kind: int, g2l16, radius: float, threshold: float, xyz):
        xyz = _f32(xyz, (-1, 3))
        g = _f32(g2l16, (16,))
        uvw = np.empty_like(xyz)
        ok = np.empty(xyz.shape[0], dtype=np.uint8)
        _chk(self.lib.tm_traits_project(self.h, C.c_int(kind), _p(g), C.c_float(radius),
                                        C.c_float(threshold), _p(xyz), C.c_uint64(xyz.shape[0]),
                                        _p(uvw), _p(ok)))
        return uvw, ok


    def uvicp_projection(self, projector, pnts4, image4, img_size, img_margin, mat_align, mat_uvw, mat_proj,
                         mat_norm, max_corr_dist):
        """tm_uvicp_projection (opencl/icp.cl icp_projection)."""
        pn, im = _f32(pnts4, (-1, 4)), _f32(image4, (-1, 4))
        n = pn.shape[0]
        sz = np.ascontiguousarray(img_size, dtype=np.int32)
        mg = np.ascontiguousarray(img_margin, dtype=np.int32)
        op = np.zeros((max(n, 1), 4), dtype=np.float32)
        mi = np.zeros(max(n, 1), dtype=np.int32)
        si = np.zeros(max(n, 1), dtype=np.int32)
        nc = C.c_uint32()
        _chk(self.lib.tm_uvicp_projection(self.h, C.c_int(projector), _p(pn), C.c_int32(n), _p(im), _p(sz), _p(mg),
                                          _p(_f32(mat_align, (16,))), _p(_f32(mat_uvw, (16,))),
                                          _p(_f32(mat_proj, (16,))), _p(_f32(mat_norm, (16,))),
                                          C.c_float(max_corr_dist), _p(op), _p(mi), _p(si), C.byref(nc)))
        return op[:n], mi[:n], si[:n], int(nc.value)

    def uvicp_correlation(self, scene4, model4, indices_scene, indices_model, centroid_scene, centroid_model,
                          want_records=True):
        """tm_uvicp_correlation (opencl/icp.cl icp_correlation + fused reduction)."""
        sc, md = _f32(scene4, (-1, 4)), _f32(model4, (-1, 4))
        is_ = np.ascontiguousarray(indices_scene, dtype=np.int32)
        im_ = np.ascontiguousarray(indices_model, dtype=np.int32)
        n = is_.shape[0]
        rec = np.zeros((max(n, 1), 16), dtype=np.float32) if want_records else None
        cov = np.zeros(9, dtype=np.float64)
        _chk(self.lib.tm_uvicp_correlation(self.h, _p(sc), C.c_uint32(sc.shape[0]), _p(md), C.c_uint32(md.shape[0]),
                                           _p(is_), _p(im_), C.c_int32(n), _p(_f32(centroid_scene, (4,))),
                                           _p(_f32(centroid_model, (4,))), _p(rec), _p(cov)))
        return (rec[:n] if rec is not None else None), cov


def host_resolution(pos) -> float:
    pos = _f32(pos, (-1, 3))
    v = CloudView(pos.ctypes.data, pos.ctypes.data, pos.ctypes.data, 3, pos.shape[0])
    return float(load().tm_host_resolution(C.byref(v)))


class HostModel:
    """model::init on the host (+ GPU grid fill when ctx is given): tm_hostmodel_build."""

    def __init__(self, ctx, pos, nrm, tgt, curv_ok=None, distance_step_count=20.0,
                 angle_step=0.17453292, min_df=0.2, max_df=1.0, resolution=-1.0, cap=200, subset=None):
        self.lib = load()
        v, self._keep = _view(pos, nrm, tgt)
        self._view = v
        co = None if curv_ok is None else np.ascontiguousarray(curv_ok, dtype=np.uint8)
        ins = None
        if subset is not None:  # model::init(subset, params)
            ins = np.zeros(int(v.n), dtype=np.uint8)
            ins[np.asarray(subset, dtype=np.int64)] = 1
        self.h = C.c_void_p()
        rc = self.lib.tm_hostmodel_build_subset(ctx.h if ctx is not None else None, C.byref(v), _p(ins), _p(co),
                                                C.c_float(distance_step_count), C.c_float(angle_step),
                                                C.c_float(min_df), C.c_float(max_df), C.c_float(resolution),
                                                C.c_uint32(cap), C.byref(self.h))
        if rc != TM_OK:
            raise TmError(rc, self.lib.tm_host_last_error().decode("utf-8", "replace"))
        d = ModelDesc()
        self.lib.tm_hostmodel_desc(self.h, C.byref(d))
        self.desc = d
        c = (C.c_uint64 * 4)()
        self.lib.tm_hostmodel_counts(self.h, C.byref(c, 0), C.byref(c, 8), C.byref(c, 16), C.byref(c, 24))
        self.n_subset, self.n_entries, self.n_keys, self.n_kept = [int(x) for x in c]
        self.extents = np.array(list(d.extents), dtype=np.int32)
        self.to_voxel16 = np.array(list(d.to_voxel), dtype=np.float32)
        self.resolution, self.diameter = float(d.resolution), float(d.diameter)
        self.feat_min = np.array(list(d.feat_min), dtype=np.float32)
        self.feat_max = np.array(list(d.feat_max), dtype=np.float32)

    def _read_desc(self):
        d = ModelDesc()
        self.lib.tm_hostmodel_desc(self.h, C.byref(d))
        self.desc = d
        c = (C.c_uint64 * 4)()
        self.lib.tm_hostmodel_counts(self.h, C.byref(c, 0), C.byref(c, 8), C.byref(c, 16), C.byref(c, 24))
        self.n_subset, self.n_entries, self.n_keys, self.n_kept = [int(x) for x in c]
        self.extents = np.array(list(d.extents), dtype=np.int32)
        self.to_voxel16 = np.array(list(d.to_voxel), dtype=np.float32)
        self.resolution, self.diameter = float(d.resolution), float(d.diameter)
        self.feat_min = np.array(list(d.feat_min), dtype=np.float32)
        self.feat_max = np.array(list(d.feat_max), dtype=np.float32)

    def save(self, path: str):
        rc = self.lib.tm_hostmodel_save(self.h, C.c_uint32(int(self._view.n)), path.encode())
        if rc != TM_OK:
            raise TmError(rc, self.lib.tm_host_last_error().decode("utf-8", "replace"))

    @classmethod
    def load(cls, path: str, pos, nrm, tgt) -> "HostModel":
        """Reload a blob written by save(); pos/nrm/tgt are the cloud it was built for."""
        self = cls.__new__(cls)
        self.lib = load()
        self._view, self._keep = _view(pos, nrm, tgt)
        self.h = C.c_void_p()
        rc = self.lib.tm_hostmodel_load(path.encode(), C.c_uint32(int(self._view.n)), C.byref(self.h))
        if rc != TM_OK:
            raise TmError(rc, self.lib.tm_host_last_error().decode("utf-8", "replace"))
        self._read_desc()
        return self

    def _arr(self, ptr, n, dtype=np.uint32):
        if not n:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n,)).copy()

    @property
    def voxel(self):
        return self._arr(self.desc.voxel, int(np.prod(self.extents.astype(np.int64))))

    @property
    def keys(self):
        return self._arr(self.desc.keys, 4 * self.n_keys).reshape(-1, 4)

    @property
    def offsets(self):
        return self._arr(self.desc.offsets, self.n_keys + 1)

    @property
    def pairs(self):
        return self._arr(self.desc.pairs, 2 * self.n_kept).reshape(-1, 2)

    @property
    def subset(self):
        return self._arr(self.lib.tm_hostmodel_subset(self.h), self.n_subset)

    def upload(self, ctx) -> "Model":
        m = Model.__new__(Model)
        m.ctx, m.lib, m._keep = ctx, ctx.lib, self._keep
        m.n = int(self._view.n)
        m.diameter, m.resolution = self.diameter, self.resolution
        m.h = C.c_void_p()
        rc = self.lib.tm_model_create(ctx.h, C.byref(self._view), self.h, C.byref(m.h))
        if rc != TM_OK:
            msg = self.lib.tm_last_error() or self.lib.tm_host_last_error()
            raise TmError(rc, msg.decode("utf-8", "replace"))
        return m

    def close(self):
        if self.h:
            self.lib.tm_hostmodel_destroy(self.h)
            self.h = C.c_void_p()


class Model:
    """Resident model: cloud + what model::init produced (include/impl/model.hpp:16-167)."""

    def __init__(self, ctx: Context, pos, nrm, tgt, *, voxel, extents, to_voxel16, resolution,
                 diameter, keys, offsets, pairs, feat_min, feat_max, distance_step_count,
                 angle_step, view: CloudView | None = None):
        self.ctx = ctx
        self.lib = ctx.lib
        if view is None:
            v, self._keep = _view(pos, nrm, tgt)
        else:
            v, self._keep = view, (pos,)
        voxel = np.ascontiguousarray(voxel, dtype=np.uint32)
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 4)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        d = ModelDesc()
        d.voxel = voxel.ctypes.data
        d.extents = (C.c_int32 * 3)(*[int(e) for e in extents])
        d.to_voxel = (C.c_float * 16)(*[float(x) for x in np.asarray(to_voxel16, dtype=np.float32).ravel()])
        d.resolution = float(resolution)
        d.diameter = float(diameter)
        d.keys = keys.ctypes.data if keys.size else None
        d.offsets = offsets.ctypes.data if offsets.size else None
        d.pairs = pairs.ctypes.data if pairs.size else None
        d.n_keys = keys.shape[0]
        d.feat_min = (C.c_float * 4)(*[float(x) for x in feat_min])
        d.feat_max = (C.c_float * 4)(*[float(x) for x in feat_max])
        d.distance_step_count = float(distance_step_count)
        d.angle_step = float(angle_step)
        self.n = int(v.n)
        self.diameter = float(np.float32(diameter))
        self.resolution = float(np.float32(resolution))
        self.h = C.c_void_p()
        _chk(self.lib.tm_model_upload(ctx.h, C.byref(v), C.byref(d), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.tm_model_destroy(self.h)
            self.h = C.c_void_p()

    def probe(self, keys, valid=None, limit: int = 200):
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 4)
        n = keys.shape[0]
        valid = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.uint64)
        _chk(self.lib.tm_probe(self.h, _p(keys), _p(valid), C.c_uint64(n), C.c_uint32(limit),
                               _p(off), None, C.c_uint64(0)))
        total = int(off[n])
        hits = np.zeros((total, 2), dtype=np.uint32)
        if total:
            _chk(self.lib.tm_probe(self.h, _p(keys), _p(valid), C.c_uint64(n), C.c_uint32(limit),
                                   _p(off), _p(hits), C.c_uint64(total)))
        return off, hits


class Scene:
    def __init__(self, ctx: Context, pos, nrm, tgt, tangent_mask, view: CloudView | None = None,
                 sort: bool = False):
        """sort=True: tm_scene_upload_sorted — the device copy is put into Z-curve order on the device;
        self.to_user[d] is the caller's index of device point d and every index exchanged with the
        scene afterwards is a device index."""
        self.ctx = ctx
        self.lib = ctx.lib
        if view is None:
            v, self._keep = _view(pos, nrm, tgt)
        else:
            v, self._keep = view, (pos,)
        tm = None if tangent_mask is None else np.ascontiguousarray(tangent_mask, dtype=np.uint8)
        self.n = int(v.n)
        self.h = C.c_void_p()
        self.to_user = None
        if sort:
            self.to_user = np.zeros(max(self.n, 1), dtype=np.uint32)
            _chk(self.lib.tm_scene_upload_sorted(ctx.h, C.byref(v), _p(tm), _p(self.to_user), C.byref(self.h)))
            self.to_user = self.to_user[:self.n]
        else:
            _chk(self.lib.tm_scene_upload(ctx.h, C.byref(v), _p(tm), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.tm_scene_destroy(self.h)
            self.h = C.c_void_p()

    def set_mask(self, mask):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        _chk(self.lib.tm_scene_set_mask(self.h, _p(m)))

    def features(self, model: Model, pair_i, pair_j, min_df: float, max_df: float):
        pi = np.ascontiguousarray(pair_i, dtype=np.uint32)
        pj = np.ascontiguousarray(pair_j, dtype=np.uint32)
        n = pi.shape[0]
        feats = np.zeros((n, 4), dtype=np.float32)
        keys = np.zeros((n, 4), dtype=np.uint32)
        valid = np.zeros(n, dtype=np.uint8)
        _chk(self.lib.tm_features(self.h, model.h, _p(pi), _p(pj), C.c_uint64(n), C.c_float(min_df),
                                  C.c_float(max_df), _p(feats), _p(keys), _p(valid)))
        return feats, keys, valid

    def hypotheses(self, model: Model, pair_i, pair_j, offsets, hits, force_up: bool = False):
        pi = np.ascontiguousarray(pair_i, dtype=np.uint32)
        pj = np.ascontiguousarray(pair_j, dtype=np.uint32)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        hits = np.ascontiguousarray(hits, dtype=np.uint32).reshape(-1, 2)
        n = pi.shape[0]
        total = int(off[n]) if n else 0
        T = np.zeros((total, 16), dtype=np.float32)
        valid = np.zeros(total, dtype=np.uint8)
        _chk(self.lib.tm_hypotheses(self.h, model.h, _p(pi), _p(pj), C.c_uint64(n), _p(off),
                                    _p(hits), C.c_int(1 if force_up else 0), _p(T), _p(valid)))
        return T, valid

    def ball_subsets(self, centres, radius: float):
        cs = np.ascontiguousarray(centres, dtype=np.uint32)
        n = cs.shape[0]
        off = np.zeros(n + 1, dtype=np.uint64)
        _chk(self.lib.tm_ball_subsets(self.h, _p(cs), C.c_uint32(n), C.c_float(radius), _p(off),
                                      None, C.c_uint64(0)))
        total = int(off[n])
        idx = np.zeros(max(total, 1), dtype=np.int32)
        _chk(self.lib.tm_ball_subsets(self.h, _p(cs), C.c_uint32(n), C.c_float(radius), _p(off),
                                      _p(idx), C.c_uint64(total)))
        return off, idx[:total]

    def score(self, model: Model, T16s, hyp_sub=None, sub_offsets=None, sub_indices=None,
              dist_thres: float = 1.0, accept_prob: float = 0.5, early_out: bool = False,
              want_scores: bool = True):
        T = _f32(T16s, (-1, 16))
        n = T.shape[0]
        counts = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.float64) if want_scores else None
        dropped = np.zeros(n, dtype=np.uint8)
        hs = so = si = None
        n_sub = 0
        if hyp_sub is not None:
            hs = np.ascontiguousarray(hyp_sub, dtype=np.uint32)
            so = np.ascontiguousarray(sub_offsets, dtype=np.uint64)
            si = np.ascontiguousarray(sub_indices, dtype=np.int32)
            if si.size == 0:
                si = np.zeros(1, dtype=np.int32)
            n_sub = so.shape[0] - 1
        _chk(self.lib.tm_score(self.h, model.h, _p(T), C.c_uint64(n), _p(hs), _p(so), _p(si),
                               C.c_uint32(n_sub), C.c_float(dist_thres), C.c_float(accept_prob),
                               C.c_int(int(early_out)), _p(counts), _p(scores), _p(dropped)))
        return counts, scores, dropped

    def correspondences(self, model: Model, T16, dist_thres: float):
        T = _f32(T16, (16,))
        sc = np.zeros(max(self.n, 1), dtype=np.uint32)
        mc = np.zeros(max(self.n, 1), dtype=np.uint32)
        n = C.c_uint32()
        score = C.c_double()
        _chk(self.lib.tm_correspondences(self.h, model.h, _p(T), C.c_float(dist_thres), _p(sc),
                                         _p(mc), C.byref(n), C.byref(score)))
        return sc[:n.value].copy(), mc[:n.value].copy(), float(score.value)

    def correspondences_batch(self, model: Model, T16s, dist_thres: float):
        """tm_correspondences_batch: (offsets, scene_corrs, model_corrs, scores) of n transforms, sized then filled."""
        T = _f32(T16s, (-1, 16))
        n = T.shape[0]
        off = np.zeros(n + 1, dtype=np.uint64)
        scores = np.zeros(n, dtype=np.float64)
        _chk(self.lib.tm_correspondences_batch(self.h, model.h, _p(T), C.c_uint32(n), C.c_float(dist_thres), _p(off),
                                               None, None, C.c_uint64(0), _p(scores)))
        tot = int(off[-1])
        sc = np.zeros(max(tot, 1), dtype=np.uint32)
        mc = np.zeros(max(tot, 1), dtype=np.uint32)
        off2 = np.zeros(n + 1, dtype=np.uint64)
        _chk(self.lib.tm_correspondences_batch(self.h, model.h, _p(T), C.c_uint32(n), C.c_float(dist_thres), _p(off2),
                                               _p(sc), _p(mc), C.c_uint64(tot), _p(scores)))
        assert np.array_equal(off, off2)
        return off, sc[:tot], mc[:tot], scores

    def icp(self, model: Model, T16s, max_iterations: int, dist_thres: float):
        T = _f32(T16s, (-1, 16))
        n = T.shape[0]
        out = np.zeros_like(T)
        counts = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.float64)
        iters = np.zeros(n, dtype=np.uint32)
        _chk(self.lib.tm_icp(self.h, model.h, _p(T), C.c_uint32(n), C.c_uint32(max_iterations),
                             C.c_float(dist_thres), _p(out), _p(counts), _p(scores), _p(iters)))
        return out, counts, scores, iters


    def knn(self, query_idx, k: int):
        q = np.ascontiguousarray(query_idx, dtype=np.uint32)
        idx = np.zeros((max(q.size, 1), k), dtype=np.int32)
        d2 = np.zeros((max(q.size, 1), k), dtype=np.float32)
        _chk(self.lib.tm_scene_knn(self.h, _p(q), C.c_uint32(q.size), C.c_uint32(k), _p(idx), _p(d2)))
        return idx[:q.size], d2[:q.size]

    def curvature(self, query_idx, k: int):
        q = np.ascontiguousarray(query_idx, dtype=np.uint32)
        mn = np.zeros(max(q.size, 1), dtype=np.float32)
        mx = np.zeros(max(q.size, 1), dtype=np.float32)
        cov = np.zeros((max(q.size, 1), 9), dtype=np.float32)
        _chk(self.lib.tm_scene_curvature(self.h, _p(q), C.c_uint32(q.size), C.c_uint32(k), _p(mn), _p(mx), _p(cov)))
        return mn[:q.size], mx[:q.size], cov[:q.size].reshape(-1, 3, 3)

    def compute_tangent_mask(self, k: int = 30, ratio: float = 0.2, apply: bool = True):
        mask = np.zeros(max(self.n, 1), dtype=np.uint8)
        cnt = C.c_uint32()
        _chk(self.lib.tm_scene_tangent_mask(self.h, C.c_uint32(k), C.c_float(ratio), _p(mask), C.c_int(int(apply)),
                                            C.byref(cnt)))
        return mask[:self.n], int(cnt.value)

    def icp_pose_sharded(self, model: Model, T16s, max_iterations: int, dist_thres: float, rank: int = 0,
                         world: int = 1, comm=None):
        """tm_icp_pose_sharded: this rank refines its slice of the poses against the whole resident scene; with a
        communicator all n results come back on every rank, without one only the slice [n*rank/world, ...)."""
        T = _f32(T16s, (-1, 16))
        n = T.shape[0]
        out = np.zeros_like(T)
        counts = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.float64)
        iters = np.zeros(n, dtype=np.uint32)
        _chk(self.lib.tm_icp_pose_sharded(self.h, model.h, comm.h if comm is not None else None,
                                          C.c_uint32(rank), C.c_uint32(world), _p(T), C.c_uint32(n),
                                          C.c_uint32(max_iterations), C.c_float(dist_thres), _p(out), _p(counts),
                                          _p(scores), _p(iters)))
        return out, counts, scores, iters

    def icp_sharded(self, model: Model, T16s, max_iterations: int, dist_thres: float, pt_begin: int,
                    pt_end: int, n_scene_total: int, comm=None, emulate_parts: int = 1):
        """tm_icp_sharded: this process accumulates scene points [pt_begin, pt_end)."""
        T = _f32(T16s, (-1, 16))
        n = T.shape[0]
        out = np.zeros_like(T)
        counts = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.float64)
        iters = np.zeros(n, dtype=np.uint32)
        _chk(self.lib.tm_icp_sharded(self.h, model.h, comm.h if comm is not None else None, _p(T),
                                     C.c_uint32(n), C.c_uint32(max_iterations), C.c_float(dist_thres),
                                     C.c_uint32(pt_begin), C.c_uint32(pt_end), C.c_uint64(n_scene_total),
                                     C.c_uint32(emulate_parts), _p(out), _p(counts), _p(scores), _p(iters)))
        return out, counts, scores, iters


class Query:
    """Resident recorded-list search (tm_query_*)."""

    def __init__(self, scene: Scene, model: Model, *, min_df=0.2, max_df=1.0, force_up=False,
                 query_limit=200, dist_thres=1.0, accept_prob=0.5, early_out=False, icp_top_k=0,
                 max_icp_iterations=0, max_hypotheses=0, hyp_limit=0):
        self.lib = scene.lib
        self.scene, self.model = scene, model
        p = QueryParams(min_df, max_df, 1 if force_up else 0, query_limit, dist_thres, accept_prob,
                        int(early_out), icp_top_k, max_icp_iterations, max_hypotheses,
                        hyp_limit)
        self.params = p
        self.h = C.c_void_p()
        _chk(self.lib.tm_query_create(scene.h, model.h, C.byref(p), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.tm_query_destroy(self.h)
            self.h = C.c_void_p()

    def set_pairs(self, outer, pair_outer, pair_j):
        o = np.ascontiguousarray(outer, dtype=np.uint32)
        po = np.ascontiguousarray(pair_outer, dtype=np.uint32)
        pj = np.ascontiguousarray(pair_j, dtype=np.uint32)
        _chk(self.lib.tm_query_set_pairs(self.h, _p(o), C.c_uint32(o.shape[0]), _p(po), _p(pj),
                                         C.c_uint64(po.shape[0])))

    def set_shard(self, rank: int, world: int):
        _chk(self.lib.tm_query_set_shard(self.h, C.c_uint32(rank), C.c_uint32(world)))

    def set_balance(self, by_tests: bool, comm=None):
        """Shards of equal hypothesis-point tests instead of equal hypothesis counts (tm_query_set_balance)."""
        _chk(self.lib.tm_query_set_balance(self.h, C.c_int(1 if by_tests else 0), comm.h if comm is not None else None))

    def run(self):
        _chk(self.lib.tm_query_run(self.h))

    def frontend_ms(self) -> float:
        ms = C.c_float()
        _chk(self.lib.tm_query_frontend_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def early_walked(self) -> int:
        """early_out = 2: hypotheses of the last run that were walked one by one (tm_query_early_walked)."""
        n = C.c_uint32()
        _chk(self.lib.tm_query_early_walked(self.h, C.byref(n)))
        return int(n.value)

    def result(self) -> QueryResult:
        r = QueryResult()
        _chk(self.lib.tm_query_result_get(self.h, C.byref(r)))
        return r

    def score_kernel_ms(self) -> float:
        ms = C.c_float()
        _chk(self.lib.tm_query_score_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def best_key_device_ptr(self) -> int:
        return int(self.lib.tm_query_best_key_device(self.h))

    def set_global_best(self, key: int):
        _chk(self.lib.tm_query_set_global_best(self.h, C.c_uint64(key)))

    def download(self):
        r = self.result()
        n = int(r.n_scored)
        counts = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.float64)
        T = np.zeros((n, 16), dtype=np.float32)
        valid = np.zeros(n, dtype=np.uint8)
        hyp_pair = np.zeros(n, dtype=np.uint32)
        dropped = np.zeros(n, dtype=np.uint8)
        if n:
            _chk(self.lib.tm_query_download(self.h, C.c_uint64(n), _p(counts), _p(scores), _p(T),
                                            _p(valid), _p(hyp_pair), _p(dropped)))
        return dict(counts=counts, scores=scores, T=T, valid=valid, hyp_pair=hyp_pair,
                    dropped=dropped, result=r)

    def download_counts(self, out=None):
        """result + per-hypothesis inlier counts only (the e2e read-back).  out: a caller-owned uint32 buffer (e.g.
        pinned host memory) of at least n_scored entries; a view of its first n_scored entries is returned."""
        r = self.result()
        n = int(r.n_scored)
        if out is None:
            counts = np.zeros(n, dtype=np.uint32)
        else:
            assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size >= n
            counts = out[:n]
        if n:
            _chk(self.lib.tm_query_download(self.h, C.c_uint64(n), _p(counts), None, None, None,
                                            None, None))
        return counts, r

    def icp_results(self):
        k = int(self.params.icp_top_k)
        ids = np.zeros(k, dtype=np.uint32)
        T = np.zeros((k, 16), dtype=np.float32)
        counts = np.zeros(k, dtype=np.uint32)
        scores = np.zeros(k, dtype=np.float64)
        iters = np.zeros(k, dtype=np.uint32)
        _chk(self.lib.tm_query_icp_results(self.h, _p(ids), _p(T), _p(counts), _p(scores), _p(iters)))
        return ids, T, counts, scores, iters


class Comm:
    def __init__(self, ctx: Context, unique_id: bytes, rank: int, world: int):
        self.lib = ctx.lib
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self.h = C.c_void_p()
        _chk(self.lib.tm_comm_create(ctx.h, buf, C.c_int(rank), C.c_int(world), C.byref(self.h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _chk(load().tm_nccl_unique_id(buf))
        return bytes(buf)

    def allreduce_best_many(self, queries):
        arr = (C.c_void_p * len(queries))(*[q.h.value for q in queries])
        _chk(self.lib.tm_queries_allreduce_best(arr, C.c_uint32(len(queries)), self.h))

    def allreduce_best(self, q: Query):
        _chk(self.lib.tm_query_allreduce_best(q.h, self.h))

    def close(self):
        if self.h:
            self.lib.tm_comm_destroy(self.h)
            self.h = C.c_void_p()


# ---- host mirrors of the device-side sharding / key packing (k_query.cu, k_score.cu) ----
def walk_stride(n: int) -> int:
    """tm_walk_stride: stride s of the evenly sampling walk p -> (p * s) mod n (early_out = 2)."""
    lib = load()
    lib.tm_walk_stride.restype = C.c_uint32
    return int(lib.tm_walk_stride(C.c_uint32(n)))


def early_level_begin(n: int, level: int) -> int:
    """tm_early_level_begin: first walk position of checkpoint range `level` of an n-element subset."""
    lib = load()
    lib.tm_early_level_begin.restype = C.c_uint32
    return int(lib.tm_early_level_begin(C.c_uint32(n), C.c_int(level)))


def walk_order(n: int) -> np.ndarray:
    """The walk as an index array: element visited at position p."""
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    return ((np.arange(n, dtype=np.uint64) * np.uint64(walk_stride(n))) % np.uint64(n)).astype(np.int64)


def shard_range(n_hyp: int, rank: int, world: int, hyp_limit: int = 0):
    """[begin, end) of the global hypothesis list scored by `rank` (shard_range_kernel)."""
    H = min(n_hyp, hyp_limit) if hyp_limit else n_hyp
    per = (H + world - 1) // world
    hb = min(rank * per, H)
    return hb, min(hb + per, H)


def balanced_bounds(hyp_per_outer, sizes, world: int):
    """Host mirror of balance_bounds_kernel (tm_query_set_balance): cuts the global hypothesis list — outer sample o
    owns hyp_per_outer[o] consecutive hypotheses, each costing sizes[o] point tests — into `world` contiguous ranges
    of (nearly) equal tests.  Returns world + 1 hypothesis indices."""
    nh = np.asarray(hyp_per_outer, dtype=np.uint64)
    sz = np.asarray(sizes, dtype=np.uint64)
    hb = np.concatenate([[0], np.cumsum(nh)]).astype(np.uint64)
    cum = np.concatenate([[0], np.cumsum(nh * sz)]).astype(np.uint64)
    H, total = int(hb[-1]), int(cum[-1])
    out = [0]
    for r in range(1, world):
        if total == 0:
            out.append(min(H, (H + world - 1) // world * r))
            continue
        target = total * r // world
        a = int(np.searchsorted(cum, np.uint64(target), side="right")) - 1  # last o with cum[o] <= target
        a = min(a, nh.size - 1)
        s = int(sz[a])
        out.append(min(int(hb[a + 1]), int(hb[a]) + (target - int(cum[a])) // s) if s else int(hb[a]))
    out.append(H)
    return out


def pose_range(n_poses: int, rank: int, world: int):
    """Poses refined by `rank` when an ICP batch is sharded over the poses (tm_icp_pose_sharded)."""
    return (n_poses * rank) // world, (n_poses * (rank + 1)) // world


def point_range(n_points: int, rank: int, world: int):
    """Scene points of `rank` when an ICP pass is sharded over the scene (tm_icp_sharded)."""
    return (n_points * rank) // world, (n_points * (rank + 1)) // world


def pack_key(inliers: int, global_id: int) -> int:
    """(inliers << 32) | (0xFFFFFFFF - id): max-reduce picks most inliers, lowest id on ties."""
    return (int(inliers) << 32) | (0xFFFFFFFF - int(global_id))


def unpack_key(key: int):
    return int(key) >> 32, 0xFFFFFFFF - (int(key) & 0xFFFFFFFF)
