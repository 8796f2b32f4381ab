"""CPU tests of the product's host side: C-ABI surface, model::init restatement (host C++)
against the oracle, loud failure without a GPU, synthetic generators."""
import os
import re

import numpy as np
import pytest

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol(built):
    from triplet_match_b200 import capi
    lib = capi.load()
    declared = set()
    for hdr in ("tm_b200.h", "tm_b200_host.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 45
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(capi.EXPORTS + capi.HOST_EXPORTS) <= declared
    assert b"sm_100a" in lib.tm_version()


@pytest.mark.skipif(not _no_gpu(), reason="box has a GPU")
def test_no_cpu_fallback(built):
    from triplet_match_b200 import capi
    with pytest.raises(capi.TmError) as e:
        capi.Context(0)
    assert e.value.code == capi.TM_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("name", ["plane_small", "cylinder_small", "freeform_small"])
def test_host_model_init_matches_oracle(built, name):
    from triplet_match_b200 import capi
    m, s, om, osc, rec = common.config(name)
    hm = capi.HostModel(None, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **common.DP, **common.SP)
    k, o, p = om.table(200)
    assert np.float32(hm.resolution) == np.float32(om.resolution)
    assert np.float32(hm.diameter) == np.float32(om.diameter)
    assert np.array_equal(hm.extents, om.extents)
    assert np.array_equal(hm.to_voxel16.view(np.uint32), om.to_voxel16.view(np.uint32))
    assert np.array_equal(hm.voxel, om.voxel)
    assert np.array_equal(hm.subset, om.subset)
    assert np.array_equal(hm.feat_min.view(np.uint32), om.feat_min.view(np.uint32))
    assert np.array_equal(hm.feat_max.view(np.uint32), om.feat_max.view(np.uint32))
    assert hm.n_entries == om.n_entries and hm.n_keys == om.n_keys
    assert np.array_equal(hm.keys, k) and np.array_equal(hm.offsets, o) and np.array_equal(hm.pairs, p)
    hm.close()


def test_host_resolution_matches_bruteforce(built):
    from oracle import pyoracle as po
    from triplet_match_b200 import capi, synth
    for seed in (1, 2, 3):
        c = synth.freeform_model(seed=seed, n_points=700, radius=0.2)
        exp = po.load().orc_resolution(po._p(po._f32(c.pos)), po.C.c_uint32(c.n))
        assert np.float32(capi.host_resolution(c.pos)) == np.float32(exp)


def test_uninitialized_model_error(built):
    from triplet_match_b200 import capi
    lib = capi.load()
    out = capi.C.c_void_p()
    v = capi.CloudView(None, None, None, 3, 0)
    rc = lib.tm_model_create(None, capi.C.byref(v), None, capi.C.byref(out))
    assert rc == capi.TM_ERR_UNINITIALIZED
    assert b"uninitialized model" in lib.tm_host_last_error()  # include/impl/model.hpp:171-173


def test_synth_is_seed_fixed():
    from triplet_match_b200 import synth
    a = synth.plane_model(seed=2, size=0.2)
    b = synth.plane_model(seed=2, size=0.2)
    c = synth.plane_model(seed=3, size=0.2)
    assert np.array_equal(a.pos, b.pos) and not np.array_equal(a.pos, c.pos)
    assert a.pos.dtype == np.float32 and a.tangent_mask.dtype == np.uint8
    on = a.tangent_mask.astype(bool)
    assert np.allclose(np.linalg.norm(a.tgt[on], axis=1), 1, atol=1e-6) and not a.tgt[~on].any()
    s = synth.make_scene(seed=4, model=a, n_points=5000, n_copies=2, extent=0.8)
    assert s.n == 5000 and len(s.poses) == 2
    p = synth.morton_order(s.pos)
    assert sorted(p.tolist()) == list(range(s.n))
    rec = synth.record_pairs(1, s, 0.3, 4, 10)
    assert np.all(np.diff(rec.pair_outer.astype(np.int64)) >= 0)
    assert np.all(s.tangent_mask[rec.pair_j] == 1) and np.all(rec.pair_j != rec.pair_i)


def test_model_blob_round_trip_and_corruption(built, tmp_path):
    """tm_hostmodel_save / tm_hostmodel_load: identical tables after a reload; magic, version, cloud
    size, checksum and index checks reject damaged files (CPU only: ctx = None build)."""
    from triplet_match_b200 import capi
    m, s, om, osc, rec = common.config("cylinder_small")
    hm = capi.HostModel(None, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, **common.DP, **common.SP)
    path = str(tmp_path / "model.tmb")
    hm.save(path)
    h2 = capi.HostModel.load(path, m.pos, m.nrm, m.tgt)
    assert (h2.n_subset, h2.n_entries, h2.n_keys, h2.n_kept) == (hm.n_subset, hm.n_entries, hm.n_keys, hm.n_kept)
    for a in ("voxel", "keys", "offsets", "pairs", "subset", "extents", "to_voxel16", "feat_min", "feat_max"):
        assert np.array_equal(getattr(h2, a), getattr(hm, a)), a
    assert h2.resolution == hm.resolution and h2.diameter == hm.diameter
    h2.close()
    raw = bytearray(open(path, "rb").read())
    def expect_fail(data, n_pts=m.n, what=""):
        p = str(tmp_path / "bad.tmb")
        open(p, "wb").write(bytes(data))
        with pytest.raises(capi.TmError) as e:
            capi.HostModel.load(p, m.pos[:n_pts], m.nrm[:n_pts], m.tgt[:n_pts])
        assert what in str(e.value)
    flipped = bytearray(raw); flipped[len(raw) // 2] ^= 0x40
    expect_fail(flipped, what="checksum")
    expect_fail(raw[: len(raw) // 3], what="truncated")
    expect_fail(b"NOTABLOB" + bytes(raw[8:]), what="not a model blob")
    v2 = bytearray(raw); v2[8] = 9
    expect_fail(v2, what="version")
    expect_fail(raw, n_pts=m.n - 1, what="points")
    with pytest.raises(capi.TmError):
        capi.HostModel.load(str(tmp_path / "missing.tmb"), m.pos, m.nrm, m.tgt)
    hm.close()


def test_walk_stride_is_a_permutation_with_even_prefixes(built):
    """tm_walk_stride (early_out = 2): coprime stride near n / golden ratio; every prefix of the walk spreads
    over the whole index range."""
    import math
    from triplet_match_b200 import capi
    for n in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 30, 64, 97, 100, 1000, 1024, 56789, 1 << 20, 999983, 2 ** 32 - 1):
        st = capi.walk_stride(n)
        if n <= 2:
            assert st == 1
            continue
        assert 1 <= st < n and math.gcd(st, n) == 1, (n, st)
        if n >= 30:
            assert abs(st / n - 0.6180339887) < 0.2, (n, st)
    # known answers: the rule is part of the C-ABI contract (include/tm_b200.h), host and device share it
    kat = {3: 1, 4: 3, 5: 3, 6: 5, 10: 7, 30: 19, 64: 39, 100: 61, 1000: 619, 1024: 633, 56789: 35097,
           1 << 20: 648055, 999983: 618023, 2 ** 32 - 1: 2654435768}
    assert {n: capi.walk_stride(n) for n in kat} == kat
    for n in (1, 2, 5, 64, 1000, 56789):
        w = capi.walk_order(n)
        assert np.array_equal(np.sort(w), np.arange(n))
    w = capi.walk_order(56789)
    head = np.sort(w[: 56789 // 20])  # the first 5 %: no gap much larger than the mean spacing
    gaps = np.diff(np.concatenate([[0], head, [56789]]))
    assert gaps.max() < 6 * 20  # three-distance theorem: at most three gap lengths, the largest a few means
