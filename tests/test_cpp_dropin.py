"""The drop-in C++ API (include/triplet_match/*): host-only checks on CPU, and a full
model::init + scene::find_all_parallel run on the GPU against known ground-truth poses."""
import os
import subprocess

import numpy as np
import pytest

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")
LIBDIR = os.path.join(ROOT, "triplet_match_b200")


@pytest.fixture(scope="module")
def exe(built, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "test_dropin")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                           SRC, "-o", out, "-L" + LIBDIR, "-ltriplet_match_b200", "-Wl,-rpath," + LIBDIR])
    return out


def test_cpp_host_api(exe):
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "cpu checks ok" in r.stdout


def _write(cloud, path):
    rec = np.concatenate([cloud.pos, cloud.nrm, cloud.tgt], axis=1).astype(np.float32)
    with open(path, "wb") as f:
        f.write(np.uint32(cloud.n).tobytes())
        f.write(np.ascontiguousarray(rec).tobytes())


def _noisy_normals(cloud, seed, sigma=0.05):
    from triplet_match_b200 import synth
    rng = np.random.default_rng(seed)
    nrm = cloud.nrm.astype(np.float64) + sigma * rng.standard_normal(cloud.nrm.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return synth.Cloud(cloud.pos, nrm.astype(np.float32), cloud.tgt, cloud.tangent_mask, cloud.poses)


@pytest.mark.gpu
@pytest.mark.parametrize("curv,shuffled,knobs", [(False, False, {}), (True, False, {}), (False, True, {}),
                                                 (False, False, {"TM_DROPIN_BATCH": "4000"}),   # rounds split into several queries
                                                 (False, False, {"TM_DROPIN_ICP_ITERS": "0"}),  # icp_ returns the match unchanged
                                                 (False, False, {"TM_DROPIN_EARLY_DROP": "1"}),  # project_(early_out = true) over the even walk
                                                 (False, False, {"TM_DROPIN_EARLY_DROP": "1", "TM_DROPIN_BATCH": "4000"})])
def test_cpp_find_all_parallel(exe, tmp_path, curv, shuffled, knobs):
    """curv=True: raw clouds with estimated (noisy) normals, tangent masks from the GPU 30-NN
    curvature criterion on both model and scene, as the reference does with PCL."""
    from triplet_match_b200 import synth
    m, s, om, osc, rec = common.config("freeform_small")
    res = om.resolution
    if curv:  # a model with genuine creases: the criterion keeps its edge points and rejects the floor's fake tangents
        m = synth.pyramid_model(seed=7, size=0.4, height=0.16, res=0.005)
        s = synth.make_scene(seed=8, model=m, n_points=60000, n_copies=3, extent=1.5, flat_copies=False, res=0.005)
        s = s.take(synth.morton_order(s.pos))
        res = 0.005
    if shuffled:  # the caller's cloud in random order: the drop-in sorts its device copy itself
        s = s.take(synth.shuffle_perm(5, 1, s.n))
    mp, sp, op = str(tmp_path / "m.bin"), str(tmp_path / "s.bin"), str(tmp_path / "o.txt")
    _write(m, mp)
    _write(s, sp)
    extra = ["curv" if curv else "nocurv"] + ([str(tmp_path / "model.tmb")] if not curv and not shuffled and not knobs else [])
    dump = str(tmp_path / "rounds.txt")
    r = subprocess.run([exe, "find", mp, sp, op] + extra, capture_output=True, text=True,
                       env=dict(os.environ, TM_DROPIN_DUMP=dump, **knobs))
    assert r.returncode == 0, r.stderr + r.stdout
    if not curv:
        _check_rounds_against_oracle(dump, s if not shuffled else s, om)
    lines = open(op).read().strip().split("\n")
    n = int(lines[0])
    assert n >= 1, r.stdout
    mpts = m.pos.astype(np.float64)
    found = set()
    for ln in lines[1:1 + n]:
        v = [float(x) for x in ln.split()]
        T = np.array(v[2:18]).reshape(4, 4).T  # column-major, model -> scene
        placed = mpts @ T[:3, :3].T + T[:3, 3]
        errs = [np.abs(placed - (mpts @ P[:3, :3].T + P[:3, 3])).max() for P in s.poses]
        k = int(np.argmin(errs))
        tol = (8 if knobs.get("TM_DROPIN_ICP_ITERS") == "0" else 3) * res  # unrefined poses are coarser
        assert errs[k] < tol, (errs, r.stdout)  # every reported instance is a real one
        assert int(v[0]) >= 0.5 * m.n
        found.add(k)
        # correspondences are reported in the caller's index space: the matched scene points lie on the instance
        corr = [int(x) for x in ln.split()[18:]]
        if corr:
            dmin = np.abs(s.pos[corr].astype(np.float64)[:, None, :] - placed[None, ::7, :]).sum(-1).min(1)
            assert np.median(dmin) < 6 * res
    assert len(found) == n  # no instance is reported twice (overlap-free acceptance)


def _check_rounds_against_oracle(dump, s, om):
    """Every find_parallel round of the drop-in dumps its recorded (p1, p2) list in the caller's indices and the number
    of hypotheses the device generated from it; the oracle must generate as many from the same list, and the list must
    honour the reference's inner budget (scene.hpp:277-305, 350-352): only pairs that pass window / collinearity /
    valid are recorded, and per outer sample at most inner_bound + 1 of them."""
    from oracle import pyoracle as po
    osc = po.OScene(s)
    rounds, cur = [], None
    for ln in open(dump).read().strip().split("\n"):
        t = ln.split()
        if t[0] == "round":
            cur = dict(n_outer=int(t[1]), n_pairs=int(t[2]), n_hyp=int(t[3]), pairs=[])
            rounds.append(cur)
        else:
            cur["pairs"].append((int(t[0]), int(t[1])))
    assert rounds and any(r["n_hyp"] > 0 for r in rounds)
    for r in rounds:
        assert len(r["pairs"]) == r["n_pairs"]
        if not r["pairs"]:
            assert r["n_hyp"] == 0
            continue
        pi = np.array([p[0] for p in r["pairs"]], dtype=np.uint32)
        pj = np.array([p[1] for p in r["pairs"]], dtype=np.uint32)
        f, k, v = osc.pair_features(om, pi, pj)
        assert v.all()  # the host-side filters of the drop-in are the oracle's: every recorded pair is a valid sample
        T, hp, *_ = osc.hypotheses(om, pi, pj)
        assert T.shape[0] == r["n_hyp"]
        # inner budget: <= inner_bound + 1 valid samples per outer sample
        n_model_all = om.n
        for o in np.unique(pi):
            nn = osc.ball_subset(int(o), om.diameter).size
            bound = int(-np.log(1.0 - 0.999) / (n_model_all / nn))
            bound = min(max(bound, 10), nn)
            assert int((pi == o).sum()) <= bound + 1


# ---- PCD I/O + the CLI ------------------------------------------------------
def _pcd_header(fields, sizes, types, n, data):
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS " + " ".join(fields) + "\nSIZE " +
            " ".join(map(str, sizes)) + "\nTYPE " + " ".join(types) + "\nCOUNT " + " ".join(["1"] * len(fields)) +
            f"\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {data}\n").encode()


def write_pcd(cloud, path, binary=True):
    """x y z normal_x normal_y normal_z rgba radius confidence curvature (tangent in the last three)."""
    n = cloud.n
    rec = np.zeros((n, 10), dtype=np.float32)
    rec[:, 0:3], rec[:, 3:6], rec[:, 7:10] = cloud.pos, cloud.nrm, cloud.tgt
    rec[:, 6] = (np.arange(n, dtype=np.uint32) * 2654435761 & 0xFFFFFF).astype(np.uint32).view(np.float32)
    fields = ["x", "y", "z", "normal_x", "normal_y", "normal_z", "rgba", "radius", "confidence", "curvature"]
    with open(path, "wb") as f:
        f.write(_pcd_header(fields, [4] * 10, ["F"] * 6 + ["U"] + ["F"] * 3, n, "binary" if binary else "ascii"))
        if binary:
            f.write(rec.tobytes())
        else:
            rg = rec[:, 6].view(np.uint32)
            for i in range(n):
                v = rec[i]
                f.write((" ".join(repr(float(x)) for x in v[:6]) + f" {int(rg[i])} " +
                         " ".join(repr(float(x)) for x in v[7:]) + "\n").encode())
    return rec


def _surfels(path, n):
    raw = np.fromfile(path, dtype=np.float32).reshape(n, 12)
    return raw


def _lzf_compress(data: bytes, literal_only=False) -> bytes:
    """A small LZF encoder for the tests (format of liblzf / PCL's binary_compressed): greedy matches found
    through a 3-byte hash of the last position, literal runs of at most 32 bytes."""
    out = bytearray()
    lit = bytearray()

    def flush():
        for k in range(0, len(lit), 32):
            chunk = lit[k:k + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lit.clear()

    last = {}
    i, n = 0, len(data)
    while i < n:
        best_len = 0
        if not literal_only and i + 3 <= n:
            key = data[i:i + 3]
            j = last.get(key)
            last[key] = i
            if j is not None and 0 < i - j <= 8192:
                m = 0
                while i + m < n and m < 264 and data[j + m] == data[i + m]:  # overlapping matches are legal
                    m += 1
                if m >= 3:
                    best_len, dist = m, i - j - 1
        if best_len:
            flush()
            ln = best_len - 2
            if ln < 7:
                out.append((ln << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(ln - 7)
            out.append(dist & 0xFF)
            i += best_len
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def test_pcd_binary_compressed(exe, tmp_path):
    """PCL's DATA binary_compressed: u32 compressed size, u32 raw size, LZF stream of the field-major records."""
    m, *_ = common.config("plane_small")
    src_b, out_b = str(tmp_path / "plain.pcd"), str(tmp_path / "plain.bin")
    rec = write_pcd(m, src_b, binary=True)
    assert subprocess.run([exe, "pcd", src_b, out_b], capture_output=True).returncode == 0
    fields = ["x", "y", "z", "normal_x", "normal_y", "normal_z", "rgba", "radius", "confidence", "curvature"]
    raw = np.ascontiguousarray(rec.T).tobytes()  # field-major
    for literal_only in (False, True):
        packed = _lzf_compress(raw, literal_only)
        if not literal_only:
            assert len(packed) < len(raw)  # the plane's constant normals compress: back-references are exercised
        p = str(tmp_path / "c.pcd")
        with open(p, "wb") as f:
            f.write(_pcd_header(fields, [4] * 10, ["F"] * 6 + ["U"] + ["F"] * 3, m.n, "binary_compressed"))
            f.write(np.uint32(len(packed)).tobytes() + np.uint32(len(raw)).tobytes() + packed)
        out = str(tmp_path / "c.bin")
        r = subprocess.run([exe, "pcd", p, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.fromfile(out, np.uint32), np.fromfile(out_b, np.uint32))
    # a run of one byte: the back-reference overlaps its own output
    ones = np.full((64, 3), 1.0, np.float32)
    raw = np.ascontiguousarray(ones.T).tobytes()
    packed = _lzf_compress(raw)
    assert len(packed) < 40
    p = str(tmp_path / "o.pcd")
    with open(p, "wb") as f:
        f.write(_pcd_header(["x", "y", "z"], [4] * 3, ["F"] * 3, 64, "binary_compressed"))
        f.write(np.uint32(len(packed)).tobytes() + np.uint32(len(raw)).tobytes() + packed)
    out = str(tmp_path / "o.bin")
    assert subprocess.run([exe, "pcd", p, out], capture_output=True).returncode == 0
    assert np.array_equal(_surfels(out, 64)[:, 0:3], ones)
    # corrupt streams and inconsistent sizes are errors, not crashes
    for blob in (np.uint32(len(packed)).tobytes() + np.uint32(len(raw)).tobytes() + packed[:-3],          # truncated
                 np.uint32(len(packed)).tobytes() + np.uint32(len(raw) + 4).tobytes() + packed,            # wrong raw size
                 np.uint32(4).tobytes() + np.uint32(len(raw)).tobytes() + bytes([0xE0, 0xFF, 0xFF, 0x00]),  # reference before start
                 np.uint32(len(packed) - 1).tobytes() + np.uint32(len(raw)).tobytes() + packed[:-1]):      # short output
        with open(p, "wb") as f:
            f.write(_pcd_header(["x", "y", "z"], [4] * 3, ["F"] * 3, 64, "binary_compressed") + blob)
        assert subprocess.run([exe, "pcd", p, out], capture_output=True).returncode == 3


def test_pcd_reader_and_writer(exe, tmp_path):
    m, *_ = common.config("plane_small")
    for binary in (True, False):
        src, out, re_a = str(tmp_path / "a.pcd"), str(tmp_path / "a.bin"), str(tmp_path / "b.pcd")
        rec = write_pcd(m, src, binary)
        r = subprocess.run([exe, "pcd", src, out, re_a, "ascii" if binary else "binary"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        s = _surfels(out, m.n)
        assert np.array_equal(s[:, 0:3].view(np.uint32), rec[:, 0:3].view(np.uint32))
        assert (s[:, 3] == 1).all()
        assert np.array_equal(s[:, 4:7].view(np.uint32), rec[:, 3:6].view(np.uint32))
        assert np.array_equal(s[:, 8].view(np.uint32), rec[:, 6].view(np.uint32))      # rgba bits
        assert np.array_equal(s[:, 9:12].view(np.uint32), rec[:, 7:10].view(np.uint32))  # tangent overlay
        # what the writer produced reads back identically (other encoding)
        out2 = str(tmp_path / "b.bin")
        assert subprocess.run([exe, "pcd", re_a, out2], capture_output=True).returncode == 0
        assert np.array_equal(np.fromfile(out2, np.uint32), np.fromfile(out, np.uint32))
    # other layouts: reordered / extra / double / missing fields, tangent_* aliases
    n = 5
    pts = np.arange(n * 3, dtype=np.float64).reshape(n, 3) * 0.25
    p = str(tmp_path / "c.pcd")
    with open(p, "wb") as f:
        f.write(_pcd_header(["intensity", "z", "y", "x", "tangent_x"], [2, 8, 8, 8, 4], ["U", "F", "F", "F", "F"], n, "binary"))
        for i in range(n):
            f.write(np.uint16(i).tobytes() + pts[i, ::-1].astype(np.float64).tobytes() + np.float32(0.5 + i).tobytes())
    out = str(tmp_path / "c.bin")
    assert subprocess.run([exe, "pcd", p, out], capture_output=True).returncode == 0
    s = _surfels(out, n)
    assert np.array_equal(s[:, 0:3], pts.astype(np.float32)) and np.array_equal(s[:, 9], np.arange(n, dtype=np.float32) + 0.5)
    assert (s[:, 4:7] == 0).all() and (s[:, 10:12] == 0).all()
    # errors: compressed data, truncated file, missing file
    bad = str(tmp_path / "d.pcd")
    with open(bad, "wb") as f:
        f.write(_pcd_header(["x", "y", "z"], [4] * 3, ["F"] * 3, 4, "binary_compressed"))
    assert subprocess.run([exe, "pcd", bad, out], capture_output=True).returncode == 3
    with open(bad, "wb") as f:
        f.write(_pcd_header(["x", "y", "z"], [4] * 3, ["F"] * 3, 4, "binary") + b"\0" * 20)
    assert subprocess.run([exe, "pcd", bad, out], capture_output=True).returncode == 3
    assert subprocess.run([exe, "pcd", str(tmp_path / "nope.pcd"), out], capture_output=True).returncode == 3
    # hostile headers: zero / negative SIZE, zero COUNT, POINTS larger than the file — an error, never a crash
    def hdr(sizes, counts, npts, data="binary"):
        return ("VERSION 0.7\nFIELDS x y z\nSIZE " + " ".join(map(str, sizes)) + "\nTYPE F F F\nCOUNT " +
                " ".join(map(str, counts)) + f"\nWIDTH {npts}\nHEIGHT 1\nPOINTS {npts}\nDATA {data}\n").encode()
    for h in (hdr([0, 4, 4], [1, 1, 1], 4), hdr([4, -4, 4], [1, 1, 1], 4), hdr([4, 4, 4], [1, 0, 1], 4),
              hdr([4, 4, 4], [1, 1, 1], 10**12), hdr([4, 4, 3], [1, 1, 1], 2), hdr([4, 4, 4], [1, 1, 1], 10**12, "ascii")):
        with open(bad, "wb") as f:
            f.write(h + b"\0" * 64)
        assert subprocess.run([exe, "pcd", bad, out], capture_output=True).returncode == 3
    # legacy ascii 'rgb F' with a denormal literal (alpha 0, R < 128), nan and inf coordinates
    with open(bad, "wb") as f:
        f.write(("VERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\n"
                 "DATA ascii\n1 2 3 4.2e-39\nnan inf -1 5.9e-39\n").encode())
    assert subprocess.run([exe, "pcd", bad, out], capture_output=True).returncode == 0
    s2 = _surfels(out, 2)
    assert np.array_equal(s2[0, 0:3], np.float32([1, 2, 3])) and np.isnan(s2[1, 0]) and np.isinf(s2[1, 1])
    assert s2[:, 8].view(np.uint32)[0] == np.float32(4.2e-39).view(np.uint32)


@pytest.fixture(scope="module")
def cli(built, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cli") / "triplet_match")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "apps", "triplet_match.cpp"), "-o", out, "-L" + LIBDIR,
                           "-ltriplet_match_b200", "-Wl,-rpath," + LIBDIR])
    return out


def test_cli_usage_and_errors(cli, tmp_path):
    assert subprocess.run([cli], capture_output=True).returncode == 2
    r = subprocess.run([cli, str(tmp_path / "no.pcd"), str(tmp_path / "no2.pcd")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


@pytest.mark.gpu
def test_cli_finds_instances_from_pcd(cli, tmp_path):
    m, s, om, osc, rec = common.config("cylinder_small")
    mp, sp = str(tmp_path / "model.pcd"), str(tmp_path / "scene.pcd")
    write_pcd(m, mp, binary=True)
    write_pcd(s, sp, binary=False)
    r = subprocess.run([cli, mp, sp, "1.0", "0.5", "5", "0"], capture_output=True, text=True)  # analytic normals
    assert r.returncode == 0, r.stderr + r.stdout
    lines = [ln for ln in r.stdout.split("\n") if ln.startswith("match ")]
    assert len(lines) >= 1, r.stdout
    mpts = m.pos.astype(np.float64)
    for ln in lines:
        tok = ln.split()
        T = np.array([float(x) for x in tok[tok.index("T") + 1:]]).reshape(4, 4).T
        placed = mpts @ T[:3, :3].T + T[:3, 3]
        errs = [np.abs(placed - (mpts @ P[:3, :3].T + P[:3, 3])).max() for P in s.poses]
        assert min(errs) < 3 * om.resolution, (errs, r.stdout)
