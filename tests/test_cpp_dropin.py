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


@pytest.mark.gpu
def test_cpp_find_all_parallel(exe, tmp_path):
    m, s, om, osc, rec = common.config("freeform_small")
    mp, sp, op = str(tmp_path / "m.bin"), str(tmp_path / "s.bin"), str(tmp_path / "o.txt")
    _write(m, mp)
    _write(s, sp)
    r = subprocess.run([exe, "find", mp, sp, op], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
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
        assert errs[k] < 3 * om.resolution, (errs, r.stdout)  # every reported instance is a real one
        assert int(v[0]) >= 0.5 * m.n
        found.add(k)
    assert len(found) == n  # no instance is reported twice (overlap-free acceptance)
