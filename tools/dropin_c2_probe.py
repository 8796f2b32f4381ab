"""Dev probe: the drop-in C++ API end to end (model::init + scene::find_all_parallel) on the C2 clouds."""
import os, subprocess, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from triplet_match_b200 import workloads as wl
import __graft_entry__ as ge
ge.build()
if os.environ.get("C3"):
    from triplet_match_b200 import synth
    model = synth.freeform_model(seed=3, n_points=50000, radius=0.01 * np.sqrt(50000 / (4 * np.pi)), n_bumps=12, n_curves=8)
    scene = synth.make_scene(seed=3, model=model, n_points=10_000_000, n_copies=8, extent=10.0 * np.sqrt(10.0), flat_copies=False)
else:
    model, scene = wl.c2_clouds()
if os.environ.get("SHUFFLE"):
    from triplet_match_b200 import synth
    scene = scene.take(synth.shuffle_perm(9, 1, scene.n))
d = tempfile.mkdtemp()
exe = os.path.join(d, "test_dropin")
lib = os.path.join(ROOT, "triplet_match_b200")
subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                       os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), "-o", exe, "-L" + lib, "-ltriplet_match_b200", "-Wl,-rpath," + lib])
def write(c, path):
    rec = np.concatenate([c.pos, c.nrm, c.tgt], axis=1).astype(np.float32)
    with open(path, "wb") as f:
        f.write(np.uint32(c.n).tobytes()); f.write(np.ascontiguousarray(rec).tobytes())
mp, sp, op = os.path.join(d, "m.bin"), os.path.join(d, "s.bin"), os.path.join(d, "o.txt")
write(model, mp); write(scene, sp)
for rep in range(2):
    t = time.time(); r = subprocess.run([exe, "find", mp, sp, op, "nocurv"], capture_output=True, text=True, env=dict(os.environ, TM_TRACE="1")); dt = time.time() - t
    print("wall", round(dt, 2), "s;", " | ".join(r.stdout.strip().split("\n")[-2:]), r.stderr[-2500:] if rep else "")
lines = open(op).read().strip().split("\n")
n = int(lines[0]); mpts = model.pos.astype(np.float64)
for ln in lines[1:1 + n]:
    v = [float(x) for x in ln.split()[:18]]
    T = np.array(v[2:18]).reshape(4, 4).T
    placed = mpts @ T[:3, :3].T + T[:3, 3]
    errs = [np.abs(placed - (mpts @ P[:3, :3].T + P[:3, 3])).max() for P in scene.poses]
    print("  instance", int(np.argmin(errs)), "inliers", int(v[0]), "max err", round(min(errs), 5))
