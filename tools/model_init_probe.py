"""Dev probe: where model::init (tm_hostmodel_build + upload) spends its time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from triplet_match_b200 import capi, synth
ctx = capi.Context(0)
for name, m in (("plane 10k (C2)", synth.plane_model(seed=2, size=1.0, res=0.01, n_curves=12)),
                ("cylinder 16k", synth.cylinder_model(seed=1, radius=0.25, height=1.0, res=0.01, n_curves=4)),
                ("free-form 50k (C3)", synth.freeform_model(seed=3, n_points=50000, radius=0.01 * np.sqrt(50000 / (4 * np.pi)), n_bumps=12, n_curves=8))):
    for rep in range(2):
        t0 = time.perf_counter(); res = capi.host_resolution(m.pos); t1 = time.perf_counter()
        hm = capi.HostModel(ctx, m.pos, m.nrm, m.tgt, curv_ok=m.tangent_mask, resolution=res, min_df=0.2, max_df=1.0, cap=200); t2 = time.perf_counter()
        ext = hm.extents.astype(np.int64)
        vox = ctx.voxel_fill(m.pos, m.nrm, m.tgt, hm.extents, hm.to_voxel16); t3 = time.perf_counter()
        gm = hm.upload(ctx); ctx.sync(); t4 = time.perf_counter()
        gm.close(); hm.close()
    print(f"{name}: n {m.n} tangent {int(m.tangent_mask.sum())} cells {int(ext.prod())} entries {hm.n_entries} | resolution {1e3*(t1-t0):.0f} ms, "
          f"hostmodel_build (incl. voxel fill) {1e3*(t2-t1):.0f} ms, voxel fill alone {1e3*(t3-t2):.0f} ms, upload {1e3*(t4-t3):.0f} ms", flush=True)
