"""K1b (convection matrices) alone on r-fold refinements of cylinder_4, split by kernel.
    python tools/bench_k1b.py [rmax]"""
import json
import sys
import numpy as np
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, fem

PEAK = 6549.1
rmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = _lib.default_context(0)
base = fem.load_mesh('cylinder_4')
for r in range(rmax + 1):
    mesh = fem.refine_uniform(base, r) if r > 0 else base
    V = fem.VectorP2Space(mesh)
    dev = _lib.ConvDevice(V, ctx)
    u = np.random.default_rng(0).standard_normal(V.dim())
    dev.convmats(u)
    ctx.profile_begin(100000)
    for _ in range(5):
        dev.convmats(u)
    prof = {k: v[1]*1e3/5 for k, v in ctx.profile_end().items()}
    t = sum(v for k, v in prof.items() if 'convmats' in k)*1e-6
    by = 3000.*mesh.num_cells
    print(json.dumps(dict(kernel='K1b', refine=r, ncell=mesh.num_cells, us=t*1e6, split_us=prof,
                          GBs=by/t/1e9, frac=by/t/1e9/PEAK)), flush=True)
