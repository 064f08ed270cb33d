"""constant-operator assembly: host numpy shim vs the device
(per-cell kernel + gather), r-fold refinements of cylinder_4.
    python tools/bench_assembly.py [rmax]"""
import json
import sys
import time
import numpy as np
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, fem

rmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = _lib.default_context(0)
base = fem.load_mesh('cylinder_4')
for r in range(rmax + 1):
    mesh = fem.refine_uniform(base, r) if r > 0 else base
    V, Q = fem.VectorP2Space(mesh), fem.P1Space(mesh)
    t0 = time.perf_counter()
    host = fem.assemble_stokes_operators(V, Q, nu=1e-3)
    t_host = time.perf_counter() - t0
    dev = _lib.device_for(V, ctx)
    dev.pattern
    dev.assemble_stokes(Q, nu=1e-3)                 # warm-up (pattern, slot lists)
    t0 = time.perf_counter()
    ctx.profile_begin(1000)
    out = dev.assemble_stokes(Q, nu=1e-3)
    prof = ctx.profile_end()
    t_dev = time.perf_counter() - t0
    err = max(abs(host[k] - out[k]).max()/abs(host[k]).max() for k in ('M', 'A', 'J', 'MP'))
    print(json.dumps(dict(refine=r, ncell=mesh.num_cells, dofs=V.dim() + Q.dim(), host_s=t_host,
                          device_call_s=t_dev, kernels_us={k: v[1]*1e3 for k, v in prof.items()},
                          relerr=err)), flush=True)
