"""kernel-level roofline of K1a (convection vector), K1b (convection matrices)
and the CSR SpMM on r-fold uniform refinements of cylinder_4 (SURVEY.md 8d.6:
inputs that leave the 126 MB L2).  Times are CUDA-event times of the kernels
alone (dnsb_profile_begin/end), bytes are the ALGORITHMIC bytes of DESIGN.md 5.

    python tools/bench_kernels.py [rmax] > gpurun_out/kernels.json
"""
import json
import sys
import numpy as np
import scipy.sparse as sps
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, fem

PEAK = 6549.1
try:
    PEAK = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass
rmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = _lib.default_context(0)
base = fem.load_mesh('cylinder_4')
out = []


def timed(fn, reps=5):
    fn()                       # warm-up
    ctx.profile_begin(100000)
    for _ in range(reps):
        fn()
    prof = ctx.profile_end()
    return {k: v[1]*1e-3/reps for k, v in prof.items()}   # seconds per call


for r in range(rmax + 1):
    mesh = fem.refine_uniform(base, r) if r > 0 else base
    V = fem.VectorP2Space(mesh)
    dev = _lib.ConvDevice(V, ctx)
    ncell, nvf = mesh.num_cells, V.dim()
    rng = np.random.default_rng(0)
    for nb in (1, 8, 64):
        if nvf*nb*8 > 2e9:
            continue
        U = rng.standard_normal((nvf, nb)) if nb > 1 else rng.standard_normal(nvf)
        t = sum(timed(lambda: dev.convvec(U)).values())
        by = 352.*ncell*nb
        out.append(dict(kernel='k_conv_elem+k_conv_gather (K1a)', refine=r, ncell=ncell,
                        dofs=nvf, nb=nb, us=t*1e6, algorithmic_MB=by/1e6, GBs=by/t/1e9, frac=by/t/1e9/PEAK))
    u = rng.standard_normal(nvf)
    indptr, indices = dev.pattern
    prof = timed(lambda: dev.convmats(u))
    t = sum(v for k, v in prof.items() if 'convmats' in k)     # f3 is K1a (timed above)
    by = 3000.*ncell
    out.append(dict(kernel='k_convmats_elem+k_convmats_gather (K1b, Newton parts)', refine=r, ncell=ncell,
                    dofs=nvf, nb=1, us=t*1e6, split_us={k: v*1e6 for k, v in prof.items()},
                    algorithmic_MB=by/1e6, GBs=by/t/1e9, frac=by/t/1e9/PEAK))
    # CSR SpMM on the P2 vector pattern (the pattern of the sym-grad stiffness)
    nnz = indices.size
    A = sps.csr_matrix((rng.standard_normal(nnz), indices, indptr), shape=(nvf, nvf))
    # device numbering: Hilbert curve of the nodes, components adjacent
    from dolfin_navier_scipy_b200 import hostsetup
    perm = hostsetup.locality_perm(np.asarray(V.tabulate_dof_coordinates()), comp=np.arange(nvf) % 2)
    A = A[perm][:, perm].tocsr()
    mat = ctx.csr(A)
    for nb in (1, 8, 64):
        if nvf*nb*8 > 2e9:
            continue
        X = rng.standard_normal((nvf, nb)) if nb > 1 else rng.standard_normal(nvf)
        prof = timed(lambda: mat.spmm(X))
        t = sum(prof.values())
        by = 12.*nnz + 4.*(nvf + 1) + 16.*nvf*nb
        out.append(dict(kernel='+'.join(sorted(prof)), refine=r, nnz=nnz, dofs=nvf, nb=nb, us=t*1e6,
                        algorithmic_MB=by/1e6, GBs=by/t/1e9, frac=by/t/1e9/PEAK))
    mat.close()
    del dev
for o in out:
    print(json.dumps(o))
