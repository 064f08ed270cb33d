"""CSR SpMV/SpMM on r-fold refinements of cylinder_4: the row
kernels against the TMA-staged kernel (dnsb_stream.cuh), CUDA-event times of
the kernels alone, ALGORITHMIC bytes 12 nnz + 4 (n+1) + 16 n nb.

    python tools/bench_spmm.py [rmin] [rmax] > gpurun_out/spmm.jsonl
"""
import json
import os
import sys
import numpy as np
import scipy.sparse as sps
import torch
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, fem, hostsetup

PEAK = 6549.1
try:
    PEAK = float(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])
except Exception:
    pass
rmin = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ROWS = [int(a) for a in sys.argv[3].split(',')] if len(sys.argv) > 3 else [64]
base = fem.load_mesh('cylinder_4')


def timed(ctx, fn, reps=20):
    """back-to-back launches on the library's stream, each bracketed by CUDA events"""
    fn()
    ctx.sync()
    ctx.profile_begin(100000)
    for _ in range(reps):
        fn()
    ctx.sync()
    prof = ctx.profile_end()
    return {k: v[1]*1e-3/reps for k, v in prof.items()}


for r in range(rmin, rmax + 1):
    mesh = fem.refine_uniform(base, r) if r > 0 else base
    V = fem.VectorP2Space(mesh)
    nvf = V.dim()
    indptr, indices = _lib.ConvDevice(V, _lib.default_context(0)).pattern
    rng = np.random.default_rng(0)
    nnz = indices.size
    A = sps.csr_matrix((rng.standard_normal(nnz), indices, indptr), shape=(nvf, nvf))
    perm = hostsetup.locality_perm(np.asarray(V.tabulate_dof_coordinates()), comp=np.arange(nvf) % 2)
    A = A[perm][:, perm].tocsr()
    cfgs = [('rows', dict(DNSB_TMA_MIN_ROWS='2000000000'))]
    for rows in ROWS:
        for stages in (2,):
            cfgs.append(('staged r%d s%d' % (rows, stages),
                         dict(DNSB_TMA_MIN_ROWS='0', DNSB_TMA_ROWS=str(rows), DNSB_TMA_STAGES=str(stages))))
    for mode, env in cfgs:
        os.environ.update(env)
        ctx = _lib.Context(0)
        mat = ctx.csr(A)
        for nb in (1,):
            X = rng.standard_normal((nvf, nb)) if nb > 1 else rng.standard_normal(nvf)
            err = np.linalg.norm(mat.spmm(X) - A@X)/np.linalg.norm(A@X)
            xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()      # device-resident operands
            yd = torch.empty(nvf*nb, dtype=torch.float64, device='cuda')
            torch.cuda.synchronize()
            prof = timed(ctx, lambda: ctx.check(ctx.lib.dnsb_spmm_dev(mat.h, None, xd.data_ptr(), yd.data_ptr(),
                                                                       nb, 1., 0.)))
            err = max(err, float(np.linalg.norm(yd.cpu().numpy().reshape(np.shape(X)) - A@X)/np.linalg.norm(A@X)))
            t = sum(prof.values())
            by = 12.*nnz + 4.*(nvf + 1) + 16.*nvf*nb
            print(json.dumps(dict(kernel='+'.join(sorted(prof)), mode=mode, refine=r, nnz=nnz, dofs=nvf, nb=nb,
                                  us=t*1e6, algorithmic_MB=by/1e6, GBs=by/t/1e9, frac=by/t/1e9/PEAK,
                                  relerr=err)), flush=True)
        mat.close()
