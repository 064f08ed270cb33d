"""per-step FGMRES iteration counts of the ensemble bench configuration
usage: python tools/dbg_ens_iters.py <mesh> <members> <guess> <nsteps> [nts]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, ensemble as ens, lin_alg_utils as lau
N, nmem, guess, nsteps = (int(a) for a in sys.argv[1:5])
nts = int(sys.argv[5]) if len(sys.argv) > 5 else 2048
ctx = _lib.default_context(0)
integ, info = ens.cylinder_ensemble(N=N, nmembers=nmem, dt=1./nts, ntimes=nsteps + 4, ctx=ctx)
sm, inv = info['sm'], np.asarray(info['femp']['invinds'])
NV = info['NV']
numean = float(np.mean(info['nus']))
Ast = numean*sm['A'] + info['Arob']
vp = lau.solve_sadpnt_smw(amat=Ast, jmat=sm['J'], jmatT=sm['JT'], rhsv=numean*info['B'][:, :1], rhsp=info['fp'],
                          krylov='gmres', vgroups=(inv//2, inv % 2), krpslvprms=dict(tol=1e-10, maxiter=1500))
integ.set_state(np.repeat(vp[:NV], integ.nb, axis=1), np.repeat(-vp[NV:], integ.nb, axis=1))
its = []
for k in range(nsteps):
    integ.run(1, tol=1e-12, guess=guess, ntimeslices=0, maxit=400)
    st = integ.stats()
    its.append(st['iters'] if st['solves'] else -1)
print('mesh', N, 'members', nmem, 'guess', guess, 'iters per step:', its)
