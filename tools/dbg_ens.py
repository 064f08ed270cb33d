import numpy as np, sys, time
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, ensemble as ens, lin_alg_utils as lau
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
guess = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ctx = _lib.default_context(0)
integ, info = ens.cylinder_ensemble(N=N, nmembers=nb, dt=1./2048, ntimes=200, ctx=ctx)
NV, NP = info['NV'], info['NP']
sm, inv = info['sm'], np.asarray(info['femp']['invinds'])
numean = float(np.mean(info['nus']))
Ast = numean*sm['A'] + info['Arob']
stats = []
t = time.time()
vp = lau.solve_sadpnt_smw(amat=Ast, jmat=sm['J'], jmatT=sm['JT'], rhsv=numean*info['B'][:, :1], rhsp=info['fp'],
                          krylov='gmres', vgroups=(inv//2, inv % 2), krpslvprms=dict(tol=1e-10, maxiter=1500, convstatsl=stats))
print('stokes its', stats, 'time', time.time()-t, 'norm v', np.linalg.norm(vp[:NV]), 'finite', np.all(np.isfinite(vp)))
import scipy.sparse as sps
K = sps.bmat([[Ast, sm['JT']], [sm['J'], None]], format='csr')
b = np.concatenate([numean*info['B'][:, 0], info['fp'].ravel()])
print('stokes true relres', np.linalg.norm(K@vp.ravel() - b)/np.linalg.norm(b))
v0 = np.repeat(vp[:NV], nb, axis=1); p0 = np.repeat(-vp[NV:], nb, axis=1)
integ.set_state(v0, p0)
for k in range(6):
    ff = integ.run(3, tol=1e-12, guess=guess, ntimeslices=0)
    v, p = integ.state(); st = integ.stats()
    print('after', 3*(k+1), 'steps: |v|', np.linalg.norm(v, axis=0)[:3], '|p|', np.linalg.norm(p, axis=0)[:3], st, 'ms', integ.engine.last_run_ms())
print('--- bench sequence')
vcur, pcur = integ.state()
integ.set_forcing(info['B'], info['U'])
integ.set_state(vcur, pcur)
ff = integ.run(10, snap_stride=1, tol=1e-12, guess=guess, ntimeslices=0)
vs, ps = integ.snapshots()
print('ff', ff, 'snap norms v', [float(np.linalg.norm(vs[k, :, 0])) for k in range(vs.shape[0])])
print('snap norms p', [float(np.linalg.norm(ps[k, :, 0])) for k in range(ps.shape[0])])
print(integ.stats())
