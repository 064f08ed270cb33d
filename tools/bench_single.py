"""single-trajectory step rate (BASELINE configs 1 and 3, IMEX part): CNAB on
cylinder_<N>, one member, device-resident loop vs. the oracle's SuperLU loop
usage: python tools/bench_single.py <mesh> <Re> <nts> <nsteps> [guess] [tol] [refine]
(refine = r: r-fold uniform refinement of the mesh)"""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, time_int_utils as tiu, lin_alg_utils as lau
N, Re, nts, nsteps = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
guess = int(sys.argv[5]) if len(sys.argv) > 5 else 16
tol = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-12
refine = int(sys.argv[7]) if len(sys.argv) > 7 else 0
meshparams = dict(refinement_level=N)
if refine:
    from dolfin_navier_scipy_b200 import fem
    meshparams['mesh'] = fem.refine_uniform(fem.load_mesh('cylinder_%d' % N), refine)
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=Re, scheme='TH', mergerhs=True,
                                   meshparams=meshparams)
inv = np.asarray(femp['invinds'])
NP, NV = sm['J'].shape
vp = lau.solve_sadpnt_smw(amat=sm['A'], jmat=sm['J'], jmatT=sm['JT'], rhsv=rhsd['fv'], rhsp=rhsd['fp'],
                          krylov='gmres', vgroups=(inv//2, inv % 2), mass_diag=sm['M'].diagonal(),
                          krpslvprms=dict(tol=1e-11, maxiter=1500))
integ = tiu.DeviceImex(sm['M'], sm['A'], sm['J'], femp['V'], inv, femp['dbcinds'], femp['dbcvals'], 1./nts,
                       fv=rhsd['fv'], fp=rhsd['fp'])
integ.set_state(vp[:NV], -vp[NV:])
integ.run(20, tol=tol, guess=guess, ntimeslices=0)          # warm-up (Heun start, graphs, projection space)
t0 = time.perf_counter()
integ.run(nsteps, tol=tol, guess=guess, ntimeslices=0)
wall = time.perf_counter() - t0
st = integ.stats()
dev_ms = integ.engine.last_run_ms()
print('mesh %d Re %g dt 1/%d: %d DoFs, device %.3f ms/step (wall %.3f), %.2f FGMRES its/step, relres %.1e, %d launches/step'
      % (N, Re, nts, NV + NP, dev_ms/nsteps, 1e3*wall/nsteps, st['iters']/max(st['solves'], 1), st['last_relres'],
         integ.ctx.launch_count()/(nsteps + 20)))
