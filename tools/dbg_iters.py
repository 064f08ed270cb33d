import numpy as np, sys
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, time_int_utils as tiu, lin_alg_utils as lau
from oracle import snu as osnu
N = int(sys.argv[1]); nts = int(sys.argv[2])
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=60, scheme='TH', mergerhs=True, meshparams=dict(refinement_level=N))
inv = femp['invinds']
vp = lau.solve_sadpnt_smw(amat=sm['A'], jmat=sm['J'], jmatT=sm['JT'], rhsv=rhsd['fv'], rhsp=rhsd['fp'], krylov='gmres', vgroups=(inv//2, inv%2), krpslvprms=dict(tol=1e-11, maxiter=1500))
NV = sm['J'].shape[1]
for guess in (0, 1, 4, 8, 12):
    integ = tiu.DeviceImex(sm['M'], sm['A'], sm['J'], femp['V'], inv, femp['dbcinds'], femp['dbcvals'], 1./nts, fv=rhsd['fv'], fp=rhsd['fp'])
    integ.set_state(vp[:NV], -vp[NV:])
    its = []
    for k in range(8):
        integ.run(5, tol=1e-11, guess=guess, ntimeslices=0)
        st = integ.stats(); its.append(round(st['iters']/st['solves'], 1))
    print('guess', guess, 'iters/solve per 5-step block', its, 'ms/step', integ.engine.last_run_ms()/5)
    integ.close()
