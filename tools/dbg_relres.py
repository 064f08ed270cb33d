"""debug: convergence statistics of ensemble runs (usage: python tools/dbg_relres.py <mesh>)"""
import sys; sys.path.insert(0,'.')
import numpy as np
sys.argv.append('1')
from dolfin_navier_scipy_b200 import ensemble as ens
N = int(sys.argv[1])
import bench
for dt in (1./512,):
    integ, info = ens.cylinder_ensemble(N=N, nmembers=64, dt=dt, ntimes=200)
    integ.set_state(*bench.ensemble_initial_state(info, 64))
    for k in (32, 40):
        integ.run(k, tol=1e-12, ntimeslices=0); print(dt, integ.stats())
    integ.close()
