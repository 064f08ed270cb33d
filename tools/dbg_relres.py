"""debug: convergence statistics of runs (usage: python tools/dbg_relres.py <mesh>)"""
import sys; sys.path.insert(0,'.')
import numpy as np
from dolfin_navier_scipy_b200 import ensemble as ens, problem_setups as dnsps, time_int_utils as tiu
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
import bench
integ, info = ens.cylinder_ensemble(N=N, nmembers=64, dt=1./1024, ntimes=200)
integ.set_state(*bench.ensemble_initial_state(info, 64))
print('guard run ff', integ.run(30, tol=1e-12, check_ff_maxv=1e8, ntimeslices=10), integ.stats())
integ.run(20, tol=1e-12, ntimeslices=0); print('plain run', integ.stats())
integ.close()
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=100., scheme='TH', mergerhs=True, meshparams=dict(refinement_level=N))
inv = np.asarray(femp['invinds'])
integ = tiu.DeviceImex(sm['M'], sm['A'], sm['J'], femp['V'], inv, femp['dbcinds'], femp['dbcvals'], 1./2048, fv=rhsd['fv'], fp=rhsd['fp'])
integ.set_state(np.zeros((inv.size,1)), np.zeros((sm['J'].shape[0],1)))
integ.run(20, tol=1e-12, ntimeslices=0); print('single a', integ.stats())
integ.run(20, tol=1e-12, ntimeslices=0); print('single b', integ.stats())
