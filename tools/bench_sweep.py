"""Newton/CN sweep rate (BASELINE config 3): cylinder_<N>, Re=100, two-call
recipe (IMEX trajectory -> lin_vel_point), 1 Picard + 1 Newton sweep.
usage: python tools/bench_sweep.py <mesh> <Nts>
(the CPU oracle's rate for the same sweeps is measured by
tests/tools/oracle_sweep_rate.py -- only tests/ may execute oracle/)"""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, stokes_navier_utils as snu
N, Nts = int(sys.argv[1]), int(sys.argv[2])
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=100, scheme='TH', mergerhs=True,
                                   meshparams=dict(refinement_level=N))
sd = dict(A=sm['A'], M=sm['M'], J=sm['J'], JT=sm['JT'], fv=rhsd['fv'], fp=rhsd['fp'], V=femp['V'],
          invinds=femp['invinds'], dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'],
          t0=0., tE=Nts/2048., Nts=Nts, start_ssstokes=True)
mod = snu
sweep_times = []
from dolfin_navier_scipy_b200 import _lib
_run = _lib.CnSweep.run


def timed_run(self, *a, **k):
    t = time.perf_counter()
    out = _run(self, *a, **k)
    sweep_times.append(time.perf_counter() - t)
    return out


_lib.CnSweep.run = timed_run
t0 = time.perf_counter()
traj = mod.solve_nse(return_dictofvelstrs=True, **sd)
t1 = time.perf_counter()
its = []
kw = dict(krpslvprms=dict(convstatsl=its))
out = mod.solve_nse(lin_vel_point=traj, treat_nonl_explicit=False, vel_pcrd_stps=1, vel_nwtn_stps=1,
                    return_dictofvelstrs=True, verbose=False, **kw, **sd)
t2 = time.perf_counter()
print('%s mesh %d: IMEX call %.2f s (%d steps), 2 sweeps %.2f s = %.1f ms per sweep step, mean FGMRES its %s'
      % ('device', N, t1 - t0, Nts, t2 - t1, 1e3*(t2 - t1)/(2*Nts),
         np.mean(its) if its else '-'))
if sweep_times:
    print('   dnsb_cnsweep_run alone (incl. upload of the linearisation trajectory and download of the new one): '
          + ', '.join('%.2f ms/step' % (1e3*t/Nts) for t in sweep_times))
