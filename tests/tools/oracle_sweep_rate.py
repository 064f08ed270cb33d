"""CPU oracle rate of the Newton/CN sweeps of BASELINE config 3 (the number
profiles/README.md quotes beside tools/bench_sweep.py): cylinder_<N>, Re=100,
two-call recipe, 1 Picard + 1 Newton sweep.  Test infrastructure: lives under
tests/ because it executes oracle/.
usage: python tests/tools/oracle_sweep_rate.py <mesh> <Nts>"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dolfin_navier_scipy_b200 import problem_setups as dnsps   # noqa: E402
from oracle import snu as osnu                                  # noqa: E402
N, Nts = int(sys.argv[1]), int(sys.argv[2])
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=100, scheme='TH', mergerhs=True,
                                   meshparams=dict(refinement_level=N))
sd = dict(A=sm['A'], M=sm['M'], J=sm['J'], JT=sm['JT'], fv=rhsd['fv'], fp=rhsd['fp'], V=femp['V'],
          invinds=femp['invinds'], dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'],
          t0=0., tE=Nts/2048., Nts=Nts, start_ssstokes=True)
t0 = time.perf_counter()
traj = osnu.solve_nse(return_dictofvelstrs=True, **sd)
t1 = time.perf_counter()
osnu.solve_nse(lin_vel_point=traj, treat_nonl_explicit=False, vel_pcrd_stps=1, vel_nwtn_stps=1,
               return_dictofvelstrs=True, verbose=False, **sd)
t2 = time.perf_counter()
print('oracle mesh %d: IMEX call %.2f s (%d steps), 2 sweeps %.2f s = %.1f ms per sweep step'
      % (N, t1 - t0, Nts, t2 - t1, 1e3*(t2 - t1)/(2*Nts)))
