import sys
import numpy as np
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, time_int_utils as tiu
from oracle import snu as osnu
from oracle.lau import solve_sadpnt_smw as olu
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=60, scheme='TH', mergerhs=True, meshparams=dict(refinement_level=1))
inv = femp['invinds']
sd = dict(A=sm['A'], M=sm['M'], J=sm['J'], JT=sm['JT'], fv=rhsd['fv'], fp=rhsd['fp'], V=femp['V'], invinds=inv, dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
v0 = osnu.solve_nse(t0=0, tE=1./512, Nts=1, start_ssstokes=True, return_vp_dict=True, **sd)[0.0]['v'][inv]
dt = 1./512
M, A, J = sm['M'], sm['A'], sm['J']
NV = M.shape[0]
def step(v):
    _, nfc, _ = osnu.get_v_conv_conts(vvec=v, V=femp['V'], invinds=inv, dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'], semi_explicit=True)
    return olu(amat=M + dt*A, jmat=J, jmatT=J.T, rhsv=M@v + dt*(rhsd['fv'] + nfc), rhsp=rhsd['fp'])[:NV]
ref = [v0]
for k in range(4):
    ref.append(step(ref[-1]))
for guess in (0, 16):
    for reorder in (True, False):
        integ = tiu.DeviceImex(M, A, J, femp['V'], inv, femp['dbcinds'], femp['dbcvals'], dt, scheme='imexeuler', nus=(1.,), fv=rhsd['fv'], fp=rhsd['fp'], reorder=reorder)
        integ.set_state(v0)
        integ.run(4, snap_stride=1, ntimeslices=0, guess=guess)
        vs, ps = integ.snapshots()
        print('guess', guess, 'reorder', reorder, [float(np.linalg.norm(vs[k, :, :1] - ref[k])/np.linalg.norm(ref[k])) for k in range(5)], integ.stats())
        integ.close()
ref = [v0]
for k in range(12):
    ref.append(step(ref[-1]))
trange = np.linspace(0., 12./512, 13)
got = tiu.semi_implicit_euler(iniv=v0, jmat=J, mmat=M, amat=A, trange=trange, fp=rhsd['fp'], V=femp['V'], invinds=inv,
                              dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'], fv=rhsd['fv'])
print('wrapper', [float(np.linalg.norm(got[k] - ref[k])/np.linalg.norm(ref[k])) for k in range(13)])
integ = tiu.DeviceImex(M, A, J, femp['V'], inv, femp['dbcinds'], femp['dbcvals'], dt, scheme='imexeuler', nus=(1.,), fv=rhsd['fv'], fp=rhsd['fp'])
integ.set_state(v0)
integ.run(12, snap_stride=1, ntimeslices=0)
vs, ps = integ.snapshots()
print('direct 12', [float(np.linalg.norm(vs[k, :, :1] - ref[k])/np.linalg.norm(ref[k])) for k in range(13)])
