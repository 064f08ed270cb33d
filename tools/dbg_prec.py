"""device preconditioner vs a numpy restatement of the same hierarchy"""
import sys
import numpy as np, scipy.sparse as sps
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, lin_alg_utils as lau, hostsetup
from oracle import convection as oconv, snu as osnu
from oracle.lau import solve_sadpnt_smw as olu
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=60, scheme='TH', mergerhs=True,
                                   meshparams=dict(refinement_level=1))
A, M, J = sm['A'].tocsr(), sm['M'].tocsr(), sm['J'].tocsr()
NP, NV = J.shape
inv = np.asarray(femp['invinds'])
fv, fp = rhsd['fv'], rhsd['fp']
ref = olu(amat=A, jmat=J, jmatT=J.T, rhsv=fv, rhsp=fp)
vfull = osnu.append_bcs_vec(ref[:NV], femp['V'].dim(), inv, femp['dbcinds'], femp['dbcvals'])
N1, N2, f3 = oconv.convmats(femp['V'], vfull.ravel())
N1c = N1[inv][:, inv]


def cheb(Fm, dinv, r, k, lmin, lmax):
    th = .5*(lmax+lmin); de = .5*(lmax-lmin); sigma = th/de; rho = 1./sigma
    z = np.zeros_like(r); res = r.copy(); d = dinv*res/th
    for i in range(k):
        z = z + d
        if i == k-1:
            break
        res = res - Fm@d
        rho_n = 1./(2*sigma - rho)
        d = rho_n*rho*d + 2*rho_n/de*(dinv*res); rho = rho_n
    return z


rng = np.random.default_rng(0)
for name, F in (('stokes', A), ('picard', (A + N1c).tocsr())):
    op = lau.SadpntOperator(F, J, J.T.tocsr(), vgroups=(inv//2, inv % 2), mass_diag=M.diagonal(), vcoarse_max=500) \
        if False else lau.SadpntOperator(F, J, J.T.tocsr(), vgroups=(inv//2, inv % 2), mass_diag=M.diagonal())
    info = op.info
    kF = 2 if name == 'picard' else 3
    vlevels, vdense = info['vhierarchy']
    lmin0, lmax0 = info['spectrum']
    slevels, sdense = info['hierarchy']
    du_inv = 1./M.diagonal()
    r = rng.standard_normal(NV + NP)
    z = op.solver.apply_prec(r).ravel()
    rv, rp = r[:NV], r[NV:]
    assert len(slevels) == 0
    t = sdense@rp
    t = du_inv*(J.T@t); t = du_inv*(F@t)
    zp = -(sdense@(J@t))
    b = rv - J.T@zp
    dinv = 1./F.diagonal()
    if len(vlevels) > 0:
        x = cheb(F, dinv, b, kF, lmin0, lmax0)
        rr = b - F@x
        xc = vdense@(vlevels[0]['R']@rr) if len(vlevels) == 1 else None
        x = x + vlevels[0]['P']@xc
        rr = b - F@x
        zv = x + cheb(F, dinv, rr, kF, lmin0, lmax0)
    else:
        zv = cheb(F, dinv, b, kF, lmin0, lmax0)
    print(name, 'vlevels', [l['A'].shape[0] for l in vlevels], 'coarse', vdense.shape, 'spectrum', info['spectrum'],
          'zp rel diff %.2e' % (np.linalg.norm(z[NV:] - zp)/np.linalg.norm(zp)),
          'zv rel diff %.2e' % (np.linalg.norm(z[:NV] - zv)/np.linalg.norm(zv)))
    op.close()
