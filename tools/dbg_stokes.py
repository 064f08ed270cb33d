import sys, time
import numpy as np, scipy.sparse as sps
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, lin_alg_utils as lau
from oracle import convection as oconv, snu as osnu
from oracle.lau import solve_sadpnt_smw as olu
which = sys.argv[1]
if which == 'dfg':
    femp, sm, rhsd = dnsps.get_sysmats(
        problem='gen_bccont', nu=1e-3, charvel=.2, scheme='TH', mergerhs=True,
        meshparams=dict(strtomeshfile='mesh/karman2D-rotcyl_lvl1.xml.gz', movingwallcntrl=False,
                        strtophysicalregions='mesh/karman2D-rotcyl_lvl1_facet_region.xml.gz',
                        strtobcsobs='mesh/karman2D-rotcyl-bm_geo_cntrlbc.json'))
else:
    femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=60, scheme='TH', mergerhs=True,
                                       meshparams=dict(refinement_level=2))
A, M, J = sm['A'].tocsr(), sm['M'].tocsr(), sm['J'].tocsr()
NP, NV = J.shape
inv = np.asarray(femp['invinds'])
fv, fp = rhsd['fv'], rhsd['fp']
ref = olu(amat=A, jmat=J, jmatT=J.T, rhsv=fv, rhsp=fp)
vfull = osnu.append_bcs_vec(ref[:NV], femp['V'].dim(), inv, femp['dbcinds'], femp['dbcvals'])
N1, N2, f3 = oconv.convmats(femp['V'], vfull.ravel())
N1c = N1[inv][:, inv]
K = lambda F: sps.bmat([[F, J.T], [J, None]], format='csr')
b = np.vstack([fv, fp])
for name, F in (('stokes', A), ('picard', (A + N1c).tocsr())):
    for schur in ('lsc', 'lumped'):
        for nsm in (2,):
            t = time.time()
            op = lau.SadpntOperator(F, J, J.T.tocsr(), vgroups=(inv//2, inv % 2), mass_diag=M.diagonal(), schur=schur, nsmooth=nsm)
            ts = time.time() - t
            t = time.time()
            vp = op.solve(fv, fp, tol=1e-12, maxit=600)
            tr = np.linalg.norm(K(F)@vp - b)/np.linalg.norm(b)
            print(which, name, schur, 'nsmooth', nsm, 'iters', op.last_iters, 'relres est', op.last_relres, 'true', tr,
                  'vamg', op.info['velocity_amg'], 'vlevels', [l['A'].shape[0] for l in op.info['vhierarchy'][0]] if op.info['vhierarchy'] else None,
                  'setup %.1fs solve %.2fs' % (ts, time.time() - t))
            op.close()
