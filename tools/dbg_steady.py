import sys, logging, time
import numpy as np
sys.path.insert(0, '.')
logging.basicConfig(level=logging.INFO, format='%(message)s')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, stokes_navier_utils as snu, lin_alg_utils as lau
femp, sm, rhsd = dnsps.get_sysmats(
    problem='gen_bccont', nu=1e-3, charvel=.2, scheme='TH', mergerhs=True,
    meshparams=dict(strtomeshfile='mesh/karman2D-rotcyl_lvl1.xml.gz', movingwallcntrl=False,
                    strtophysicalregions='mesh/karman2D-rotcyl_lvl1_facet_region.xml.gz',
                    strtobcsobs='mesh/karman2D-rotcyl-bm_geo_cntrlbc.json'))
orig = lau.solve_sadpnt_smw
def wrapped(**kw):
    t = time.time()
    st = []
    kw['krpslvprms'] = dict(kw.get('krpslvprms', {}), convstatsl=st)
    out = orig(**kw)
    print('   solve: iters', st, 'time %.2fs' % (time.time() - t))
    return out
lau.solve_sadpnt_smw = wrapped
d = dict(A=sm['A'], M=sm['M'], J=sm['J'], JT=sm['JT'], fv=rhsd['fv'], fp=rhsd['fp'], V=femp['V'],
         invinds=femp['invinds'], dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
try:
    (v, p), norms = snu.solve_steadystate_nse(return_vp=True, return_nwtnupd_norms=True, verbose=True, **d)
    print('newton norms', norms)
except UserWarning as e:
    print('FAILED', e)
