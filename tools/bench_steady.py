#!/usr/bin/env python
"""BASELINE config 2 (DFG 2D-1 steady state: Stokes -> Picard -> Newton on
karman2D-rotcyl_lvl1): device path against the oracle's SuperLU path.

    python tools/bench_steady.py [--oracle]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np                                                    # noqa: E402


def main():
    from dolfin_navier_scipy_b200 import problem_setups as dnsps
    from dolfin_navier_scipy_b200 import stokes_navier_utils as snu
    from conftest import soldict
    femp, sm, rhsd = dnsps.get_sysmats(
        problem='gen_bccont', nu=1e-3, charvel=.2, scheme='TH', mergerhs=True,
        meshparams=dict(strtomeshfile='mesh/karman2D-rotcyl_lvl1.xml.gz',
                        movingwallcntrl=False,
                        strtophysicalregions='mesh/karman2D-rotcyl_lvl1_'
                        'facet_region.xml.gz',
                        strtobcsobs='mesh/karman2D-rotcyl-bm_geo_cntrlbc.json'))
    sd = soldict(femp, sm, rhsd)
    out = dict(dofs=int(sm['J'].shape[0] + sm['J'].shape[1]))
    for rep in range(2):           # second call: contexts, meshes and kernels warm
        t0 = time.perf_counter()
        v, p = snu.solve_steadystate_nse(return_vp=True, verbose=False, **sd)
        out['device_s_call%d' % rep] = time.perf_counter() - t0
    if '--profile' in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        snu.solve_steadystate_nse(return_vp=True, verbose=False, **sd)
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats('cumulative').print_stats(45)
    if '--oracle' in sys.argv:
        from oracle import snu as osnu
        t0 = time.perf_counter()
        vo, po = osnu.solve_steadystate_nse(return_vp=True, verbose=False, **sd)
        out['oracle_s'] = time.perf_counter() - t0
        out['v_rel'] = float(np.linalg.norm(v - vo)/np.linalg.norm(vo))
        out['p_rel'] = float(np.linalg.norm(p - po)/np.linalg.norm(po))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
