#!/usr/bin/env python
"""determinism check of the programmatic-dependent-launch path: the same
64-member run under different switch settings, states compared bitwise"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from dolfin_navier_scipy_b200 import _lib, problem_setups as dnsps     # noqa: E402
from dolfin_navier_scipy_b200 import time_int_utils as tiu            # noqa: E402


_SYS = {}


def run(env, nsteps=int(os.environ.get('NSTEPS', '30')), lvl=int(os.environ.get('LVL', '2'))):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctx = _lib.Context()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if lvl not in _SYS:
        _SYS[lvl] = dnsps.get_sysmats(problem='cylinderwake', Re=80., scheme='TH', mergerhs=True,
                                      meshparams=dict(refinement_level=lvl))
    femp, sm, rhsd = _SYS[lvl]
    inv = femp['invinds']
    nus = femp['nu']*np.linspace(.6, 1.6, 64)
    integ = tiu.DeviceImex(sm['M'], sm['A']/femp['nu'], sm['J'], femp['V'], inv, femp['dbcinds'],
                           femp['dbcvals'], 1./1024, nus=nus, fv=rhsd['fv'], fp=rhsd['fp'], ctx=ctx)
    rng = np.random.default_rng(0)
    v0 = np.zeros((sm['M'].shape[0], 1))
    integ.set_state(v0, np.zeros((sm['J'].shape[0], 1)))
    integ.run(nsteps, tol=1e-12, guess=int(os.environ.get('GUESS', '8')))
    v, p = integ.state()
    st = integ.stats()
    integ.close()
    return v.copy(), p.copy(), st


def main():
    cases = [dict(DNSB_PDL='0')]
    for pat in sys.argv[1:]:
        key, val = pat.split('=')
        cases += [{'DNSB_PDL': '1', key: val, ('DNSB_PDL_SKIP' if key == 'DNSB_PDL_ONLY' else 'DNSB_PDL_ONLY'): ''}]*int(os.environ.get('REPEAT', '3'))
    ref = None
    for env in cases:
        v, p, st = run(env)
        if ref is None:
            ref = (v, p)
        print(env, 'iters', st['iters'], 'relres', st['max_relres'],
              'bitwise == first:', bool(np.array_equal(v, ref[0]) and np.array_equal(p, ref[1])),
              'maxdiff', float(np.abs(v - ref[0]).max()))


if __name__ == '__main__':
    main()
