import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


@pytest.fixture(scope='session')
def cyl1():
    """cylinder wake, coarsest valid mesh, Re=60 (BASELINE config 1)"""
    from dolfin_navier_scipy_b200 import problem_setups as dnsps
    femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=60,
                                       scheme='TH', mergerhs=True,
                                       meshparams=dict(refinement_level=1))
    return femp, sm, rhsd


def soldict(femp, sm, rhsd, **kw):
    d = dict(A=sm['A'], M=sm['M'], J=sm['J'], JT=sm['JT'], fv=rhsd['fv'],
             fp=rhsd['fp'], V=femp['V'], invinds=femp['invinds'],
             dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
    d.update(kw)
    return d


@pytest.fixture(scope='session')
def cyl1_re30():
    """cylinder wake on the coarsest mesh at Re=30: the steady Picard/Newton
    systems of this Reynolds number are inside the envelope of the device's
    iterative solve on this mesh (at Re=60 the element Peclet number of the
    unstabilised Galerkin convection is too large: `NotConverged`)"""
    from dolfin_navier_scipy_b200 import problem_setups as dnsps
    return dnsps.get_sysmats(problem='cylinderwake', Re=30, scheme='TH',
                             mergerhs=True,
                             meshparams=dict(refinement_level=1))
