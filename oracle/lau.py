"""Oracle: `sadptprj_riclyap_adi.lin_alg_utils.solve_sadpnt_smw`.

TEST INFRASTRUCTURE (see `oracle/__init__.py`).

The module is a third-party dependency that is *not* vendored under
/root/reference (bare name in `requirements.txt:6`, no version pin).  Its
contract is fixed by the reference's call sites
(`stokes_navier_utils.py:401,458,497,903-907,1505-1512,1629-1633`,
`time_int_utils.py:402,408,466,605`): solve

    [ amat  jmatT ] [ v ]   [ rhsv ]
    [ jmat    0   ] [ p ] = [ rhsp ]

exactly; returns the stacked ``(NV+NP, k)`` array [and the factorisation with
``return_alu``].  Restated with scipy's SuperLU on the assembled block matrix,
the same thing `time_int_utils.py:89-91` does inline.
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla


def saddle_matrix(amat, jmat, jmatT=None):
    jmatT = jmat.T if jmatT is None else jmatT
    NP = jmat.shape[0]
    return sps.vstack([sps.hstack([amat, jmatT]),
                       sps.hstack([jmat, sps.csr_matrix((NP, NP))])],
                      format='csc')


class SadLU(object):
    """the factorisation handed out with ``return_alu``: the reference CALLS
    it on a stacked right-hand side (`time_int_utils.py:604-613`), so it is
    a `scipy.sparse.linalg.factorized`-like callable"""

    def __init__(self, kmat):
        self.lu = spsla.splu(kmat)

    def solve(self, rhs):
        return self.lu.solve(rhs)

    __call__ = solve


def solve_sadpnt_smw(amat=None, jmat=None, rhsv=None, jmatT=None,
                     rhsp=None, sadlu=None, return_alu=False, **kw):
    NP, NV = jmat.shape
    if sadlu is None:
        sadlu = SadLU(saddle_matrix(amat, jmat, jmatT))
    rhsv = np.asarray(rhsv, dtype=float).reshape(NV, -1)
    if rhsp is None:
        rhsp = np.zeros((NP, rhsv.shape[1]))
    rhsp = np.asarray(rhsp, dtype=float).reshape(NP, -1)
    vp = sadlu.solve(np.vstack([rhsv, rhsp]))
    if return_alu:
        return vp, sadlu
    return vp
