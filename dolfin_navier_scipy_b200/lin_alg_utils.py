"""Saddle-point solves on the device.

Stand-in for `sadptprj_riclyap_adi.lin_alg_utils` (`lau`), the un-vendored
third-party module the reference calls at `stokes_navier_utils.py:401,458,497,
903-907,1505-1512,1629-1633` and `time_int_utils.py:402,408,466,605`: the
sparse-LU solve of ``[[amat, jmatT], [jmat, 0]]`` is replaced by the
block-preconditioned FGMRES of ``libdnsb200`` (``dnsb_solver_*``).
"""
import numpy as np
import scipy.sparse as sps

from . import _lib
from . import hostsetup

__all__ = ['solve_sadpnt_smw', 'SadpntOperator']


class SadpntOperator(object):
    """reusable device solver for one saddle-point matrix (``return_alu``)"""

    def __init__(self, amat, jmat, jmatT=None, ncols=1, ctx=None,
                 cheb_steps=3, restart=60, coarse_max=4096, schur_diag=None,
                 spectrum=None, hierarchy=None, velocity_amg='auto',
                 vgroups=None, vhierarchy=None):
        self.ctx = _lib.default_context() if ctx is None else ctx
        self.NP, self.NV = jmat.shape
        self.ncols = ncols
        self.solver, self.info = hostsetup.make_saddle_solver(
            self.ctx, amat, jmat, jmatT, nb=ncols, restart=restart,
            cheb_steps=cheb_steps, coarse_max=coarse_max,
            schur_diag=schur_diag, spectrum=spectrum, hierarchy=hierarchy,
            velocity_amg=velocity_amg, vgroups=vgroups, vhierarchy=vhierarchy)

    def solve(self, rhsv, rhsp=None, x0=None, tol=1e-12, maxit=800):
        rhsv = np.asarray(rhsv, dtype=float).reshape(self.NV, self.ncols)
        if rhsp is not None:
            rhsp = np.asarray(rhsp, dtype=float).reshape(self.NP, self.ncols)
        vp, iters, relres = self.solver.solve(rhsv, rhsp, x0=x0, tol=tol,
                                              maxit=maxit)
        self.last_iters, self.last_relres = iters, relres
        return vp

    def close(self):
        self.solver.close()


def solve_sadpnt_smw(amat=None, jmat=None, rhsv=None, jmatT=None,
                     rhsp=None, umat=None, vmat=None, krylov=None,
                     krpslvprms={}, krplsprms={}, return_alu=False,
                     sadlu=None, decouplevp=False, solve_A=None,
                     symmetric=False, cgtol=1e-8, vgroups=None, **kw):
    """solve ``[[amat, jmatT], [jmat, 0]] [v; p] = [rhsv; rhsp]`` on the device

    Same arguments and return value as `lau.solve_sadpnt_smw` (stacked
    ``(NV+NP, k)`` array [and the reusable operator with ``return_alu``]).
    ``krpslvprms['tol'|'maxiter'|'x0']`` are honoured; ``convstatsl`` receives
    the iteration counts like krypy's convergence statistics
    (`tests/time_dep_nse_krylov.py:5,47`).  Low-rank updates (``umat``,
    ``vmat``) and ``decouplevp`` are outside the hot path and not supported.
    """
    if umat is not None or vmat is not None:
        raise NotImplementedError('Sherman-Morrison-Woodbury updates')
    if decouplevp:
        raise NotImplementedError('decoupled v/p solves')
    NP, NV = jmat.shape
    rhsv = np.asarray(rhsv, dtype=float).reshape(NV, -1)
    k = rhsv.shape[1]
    op = sadlu if sadlu is not None else \
        SadpntOperator(sps.csr_matrix(amat), sps.csr_matrix(jmat), jmatT,
                       ncols=k, vgroups=vgroups)
    tol = krpslvprms.get('tol', 1e-12) if krylov is not None else 1e-12
    maxit = krpslvprms.get('maxiter', 800) if krylov is not None else 800
    x0 = krpslvprms.get('x0', None) if krylov is not None else None
    vp = op.solve(rhsv, rhsp, x0=x0, tol=tol, maxit=maxit)
    if 'convstatsl' in krpslvprms:
        krpslvprms['convstatsl'].append(int(op.last_iters.max()))
    if return_alu:
        return vp, op
    if sadlu is None:
        op.close()
    return vp
