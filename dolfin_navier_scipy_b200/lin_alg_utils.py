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
                 cheb_steps=3, restart=None, coarse_max=4096, schur_diag=None,
                 spectrum=None, hierarchy=None, velocity_amg='auto',
                 vgroups=None, vhierarchy=None, mass_diag=None, cache=None,
                 schur='auto', nsmooth=2):
        self.ctx = _lib.default_context() if ctx is None else ctx
        self.NP, self.NV = jmat.shape
        self.ncols = ncols
        if restart is None:
            # long recurrences for single systems (the Oseen/Newton matrices
            # need ~150 iterations), bounded basis memory for wide batches
            restart = 250 if ncols == 1 else 60
        amat = sps.csr_matrix(amat)
        if cache is not None:
            # Picard/Newton sequences: keep the (expensive, host-side)
            # hierarchies of the first matrix for the later ones
            hierarchy = cache.get('hierarchy', hierarchy)
            vhierarchy = cache.get('vhierarchy', vhierarchy)
        Fsym = None
        if abs(amat - amat.T).max() > 1e-12*abs(amat).max():
            Fsym = (.5*(amat + amat.T)).tocsr()
            # convection: D^-1 F has complex eigenvalues, on which Chebyshev
            # polynomials of degree > 2 amplify (the V-cycle diverges)
            cheb_steps = min(cheb_steps, 2)
        self.solver, self.info = hostsetup.make_saddle_solver(
            self.ctx, amat, jmat, jmatT, nb=ncols, restart=restart,
            cheb_steps=cheb_steps, coarse_max=coarse_max,
            schur_diag=schur_diag, spectrum=spectrum, hierarchy=hierarchy,
            velocity_amg=velocity_amg, vgroups=vgroups, vhierarchy=vhierarchy,
            mass_diag=mass_diag, Fsym=Fsym, schur=schur, nsmooth=nsmooth)
        if cache is not None and self.info['velocity_amg']:
            cache.setdefault('hierarchy', self.info['hierarchy'])
            cache.setdefault('vhierarchy', self.info['vhierarchy'])

    def solve(self, rhsv, rhsp=None, x0=None, tol=1e-12, maxit=800,
              allow_unconverged=False):
        """FGMRES to the relative residual ``tol``.  The solve replaces an
        exact sparse LU: if any column stops at ``maxit`` above ``tol`` (or
        produces a NaN) `_lib.NotConverged` is raised unless
        ``allow_unconverged`` -- the iterate is then in ``self.last_vp``."""
        rhsv = np.asarray(rhsv, dtype=float).reshape(self.NV, self.ncols)
        if rhsp is not None:
            rhsp = np.asarray(rhsp, dtype=float).reshape(self.NP, self.ncols)
        vp, iters, relres = self.solver.solve(rhsv, rhsp, x0=x0, tol=tol,
                                              maxit=maxit)
        self.last_iters, self.last_relres, self.last_vp = iters, relres, vp
        bad = ~(relres <= tol)
        if bad.any() and not allow_unconverged:
            raise _lib.NotConverged(
                'FGMRES stopped after {0} iterations above tol={1:.1e} for '
                '{2} of {3} right-hand sides (largest relative residual '
                '{4:.3e})'.format(int(iters.max()), tol, int(bad.sum()),
                                  relres.size, float(np.nanmax(relres))
                                  if not np.isnan(relres).all() else
                                  float('nan')))
        return vp

    def solve_minres(self, rhsv, rhsp=None, x0=None, tol=1e-12, maxit=800):
        """MINRES for a SYMMETRIC saddle-point matrix (the Stokes / IMEX
        systems of `cylinderwake`; `north_star` K3, the reference's own
        experiment is the unpreconditioned `scipy` call at `snu:879-890`).

        The recurrence (`scipy.sparse.linalg.minres`) runs on the host; every
        product with ``K`` and every application of the preconditioner --
        ``diag(Fh^-1, Sh^-1)``, the symmetric positive definite form of the
        device's block preconditioner -- runs on the device
        (`dnsb_solver_apply_k`, `dnsb_solver_apply_prec`).  MINRES stops on
        the preconditioned residual; the true relative residual is checked
        afterwards and the iteration is continued from the iterate until it
        meets ``tol``.  Single right-hand side.  Against the device FGMRES
        (block triangular preconditioner, device resident) it needs about
        twice the iterations and a host round trip each: it is here for
        symmetric problems where a short recurrence is wanted, not for speed
        (DESIGN.md section 8)."""
        import scipy.sparse.linalg as spsla
        if self.ncols != 1:
            raise NotImplementedError('MINRES: one right-hand side at a time')
        n = self.NV + self.NP
        b = np.zeros(n)
        b[:self.NV] = np.asarray(rhsv, dtype=float).ravel()
        if rhsp is not None:
            b[self.NV:] = np.asarray(rhsp, dtype=float).ravel()
        kop = spsla.LinearOperator(
            (n, n), matvec=lambda x: self.solver.apply_k(x).ravel(),
            dtype=float)
        pop = spsla.LinearOperator(
            (n, n), matvec=lambda r: self.solver.apply_prec(r).ravel(),
            dtype=float)
        count = [0]

        def cb(xk):
            count[0] += 1
        bn = np.linalg.norm(b)
        x = np.zeros(n) if x0 is None else np.asarray(x0, float).ravel().copy()
        self.solver.set_prec_mode(True)
        try:
            rel, rtol = np.inf, tol
            for attempt in range(6):
                left = maxit - count[0]
                if left <= 0:
                    break
                x, _ = spsla.minres(kop, b, x0=x, rtol=rtol, maxiter=left,
                                    M=pop, callback=cb)
                rel = np.linalg.norm(b - kop.matvec(x))/bn if bn > 0 else 0.
                if rel <= tol:
                    break
                rtol = max(1e-3*rtol, 1e-16)
        finally:
            self.solver.set_prec_mode(False)
        self.last_iters = np.array([count[0]])
        self.last_relres = np.array([rel])
        self.last_vp = x.reshape(-1, 1)
        if not rel <= tol:
            raise _lib.NotConverged(
                'MINRES stopped after {0} iterations at a relative residual '
                'of {1:.3e} (tol {2:.1e})'.format(count[0], rel, tol))
        return x.reshape(-1, 1)

    def update_values(self, vals):
        """new values of ``amat`` on the pattern given at construction"""
        self.solver.update_fvalues(vals)

    def close(self):
        self.solver.close()


def _solve_smw(amat, jmat, rhsv, jmatT, rhsp, umat, vmat, krylov, krpslvprms,
               vgroups, mass_diag, cache):
    """``[[amat + umat vmat, jmatT], [jmat, 0]] x = b`` by Sherman-Morrison-
    Woodbury (the `smw` of the reference's `solve_sadpnt_smw`, call site
    `snu:1505-1512`): with ``K = [[amat, jmatT], [jmat, 0]]`` and the padded
    factors ``U = [umat; 0]``, ``V = [vmat, 0]``,

        x = y - Z (I + V Z)^-1 V y,    K y = b,   K Z = U,

    one batched device solve for the ``1 + rank`` columns of ``[b, U]``."""
    NP, NV = jmat.shape
    U = np.asarray(umat.todense() if sps.issparse(umat) else umat,
                   dtype=float).reshape(NV, -1)
    V = sps.csr_matrix(vmat) if sps.issparse(vmat) else \
        np.asarray(vmat, dtype=float).reshape(-1, NV)
    rhsv = np.asarray(rhsv, dtype=float).reshape(NV, -1)
    nrhs, rank = rhsv.shape[1], U.shape[1]
    bp = np.zeros((NP, nrhs)) if rhsp is None else \
        np.asarray(rhsp, dtype=float).reshape(NP, nrhs)
    op = SadpntOperator(sps.csr_matrix(amat), sps.csr_matrix(jmat), jmatT,
                        ncols=nrhs + rank, vgroups=vgroups,
                        mass_diag=mass_diag, cache=cache)
    tol = krpslvprms.get('tol', 1e-12) if krylov is not None else 1e-12
    maxit = krpslvprms.get('maxiter', 800) if krylov is not None else 800
    try:
        sol = op.solve(np.hstack([rhsv, U]),
                       np.hstack([bp, np.zeros((NP, rank))]), tol=tol,
                       maxit=maxit)
    finally:
        op.close()
    y, Z = sol[:, :nrhs], sol[:, nrhs:]
    small = np.eye(rank) + V@Z[:NV]
    return y - Z@np.linalg.solve(small, V@y[:NV])


def solve_sadpnt_smw(amat=None, jmat=None, rhsv=None, jmatT=None,
                     rhsp=None, umat=None, vmat=None, krylov=None,
                     krpslvprms={}, krplsprms={}, return_alu=False,
                     sadlu=None, decouplevp=False, solve_A=None,
                     symmetric=False, cgtol=1e-8, vgroups=None,
                     mass_diag=None, cache=None, **kw):
    """solve ``[[amat, jmatT], [jmat, 0]] [v; p] = [rhsv; rhsp]`` on the device

    Same arguments and return value as `lau.solve_sadpnt_smw` (stacked
    ``(NV+NP, k)`` array [and the reusable operator with ``return_alu``]).
    ``krpslvprms['tol'|'maxiter'|'x0']`` are honoured; ``convstatsl`` receives
    the iteration counts like krypy's convergence statistics
    (`tests/time_dep_nse_krylov.py:5,47`).  Extensions (ignored by the
    reference's signature): ``vgroups`` (node/component of the velocity
    unknowns for the vector AMG), ``mass_diag`` (diagonal of the velocity mass
    matrix for the LSC Schur approximation of stiffness dominated systems),
    ``cache`` (dict that carries the host-side hierarchies through a
    Picard/Newton sequence).  Low-rank updates ``amat + umat @ vmat`` go
    through the Sherman-Morrison-Woodbury formula on one batched device solve;
    ``decouplevp`` (a different algorithm, off the hot path) is not supported.
    """
    if (umat is None) != (vmat is None):
        raise ValueError('`umat` and `vmat` come together')
    if umat is not None:
        return _solve_smw(amat, jmat, rhsv, jmatT, rhsp, umat, vmat, krylov,
                          krpslvprms, vgroups, mass_diag, cache)
    if decouplevp:
        raise NotImplementedError('decoupled v/p solves')
    NP, NV = jmat.shape
    amat = sps.csr_matrix(amat)
    rhsv = np.asarray(rhsv, dtype=float).reshape(NV, -1)
    k = rhsv.shape[1]
    op = sadlu if sadlu is not None else \
        SadpntOperator(sps.csr_matrix(amat), sps.csr_matrix(jmat), jmatT,
                       ncols=k, vgroups=vgroups, mass_diag=mass_diag,
                       cache=cache)
    tol = krpslvprms.get('tol', 1e-12) if krylov is not None else 1e-12
    maxit = krpslvprms.get('maxiter', 800) if krylov is not None else 800
    x0 = krpslvprms.get('x0', None) if krylov is not None else None
    if krylov is not None and str(krylov).lower() == 'minres':
        if abs(amat - amat.T).max() > 1e-12*abs(amat).max():
            raise ValueError('MINRES needs a symmetric velocity block')
        vp = op.solve_minres(rhsv, rhsp, x0=x0, tol=tol, maxit=maxit)
    else:
        vp = op.solve(rhsv, rhsp, x0=x0, tol=tol, maxit=maxit)
    if 'convstatsl' in krpslvprms:
        krpslvprms['convstatsl'].append(int(op.last_iters.max()))
    if return_alu:
        return vp, op
    if sadlu is None:
        op.close()
    return vp
