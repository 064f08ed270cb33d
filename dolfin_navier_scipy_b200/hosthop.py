"""IMEX integration with opaque Python callbacks: the per-step host hop.

The device-resident loop (``dnsb_imex_run``) knows its right-hand side: the P2
convection, a low-rank time series of forcings, constant boundary data.  The
reference's integrators accept more -- arbitrary ``f_vdp(v)``, ``f_tvdp(t, v)``,
``dynamic_rhs(t, vc=, memory=, mode=)`` (observers / output feedback,
`tiu:148-257`), time dependent ``g_tdp`` and Dirichlet control through
``getbcs`` / ``applybcs`` (`tiu:23-145,260-355,366-477`).  Those cannot run
inside a CUDA loop, so this module runs the loop on the host and keeps the
heavy parts on the device ("state-dependent callbacks via a
per-step D2H -> callback -> H2D hop"):

 * every saddle-point solve: ``lin_alg_utils.SadpntOperator`` (device FGMRES),
   one operator per distinct matrix (IMEX-Euler predictor, mass-matrix
   corrector, loop matrix), each warm-started from its previous solution;
 * whatever the callbacks themselves call (``dts.get_convvec`` evaluates the
   convection on the device).

The arithmetic of a step is the reference's (SURVEY A.4/A.5 with the boundary
and dynamic terms), written once for both multistep schemes.
"""
import numpy as np
import scipy.sparse as sps

from . import lin_alg_utils as lau

__all__ = ['imex_with_callbacks', 'euler_with_callbacks']


def _timeslices(times, nslices):
    """`tiu:480-489`: ``nslices`` equal chunks and the remainder"""
    times = list(times)
    n = len(times)//nslices
    chunks = [times[k*n:(k + 1)*n] for k in range(nslices)]
    chunks.append(times[nslices*n:])
    return chunks


class _WarmSolver(object):
    """device solve of ``[[F, J.T], [J, 0]]`` that remembers its last solution"""

    def __init__(self, F, J, ctx=None, tol=1e-12, maxit=800):
        self.op = lau.SadpntOperator(sps.csr_matrix(F), sps.csr_matrix(J),
                                     sps.csr_matrix(J).T.tocsr(), ncols=1,
                                     ctx=ctx)
        self.tol, self.maxit, self.last = tol, maxit, None
        self.nv = F.shape[0]

    def __call__(self, rhsv, rhsp):
        vp = self.op.solve(rhsv, rhsp, x0=self.last, tol=self.tol,
                           maxit=self.maxit)
        self.last = vp.copy()
        return vp[:self.nv].reshape(-1, 1), vp[self.nv:].reshape(-1, 1)

    def close(self):
        self.op.close()


def imex_with_callbacks(scheme, trange=None, inivel=None, inip=None,
                        bcs_ini=[], M=None, A=None, J=None, f_vdp=None,
                        f_tdp=None, g_tdp=None, f_tvdp=None, scalep=-1.,
                        getbcs=None, applybcs=None, appndbcs=None,
                        savevp=None, dynamic_rhs=None, dynamic_rhs_memory={},
                        check_ff_maxv=None, ntimeslices=10, ctx=None,
                        tol=1e-12, maxit=800, **kw):
    """CNAB (``scheme='cnab'``, `tiu:23-145`) or SBDF2 (``'sbdf2'``,
    `tiu:260-355`) after the Heun start (`tiu:366-477`) with the reference's
    callback interface; returns ``(v_n, p_n, ffflag)``."""
    trange = np.asarray(trange, dtype=float)
    steps = np.diff(trange)
    if not np.allclose(np.linalg.norm(np.diff(steps)), 0):
        raise NotImplementedError()                      # `tiu:358-363`
    dt = trange[1] - trange[0]
    M, A, J = sps.csr_matrix(M), sps.csr_matrix(A), sps.csr_matrix(J)
    NP, NV = J.shape
    zero = np.zeros((NV, 1))
    theta = dict(cnab=.5, sbdf2=2./3)[scheme]

    # ---- defaults of the optional callbacks ---------------------------------
    conv = f_vdp if f_vdp is not None else (lambda vfull: zero)
    frc = f_tdp if f_tdp is not None else (lambda t: zero)
    div = g_tdp if g_tdp is not None else (lambda t: np.zeros((NP, 1)))
    if getbcs is None:
        def getbcs(time, vvec, pvec, mode=None):
            return []
    if applybcs is None:
        def applybcs(bcs):
            return 0., 0., 0.
    if appndbcs is None:
        def appndbcs(vvec, bcs):
            return vvec
    if savevp is None:
        def savevp(vvec, pvec, time=None):
            return
    if dynamic_rhs is None:
        def dynamic_rhs(t, vc=None, memory={}, mode=None):
            return zero, memory
    if f_tvdp is not None:
        plain_dyn = dynamic_rhs

        def dynamic_rhs(t, vc=None, memory={}, mode=None):     # `tiu:52-58`
            val, memory = plain_dyn(t, vc=vc, memory=memory, mode=mode)
            return val + f_tvdp(t, vc), memory
    vmax = np.inf if check_ff_maxv is None else check_ff_maxv

    solve_pred = _WarmSolver(M + dt*A, J, ctx=ctx, tol=tol, maxit=maxit)
    solve_corr = _WarmSolver(M, J, ctx=ctx, tol=tol, maxit=maxit)
    solve_loop = _WarmSolver(M + theta*dt*A, J, ctx=ctx, tol=tol, maxit=maxit)
    try:
        # ---- t0 ----------------------------------------------------------------
        d_c, mem = dynamic_rhs(trange[0], vc=inivel,
                               memory=dynamic_rhs_memory, mode='init')
        savevp(appndbcs(inivel, bcs_ini), inip, time=trange[0])
        # ---- Heun start: IMEX-Euler predictor, trapezoidal corrector -----------
        t0, t1 = trange[0], trange[1]
        b_c, _, mb_c = applybcs(bcs_ini)
        f_c, n_c = frc(t0), conv(appndbcs(inivel, bcs_ini))
        d_pred, mem = dynamic_rhs(t1, vc=inivel, memory=mem, mode='heunpred')
        bcs_pred = getbcs(t1, appndbcs(inivel, bcs_ini), inip, mode='heunpred')
        b_pred, bp_pred, mb_pred = applybcs(bcs_pred)
        f_n, g_n = frc(t1), div(t1)
        v_pred, q_pred = solve_pred(
            M@inivel + dt*(f_n + b_pred + d_pred) + dt*n_c - (mb_pred - mb_c),
            g_n + bp_pred)
        p_pred = scalep/dt*q_pred
        d_n, mem = dynamic_rhs(t1, vc=v_pred, memory=mem, mode='heuncorr')
        n_pred = conv(appndbcs(v_pred, bcs_pred))
        bcs_n = getbcs(t1, appndbcs(v_pred, bcs_pred), p_pred, mode='heuncorr')
        b_n, bp_n, mb_n = applybcs(bcs_n)
        v_n, q_n = solve_corr(
            M@inivel - (mb_n - mb_c) - .5*dt*(A@(inivel + v_pred))
            + .5*dt*(f_c + f_n + b_n + b_c + d_n + d_c + n_c + n_pred),
            g_n + bp_n)
        p_n = scalep/dt*q_n
        savevp(appndbcs(v_n, bcs_n), p_n, time=t1)
        # state of the multistep loops: (current) <- (new), (old) <- (current)
        v_c, n_o = inivel, n_c           # n_o: convection one level back
        ffflag = 0
        for chunk in _timeslices(trange[2:], ntimeslices):
            guard = np.linalg.norm(v_n if scheme == 'cnab' else v_c)
            if guard > vmax or np.isnan(guard):                   # `tiu:99-103,311-319`
                ffflag = 1
                break
            for t in chunk:
                v_o, mb_o = v_c, mb_c
                v_c, p_c, bcs_c, b_c, mb_c = v_n, p_n, bcs_n, b_n, mb_n
                f_c, d_c = f_n, d_n
                n_c = conv(appndbcs(v_c, bcs_c))
                bcs_n = getbcs(t, appndbcs(v_c, bcs_c), p_c, mode='abtwo')
                b_n, bp_n, mb_n = applybcs(bcs_n)
                f_n, g_n = frc(t), div(t)
                d_n, mem = dynamic_rhs(t, vc=v_c, memory=mem, mode='abtwo')
                if scheme == 'cnab':                              # `tiu:125-128`
                    rhs = M@v_c - .5*dt*(A@v_c) - (mb_n - mb_c) \
                        + .5*dt*(3*n_c - n_o) \
                        + .5*dt*(f_c + f_n + b_n + b_c + d_n + d_c)
                else:                                             # `tiu:342-346`
                    rhs = 1./3*(M@(4*v_c - v_o)) \
                        - (mb_n - 4./3*mb_c + 1./3*mb_o) + 2./3*dt*b_n \
                        + 2./3*dt*(2*n_c - n_o) + 2./3*dt*(f_n + d_n)
                v_n, q_n = solve_loop(rhs, g_n + bp_n)
                p_n = scalep/dt*q_n
                n_o = n_c
                savevp(appndbcs(v_n, bcs_n), p_n, time=t)
    finally:
        for s in (solve_pred, solve_corr, solve_loop):
            s.close()
    return v_n, p_n, ffflag


def euler_with_callbacks(iniv=None, jmat=None, mmat=None, amat=None, rhsv=None,
                         trange=None, data_trange=None, fp=None, ctx=None,
                         tol=1e-12, maxit=800):
    """`tiu.semi_implicit_euler` (`tiu:566-635`) with its opaque ``rhsv(t, v)``:
    ``(M + dt A) v+ + J.T q = M v + dt rhsv(t+, v)``, ``J v+ = fp``"""
    trange = np.asarray(trange, dtype=float)
    wanted = list(trange if data_trange is None else data_trange)[1:]
    mmat, amat = sps.csr_matrix(mmat), sps.csr_matrix(amat)
    NP, NV = jmat.shape
    g = np.zeros((NP, 1)) if fp is None else np.asarray(fp).reshape(NP, 1)
    dt = trange[1] - trange[0]
    step = _WarmSolver(mmat + dt*amat, jmat, ctx=ctx, tol=tol, maxit=maxit)
    out, v = [iniv], np.asarray(iniv, dtype=float).reshape(NV, 1)
    try:
        for t in trange[1:]:
            v, _ = step(mmat@v + dt*np.asarray(rhsv(t, v)).reshape(NV, 1), g)
            if wanted and t == wanted[0]:
                out.append(v)
                wanted.pop(0)
    finally:
        step.close()
    return out
