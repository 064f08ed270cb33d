"""Oracle: `dolfin_navier_scipy/stokes_navier_utils.py` on the hot path.

TEST INFRASTRUCTURE (see `oracle/__init__.py`).

Restates `get_v_conv_conts` (`snu:40-133`), `m_innerproduct` (`snu:136-143`),
`solve_steadystate_nse` (`snu:212-545`), `solve_nse` (`snu:548-1599`: the IMEX
branch `:1100-1298` and the Picard/Newton + trapezoidal sweeps `:1304-1587`),
`get_pfromv` (`snu:1602-1633`) and the drag/lift residual functional
(`residual_checks.py:40-56`).  Scope: no Dirichlet *control* boundaries
(``diricontbcinds=None``), ``nsects=1``, no closed loop -- what the BASELINE
configs use.  Documented deviations from HEAD:
 * trajectories are kept in memory: ``lin_vel_point`` / the returned
   ``dictofvelstrs`` map ``t -> full velocity array`` instead of ``t -> path``;
 * `time.clock`, `np.float`, `np.int` fixes; dead closed-loop branch dropped;
 * the Newton/CN mode is the two-call recipe (IMEX trajectory first, then
   ``lin_vel_point=<dict>, treat_nonl_explicit=False``).
Bug-compatible quirks kept: ``get_pfromv(..., A=cmmat)`` for the initial
pressure (`snu:934`), ``p = -q/dt`` scaling, Heun corrector with ``amat=M``.
"""
import numpy as np
import scipy.sparse as sps

from . import convection as oconv
from . import tiu as otiu
from .lau import solve_sadpnt_smw


def _unroll(bcinds, bcvals):
    if bcinds is None or len(bcinds) == 0:
        return [], []
    if isinstance(bcinds[0], (list, np.ndarray)):
        ui, uv = [], []
        for k, c in enumerate(bcinds):
            ui.extend(c)
            uv.extend(bcvals[k])
        return ui, uv
    return bcinds, bcvals


def append_bcs_vec(vvec, vdim, invinds, bcinds, bcvals):
    """`dolfin_to_sparrays.py:49-64`"""
    vw = np.full((vdim, 1), np.nan)
    bi, bv = _unroll(bcinds, bcvals)
    vw[invinds] = np.asarray(vvec).reshape(-1, 1)
    if len(bi) > 0:
        vw[bi, 0] = bv
    return vw


def condense_velmat(A, invinds, dbcinds, dbcvals):
    """`dolfin_to_sparrays.py:576-642`: ``A[inv,:][:,inv]``, ``-(A u_bc)[inv]``"""
    nv = A.shape[0]
    bi, bv = _unroll(dbcinds, dbcvals)
    bcsv = np.zeros((nv, 1))
    if len(bi) > 0:
        bcsv[bi, 0] = bv
    fvbc = -A@bcsv
    return A[invinds, :][:, invinds], fvbc[invinds, :]


def get_v_conv_conts(vvec=None, V=None, invinds=None, dbcvals=[], dbcinds=[],
                     semi_explicit=False, Picard=False):
    """`snu:40-133`"""
    vvec = np.asarray(vvec)
    if len(vvec) == V.dim():
        ve = vvec.reshape(-1)
    else:
        ve = append_bcs_vec(vvec, V.dim(), invinds, dbcinds,
                            dbcvals).reshape(-1)
    if semi_explicit:
        rhs_con = oconv.convvec(V, ve)[invinds].reshape(-1, 1)
        return 0., -rhs_con, 0.
    N1, N2, rhs_con = oconv.convmats(V, ve)
    if Picard:
        convc_mat, rhsv_conbc = condense_velmat(N1, invinds, dbcinds, dbcvals)
        return convc_mat, None, rhsv_conbc
    convc_mat, rhsv_conbc = condense_velmat(N1 + N2, invinds, dbcinds, dbcvals)
    return convc_mat, rhs_con[invinds, ], rhsv_conbc


def m_innerproduct(M, v1, v2=None):
    v2 = v1 if v2 is None else v2
    return np.dot(v1.T, M@v2)


def get_pfromv(v=None, V=None, M=None, A=None, J=None, fv=None,
               stokes_flow=False, dbcinds=None, dbcvals=None, invinds=None):
    """`snu:1602-1633`: ``[M J.T; J 0][.; q] = [-A v - c(v) + fv; 0]``, p=-q"""
    if stokes_flow:
        rhs_con = 0.
    else:
        _, rhs_con, _ = get_v_conv_conts(vvec=v, V=V, invinds=invinds,
                                         dbcinds=dbcinds, dbcvals=dbcvals)
    vp = solve_sadpnt_smw(amat=M, jmat=J, jmatT=J.T, rhsv=-A@v - rhs_con + fv)
    return -vp[J.shape[1]:, :]


def solve_steadystate_nse(A=None, J=None, JT=None, M=None, fv=None, fp=None,
                          V=None, invinds=None, dbcvals=None, dbcinds=None,
                          return_vp=False, return_nwtnupd_norms=False,
                          only_stokes=False,
                          vel_pcrd_stps=10, vel_pcrd_tol=1e-4,
                          vel_nwtn_stps=20, vel_nwtn_tol=5e-15,
                          vel_start_nwtn=None, **kw):
    """`snu:212-545` (Stokes -> Picard -> Newton), SURVEY A.7"""
    JT = J.T if JT is None else JT
    cnv = A.shape[0]
    dbcinds, dbcvals = _unroll(dbcinds, dbcvals)
    norm_nwtnupd_list = []

    def _appbcs(vvec):
        return append_bcs_vec(vvec, V.dim(), invinds, dbcinds, dbcvals)

    if vel_start_nwtn is None or only_stokes:
        vp_k = solve_sadpnt_smw(amat=A, jmat=J, jmatT=JT, rhsv=fv, rhsp=fp)
        vp_k[cnv:] = -vp_k[cnv:]
        vel_k = vp_k[:cnv, ]
    else:
        vel_k = vel_start_nwtn[invinds, :]
        vp_k = np.vstack([vel_k, np.zeros((J.shape[0], 1))])

    for k in range(vel_pcrd_stps):
        if only_stokes:
            break
        N1, _, _ = oconv.convmats(V, _appbcs(vel_k).reshape(-1))
        pcrdcnvmat, rhsv_conbc = condense_velmat(N1, invinds, dbcinds, dbcvals)
        vp_k = solve_sadpnt_smw(amat=A + pcrdcnvmat, jmat=J, jmatT=JT,
                                rhsv=fv + rhsv_conbc, rhsp=fp)
        normpicupd = np.sqrt(m_innerproduct(M, vel_k - vp_k[:cnv, ]))[0]
        vel_k = vp_k[:cnv, ]
        vp_k[cnv:] = -vp_k[cnv:]
        if normpicupd < vel_pcrd_tol:
            break

    for k in range(vel_nwtn_stps):
        if only_stokes:
            break
        convc_mat, rhs_con, rhsv_conbc = \
            get_v_conv_conts(vvec=_appbcs(vel_k), V=V, invinds=invinds,
                             dbcinds=dbcinds, dbcvals=dbcvals)
        vp_k = solve_sadpnt_smw(amat=A + convc_mat, jmat=J, jmatT=JT,
                                rhsv=fv + rhs_con + rhsv_conbc, rhsp=fp)
        norm_nwtnupd = np.sqrt(m_innerproduct(M, vel_k - vp_k[:cnv, :]))[0]
        norm_nwtnupd_list.append(float(np.ravel(norm_nwtnupd)[0]))
        vel_k = vp_k[:cnv, ]
        vp_k[cnv:] = -vp_k[cnv:]
        if norm_nwtnupd < vel_nwtn_tol:
            break
    else:
        if vel_nwtn_stps > 0 and not only_stokes:
            raise UserWarning('Steady State NSE: Newton has not converged')

    vwc = _appbcs(vel_k).reshape((V.dim(), 1))
    retthing = (vwc, vp_k[cnv:, :]) if return_vp else vwc
    if return_nwtnupd_norms:
        return retthing, norm_nwtnupd_list
    return retthing


def solve_nse(A=None, M=None, J=None, JT=None, fv=None, fp=None,
              fvtd=None, fvss=0., iniv=None, inip=None, lin_vel_point=None,
              stokes_flow=False, trange=None, t0=None, tE=None, Nts=None,
              time_int_scheme='cnab', V=None, invinds=None,
              dbcinds=None, dbcvals=None,
              vel_nwtn_stps=20, vel_nwtn_tol=5e-15, vel_pcrd_stps=4,
              return_dictofvelstrs=False, treat_nonl_explicit=True,
              return_final_vp=False, return_vp_dict=False,
              check_ff=False, check_ff_maxv=1e8, start_ssstokes=False, **kw):
    """`snu:548-1599` restricted as stated in the module docstring"""
    if trange is None:
        trange = np.linspace(t0, tE, int(Nts) + 1)
    if treat_nonl_explicit and lin_vel_point is not None:
        raise UserWarning('cant use `lin_vel_point` ' +
                          'and explicit treatment of the nonlinearity')
    JT = J.T if JT is None else JT
    dbcinds, dbcvals = _unroll(dbcinds, dbcvals)
    cnv = len(invinds)
    vdim = V.dim()
    NP = J.shape[0]
    fv = np.zeros((cnv, 1)) if fv is None else fv
    fp = np.zeros((NP, 1)) if fp is None else fp
    cmmat, camat, cj, cjt, cfv, cfp = M, A, J, JT, fv, fp

    def _appbcs(vvec):
        return append_bcs_vec(vvec, vdim, invinds, dbcinds, dbcvals)

    if iniv is None:
        if not start_ssstokes:
            raise ValueError('No initial value given')
        vp_stokes = solve_sadpnt_smw(amat=camat, jmat=cj, jmatT=cjt,
                                     rhsv=cfv + fvss, rhsp=cfp)   # snu:903-907
        iniv = vp_stokes[:cnv].reshape((-1, 1))
    else:
        iniv = np.asarray(iniv).reshape(-1, 1)[invinds]            # snu:914
    if inip is None:
        inip = get_pfromv(v=iniv, V=V, M=cmmat, A=cmmat, J=cj,     # snu:934 (!)
                          fv=cfv + fvss, stokes_flow=stokes_flow,
                          dbcinds=dbcinds, dbcvals=dbcvals, invinds=invinds)

    if stokes_flow:
        vel_nwtn_stps, vel_pcrd_stps = 1, 0

    if lin_vel_point is None:          # semi-explicit integration
        def rhsv(t):
            return cfv if fvtd is None else cfv + fvtd(t)

        def rhsp(t):
            return fp

        def nonlvfunc(vvec):
            _, convvec, _ = get_v_conv_conts(vvec=vvec, V=V, invinds=invinds,
                                             semi_explicit=True)
            return convvec
        f_vdp = None if stokes_flow else nonlvfunc
        vp_dict = {}

        def _svpplz(vvec, pvec, time=None):
            vp_dict.update({time: dict(p=pvec, v=vvec)})
        timintsc = dict(cnab=otiu.cnab, sbdf2=otiu.sbdftwo)[time_int_scheme]
        v_end, p_end, ffflag = timintsc(trange=trange, inivel=iniv, inip=inip,
                                        scalep=-1., M=cmmat, A=camat, J=cj,
                                        f_vdp=f_vdp, f_tdp=rhsv, g_tdp=rhsp,
                                        appndbcs=_appbcs, savevp=_svpplz,
                                        check_ff_maxv=check_ff_maxv)

        def _flag(thing):
            return (thing, ffflag) if check_ff else thing
        if return_vp_dict:
            return _flag(vp_dict)
        elif return_final_vp:
            return _flag((v_end, p_end))
        elif return_dictofvelstrs:
            return _flag({t: d['v'] for t, d in vp_dict.items()})
        return

    # ---- Picard/Newton sweeps with the trapezoidal rule -------
    cur_linvel_point = lin_vel_point
    newtk, norm_nwtnupd = 0, 1
    loc_nwtn_tol = vel_nwtn_tol
    nwtnupd_norms = []
    while newtk < vel_nwtn_stps and norm_nwtnupd > loc_nwtn_tol:
        v_old, p_old = iniv, inip
        if vel_pcrd_stps > 0:
            vel_pcrd_stps -= 1
            pcrd_anyone = True
        else:
            pcrd_anyone = False
            newtk += 1
        dictofvelstrs = {trange[0]: _appbcs(iniv)}
        if stokes_flow:
            convc_mat_c = sps.csr_matrix((cnv, cnv))
            rhs_con_c = np.zeros((cnv, 1))
            rhsv_conbc_c = np.zeros((cnv, 1))
        else:
            convc_mat_c, rhs_con_c, rhsv_conbc_c = \
                get_v_conv_conts(vvec=_appbcs(v_old), V=V, invinds=invinds,
                                 dbcinds=dbcinds, dbcvals=dbcvals,
                                 Picard=pcrd_anyone)               # snu:1351
        _rhsconvc = 0. if pcrd_anyone else rhs_con_c
        fvn_c = cfv + rhsv_conbc_c + _rhsconvc                     # snu:1365
        norm_nwtnupd = 0
        for tk, t in enumerate(trange[1:]):
            cts = t - trange[tk]
            if stokes_flow:
                convc_mat_n = sps.csr_matrix((cnv, cnv))
                rhs_con_n = np.zeros((cnv, 1))
                rhsv_conbc_n = np.zeros((cnv, 1))
                prev_v = v_old
            else:
                try:
                    prev_v = cur_linvel_point[t]
                except KeyError:
                    prev_v = cur_linvel_point[None]
                convc_mat_n, rhs_con_n, rhsv_conbc_n = \
                    get_v_conv_conts(vvec=prev_v, V=V, invinds=invinds,
                                     dbcinds=dbcinds, dbcvals=dbcvals,
                                     Picard=pcrd_anyone)           # snu:1443
            _rhsconvn = 0. if pcrd_anyone else rhs_con_n
            fvn_n = cfv + rhsv_conbc_n + _rhsconvn                 # snu:1459
            # snu:1034-1035
            solvmat = cmmat + 0.5*cts*(camat + convc_mat_n)
            rhsv = cmmat@v_old + 0.5*cts*(fvn_n + fvn_c -
                                          (camat + convc_mat_c)@v_old)
            vp_new = solve_sadpnt_smw(amat=solvmat, jmat=cj, jmatT=cjt,
                                      rhsv=rhsv, rhsp=cfp)         # snu:1505
            v_old = vp_new[:cnv, ]
            if not stokes_flow:
                convc_mat_c, rhs_con_c, rhsv_conbc_c = \
                    get_v_conv_conts(vvec=_appbcs(v_old), V=V, invinds=invinds,
                                     dbcinds=dbcinds, dbcvals=dbcvals,
                                     Picard=pcrd_anyone)           # snu:1529
            _rhsconvc = 0. if pcrd_anyone else rhs_con_c
            fvn_c = (fvn_n - _rhsconvn - rhsv_conbc_n
                     + rhsv_conbc_c + _rhsconvc)                   # snu:1537
            dictofvelstrs[t] = _appbcs(v_old)
            p_old = -1/cts*vp_new[cnv:, ]                          # snu:1542
            if stokes_flow:
                norm_nwtnupd = None
            else:
                prev_in = prev_v[invinds, :] if len(prev_v) > cnv else prev_v
                norm_nwtnupd += float((cts*m_innerproduct(
                    cmmat, v_old - prev_in)).flatten()[0])         # snu:1559
        nwtnupd_norms.append(norm_nwtnupd)
        cur_linvel_point = dictofvelstrs
        if stokes_flow:
            break
    if return_final_vp:
        return (_appbcs(v_old), p_old)
    elif return_dictofvelstrs:
        return dictofvelstrs
    return nwtnupd_norms


def steady_state_res(Afull, JTfull, V, vfull, pfull):
    """weak residual ``A v + c(v,v) - JT p`` on *all* dofs

    (`residual_checks.py:40-56`: ``diffrm + cnvfrm - pfrm`` with the
    outflow-corrected symmetric-gradient form, which is exactly `Afull`).
    """
    vfull = np.asarray(vfull).reshape(-1)
    return Afull@vfull + oconv.convvec(V, vfull) \
        - JTfull@np.asarray(pfull).reshape(-1)


def drag_lift(Afull, JTfull, V, vfull, pfull, ldsbcinds, rho=1., L=0.1,
              Um=0.2):
    """Cd, Cl by testing the residual with the surface indicator

    (`tests/steadystate_schaefer-turek_2D-1.py:68-85`)."""
    res = steady_state_res(Afull, JTfull, V, vfull, rho*np.asarray(pfull))
    ld = np.asarray(ldsbcinds)
    drag = res[ld[ld % 2 == 0]].sum()
    lift = res[ld[ld % 2 == 1]].sum()
    fac = 2./(rho*L*Um**2)
    return fac*drag, fac*lift
