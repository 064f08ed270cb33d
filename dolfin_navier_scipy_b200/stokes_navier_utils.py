"""Navier-Stokes solvers with the reference's Python API, device internals.

Drop-in for `dolfin_navier_scipy/stokes_navier_utils.py` (`snu`):
`solve_nse` (`snu:548-1599`), `solve_steadystate_nse` (`snu:212-545`),
`get_v_conv_conts` (`snu:40-133`), `get_pfromv` (`snu:1602-1633`),
`m_innerproduct` (`snu:136-143`), `get_datastr_snu` (`snu:21-30`).

 * IMEX branch (`treat_nonl_explicit=True`, the reference's default): the whole
   time loop runs device-resident (`time_int_utils.DeviceImex`).
 * Picard/Newton sweeps and the steady-state solver: convection matrices are
   assembled by the device kernel K1b and every saddle-point system is solved
   by the device FGMRES; the (cheap) condensation and bookkeeping stay in
   Python, one host hop per linear solve.

Differences to HEAD, all documented in SURVEY.md 8c: trajectories are kept in
memory (``lin_vel_point`` / ``return_dictofvelstrs`` map ``t -> array``);
Dirichlet *control* boundaries, closed-loop feedback and paraview output are
outside the accelerated path and raise ``NotImplementedError``.
"""
import logging

import numpy as np
import scipy.sparse as sps

from . import dolfin_to_sparrays as dts
from . import lin_alg_utils as lau
from . import time_int_utils as tiu

__all__ = ['get_datastr_snu', 'get_v_conv_conts', 'solve_nse',
           'solve_steadystate_nse', 'get_pfromv', 'm_innerproduct']


def get_datastr_snu(time=None, meshp=None, nu=None, Nts=None, data_prfx='',
                    semiexpl=False):
    """file-name key of the reference's trajectory cache (`snu:21-30`)"""
    parts = [data_prfx,
             'timeNone' if time is None or isinstance(time, str)
             else 'time{0:.5e}'.format(time),
             '_nuNone' if nu is None else '_nu{0:.3e}'.format(nu),
             '_mesh{0}'.format(meshp),
             '_NtsNone' if Nts is None else '_Nts{0}'.format(Nts),
             '_semexp' if semiexpl else '']
    return ''.join(parts)


def m_innerproduct(M, v1, v2=None):
    """``v1.T M v2`` (`snu:136-143`)"""
    v2 = v1 if v2 is None else v2
    return np.dot(v1.T, M*v2)


def _reject_unsupported(diricontbcinds=None, **kw):
    if diricontbcinds not in (None, []):
        raise NotImplementedError('Dirichlet control boundaries')


def get_v_conv_conts(vvec=None, V=None,
                     invinds=None, dbcvals=[], dbcinds=[],
                     semi_explicit=False, Picard=False, retparts=False):
    """linearised convection, condensed to the inner nodes (`snu:40-133`)

    Returns ``(convc_mat, rhs_con, rhsv_conbc)``; with ``semi_explicit``
    ``(0., -N(v)v, 0.)``.  Assembly runs on the device (K1a/K1b).
    """
    vvec = np.asarray(vvec, dtype=float)
    if len(vvec) == V.dim():
        ve = vvec.reshape(-1)
    else:
        ve = dts.append_bcs_vec(vvec, V=V, invinds=invinds, bcinds=dbcinds,
                                bcvals=dbcvals).reshape(-1)
    if semi_explicit:
        rhs_con = dts.get_convvec(V=V, u0_vec=ve, invinds=invinds)
        return 0., -rhs_con, 0.
    N1, N2, rhs_con = dts.get_convmats(u0_vec=ve, V=V)
    cnd = dts.condense_velmatsbybcs
    if Picard:
        convc_mat, rhsv_conbc = cnd(N1, invinds=invinds, dbcinds=dbcinds,
                                    dbcvals=dbcvals)
        return convc_mat, None, rhsv_conbc
    elif retparts:
        pm, pr = cnd(N1, invinds=invinds, dbcinds=dbcinds, dbcvals=dbcvals)
        am, ar = cnd(N2, invinds=invinds, dbcinds=dbcinds, dbcvals=dbcvals)
        return (pm, am), rhs_con[invinds, ], (pr, ar)
    convc_mat, rhsv_conbc = cnd(N1 + N2, invinds=invinds, dbcinds=dbcinds,
                                dbcvals=dbcvals)
    return convc_mat, rhs_con[invinds, ], rhsv_conbc


def get_pfromv(v=None, V=None, M=None, A=None, J=None, fv=None, fp=None,
               decouplevp=False, solve_M=None, symmetric=False,
               cgtol=1e-8, stokes_flow=False,
               diribcs=None, dbcinds=None, dbcvals=None, invinds=None,
               **kwargs):
    """pressure that belongs to a velocity (`snu:1602-1633`)"""
    if stokes_flow:
        rhs_con = 0.
    else:
        _, rhs_con, _ = get_v_conv_conts(vvec=v, V=V, invinds=invinds,
                                         dbcinds=dbcinds, dbcvals=dbcvals)
    vp = lau.solve_sadpnt_smw(amat=M, jmat=J, jmatT=J.T,
                              rhsv=-A*v - rhs_con + fv)
    return -vp[J.shape[1]:, :]


def solve_steadystate_nse(A=None, J=None, JT=None, M=None,
                          fv=None, fp=None,
                          V=None, Q=None, invinds=None, diribcs=None,
                          dbcvals=None, dbcinds=None,
                          diricontbcinds=None, diricontbcvals=None,
                          diricontfuncs=None, diricontfuncmems=None,
                          return_vp=False, ppin=None,
                          return_nwtnupd_norms=False,
                          N=None, nu=None,
                          only_stokes=False,
                          vel_pcrd_stps=10, vel_pcrd_tol=1e-4,
                          vel_nwtn_stps=20, vel_nwtn_tol=5e-15,
                          clearprvdata=False, useolddata=False,
                          vel_start_nwtn=None, get_datastring=None,
                          data_prfx='', paraviewoutput=False,
                          save_data=False, vfileprfx='', pfileprfx='',
                          verbose=True, lin_tol=1e-12, **kw):
    """steady state by Stokes -> Picard -> Newton (`snu:212-545`)

    Same iteration as the reference; each linear system ``[[A+N, JT],[J, 0]]``
    goes to the device FGMRES (relative residual ``lin_tol``).
    """
    _reject_unsupported(diricontbcinds=diricontbcinds)
    JT = J.T if JT is None else JT
    dbcinds, dbcvals = dts.unroll_dlfn_dbcs(diribcs, bcinds=dbcinds,
                                            bcvals=dbcvals)
    cnv = A.shape[0]
    norm_nwtnupd_list = []
    # paraview files of the iterates (`snu:348-357`), written without dolfin
    prvoutdict = dict(writeoutput=False)
    if paraviewoutput:
        from . import data_output_utils as dou
        prvoutdict = dict(V=V, Q=Q, invinds=invinds, dbcinds=dbcinds,
                          dbcvals=dbcvals, ppin=ppin, writeoutput=True,
                          vfile=dou.PvdFile(vfileprfx + '__steadystates.pvd',
                                            V, name='v'),
                          pfile=dou.PvdFile(pfileprfx + '__steadystates.pvd',
                                            Q, name='p'))
    nprv = [0]

    def _prvout(vp):
        if prvoutdict['writeoutput']:
            from . import data_output_utils as dou
            dou.output_paraview(vp=vp, t=nprv[0], **prvoutdict)
            nprv[0] += 1

    def _appbcs(vvec):
        return dts.append_bcs_vec(vvec, V=V, invinds=invinds,
                                  bcinds=dbcinds, bcvals=dbcvals)

    hcache = {}
    mdiag = None if M is None else sps.csr_matrix(M).diagonal()

    warm = [None]     # last solution in the solver's sign convention

    def _solve(amat, rhsv):
        # every Picard/Newton system is started from the previous iterate: the
        # updates shrink from step to step, and so does the work of the
        # iterative solve (the stopping test is relative to |rhs| either way)
        prms = dict(tol=lin_tol, maxiter=2000)
        if warm[0] is not None:
            prms['x0'] = warm[0]
        vp = lau.solve_sadpnt_smw(amat=amat, jmat=J, jmatT=JT, rhsv=rhsv,
                                  rhsp=fp, krylov='gmres',
                                  vgroups=(np.asarray(invinds)//2,
                                           np.asarray(invinds) % 2),
                                  mass_diag=mdiag, cache=hcache,
                                  krpslvprms=prms)
        warm[0] = vp.copy()
        return vp

    if vel_start_nwtn is None or only_stokes:
        vp_k = _solve(A, fv)
        vp_k[cnv:] = -vp_k[cnv:]            # p was flipped for symmetry
        vel_k = vp_k[:cnv, ]
        _prvout(vp_k)                       # `snu:412-416`
    else:
        vel_k = vel_start_nwtn[invinds, :]
        vp_k = np.vstack([vel_k, np.zeros((J.shape[0], 1))])

    for k in range(vel_pcrd_stps):
        if only_stokes:
            break
        N1, _, _ = dts.get_convmats(u0_vec=_appbcs(vel_k), V=V)
        pcrdcnvmat, rhsv_conbc = dts.condense_velmatsbybcs(
            N1, invinds=invinds, dbcinds=dbcinds, dbcvals=dbcvals)
        vp_k = _solve(A + pcrdcnvmat, fv + rhsv_conbc)
        normpicupd = np.sqrt(m_innerproduct(M, vel_k - vp_k[:cnv, ]))[0]
        if verbose:
            logging.info('Picard iteration: {0} -- norm of update: {1}'.
                         format(k+1, normpicupd))
        vel_k = vp_k[:cnv, ]
        vp_k[cnv:] = -vp_k[cnv:]
        _prvout(vp_k)                       # `snu:472-473`
        if normpicupd < vel_pcrd_tol:
            break

    for vel_newtk in range(vel_nwtn_stps):
        if only_stokes:
            break
        convc_mat, rhs_con, rhsv_conbc = get_v_conv_conts(
            vvec=_appbcs(vel_k), V=V, invinds=invinds, dbcinds=dbcinds,
            dbcvals=dbcvals)
        vp_k = _solve(A + convc_mat, fv + rhs_con + rhsv_conbc)
        norm_nwtnupd = np.sqrt(m_innerproduct(M, vel_k - vp_k[:cnv, :]))[0]
        norm_nwtnupd_list.append(float(np.ravel(norm_nwtnupd)[0]))
        vel_k = vp_k[:cnv, ]
        vp_k[cnv:] = -vp_k[cnv:]
        _prvout(vp_k)                       # `snu:514-516`
        if verbose:
            logging.info('Steady State NSE: Newton iteration: {0} -- norm of '
                         'update: {1}'.format(vel_newtk, norm_nwtnupd))
        # the linear systems are solved to a relative residual `lin_tol`, not
        # by LU: an update at that level is converged (the reference's default
        # `vel_nwtn_tol=5e-15` is below what an iterative solve can resolve)
        floor = 100.*lin_tol*np.sqrt(m_innerproduct(M, vel_k))[0]
        if norm_nwtnupd < max(vel_nwtn_tol, float(np.ravel(floor)[0])):
            break
    else:
        if vel_nwtn_stps == 0 or only_stokes:
            pass
        else:
            raise UserWarning('Steady State NSE: Newton has not converged')

    vwc = _appbcs(vel_k).reshape((V.dim(), 1))
    retthing = (vwc, vp_k[cnv:, :]) if return_vp else vwc
    if return_nwtnupd_norms:
        return retthing, norm_nwtnupd_list
    return retthing


def solve_nse(A=None, M=None, J=None, JT=None,
              fv=None, fp=None,
              fvtd=None, fvss=0.,
              fvtvd=None, fv_tmdp=None,
              iniv=None, inip=None, lin_vel_point=None,
              stokes_flow=False,
              trange=None,
              t0=None, tE=None, Nts=None,
              time_int_scheme='cnab',
              V=None, Q=None, invinds=None, diribcs=None,
              dbcinds=None, dbcvals=None,
              diricontbcinds=None, diricontbcvals=None,
              diricontfuncs=None, diricontfuncmems=None,
              N=None, nu=None,
              ppin=None,
              closed_loop=False,
              static_feedback=False,
              dynamic_feedback=False, dyn_fb_dict={},
              dyn_fb_disc='trapezoidal',
              b_mat=None,
              vel_nwtn_stps=20, vel_nwtn_tol=5e-15,
              nsects=1, loc_nwtn_tol=5e-15, loc_pcrd_stps=True,
              addfullsweep=False,
              vel_pcrd_stps=4,
              krylov=None, krpslvprms={}, krplsprms={},
              clearprvdata=False,
              get_datastring=None,
              data_prfx='',
              paraviewoutput=False, plttrange=None, prvoutpnts=None,
              vfileprfx='', pfileprfx='',
              return_dictofvelstrs=False,
              return_dictofpstrs=False,
              treat_nonl_explicit=True, no_data_caching=True,
              use_custom_nonlinearity=False,
              custom_nonlinear_vel_function=None,
              datatrange=None, dataoutpnts=None,
              return_final_vp=False,
              return_vp_dict=False,
              return_y_list=False, cv_mat=None,
              check_ff=False, check_ff_maxv=1e8,
              verbose=True,
              start_ssstokes=False,
              lin_tol=1e-12, guess=16, cheb_steps=4,
              **kw):
    """time-dependent Navier-Stokes -- `snu:548-1599`

    ``M v' + A v + N(v) v + J.T p = f_v,  J v = f_p``.  With
    ``treat_nonl_explicit`` (default, like the reference) the IMEX scheme
    ``time_int_scheme`` in {'cnab', 'sbdf2'} runs device-resident; otherwise
    Picard/Newton sweeps with the trapezoidal rule about ``lin_vel_point``
    (``{t: full velocity array}``, e.g. the ``return_dictofvelstrs`` result of
    an IMEX run).  Extra keywords: ``lin_tol`` (FGMRES relative residual),
    ``guess`` (initial-guess mode of ``dnsb_imex_run``), ``cheb_steps``.
    """
    _reject_unsupported(diricontbcinds=diricontbcinds)
    if fv_tmdp is not None:
        raise DeprecationWarning()
    if nsects != 1 or addfullsweep:
        # dead at HEAD: the second section looks its linearisation point up in
        # the dict of the first one and dies with `KeyError: None`
        # (`snu:1425-1431`; tests/test_reference_pin.py::
        # test_live_reference_time_sections_are_broken_at_head)
        raise NotImplementedError(
            'time sections (`nsects` > 1, `addfullsweep`) fail in the '
            'reference itself (`KeyError` at `stokes_navier_utils.py:1428`)')
    # argument combinations are checked BEFORE any device work
    if lin_vel_point is None:
        if time_int_scheme not in ('cnab', 'sbdf2'):
            raise ValueError("`time_int_scheme` must be 'cnab' or 'sbdf2' "
                             "(`snu:1199-1202`), got {0!r}".
                             format(time_int_scheme))
        if not treat_nonl_explicit:
            raise NotImplementedError(
                'IMEX -> Newton hand-off inside one call is broken at HEAD '
                '(SURVEY.md 8c): run the IMEX integration with '
                '`return_dictofvelstrs=True` first and call again with '
                '`lin_vel_point=<that dict>, treat_nonl_explicit=False`')
        if stokes_flow:
            raise NotImplementedError('stokes_flow in the IMEX branch')
    if trange is None:
        trange = np.linspace(t0, tE, int(Nts) + 1)
    trange = np.asarray(trange, dtype=float)
    if treat_nonl_explicit and lin_vel_point is not None:
        raise UserWarning('cant use `lin_vel_point` ' +
                          'and explicit treatment of the nonlinearity')
    JT = J.T if JT is None else JT
    dbcinds, dbcvals = dts.unroll_dlfn_dbcs(diribcs, bcinds=dbcinds,
                                            bcvals=dbcvals)
    invinds = np.asarray(invinds)
    cnv = invinds.size
    NP = J.shape[0]
    fv = np.zeros((cnv, 1)) if fv is None else np.asarray(fv).reshape(cnv, 1)
    fp = np.zeros((NP, 1)) if fp is None else np.asarray(fp).reshape(NP, 1)

    def _appbcs(vvec):
        return dts.append_bcs_vec(vvec, V=V, invinds=invinds,
                                  bcinds=dbcinds, bcvals=dbcvals)

    def _paraview_trajectory(vdict, pdict):
        """`vfile << v, t` for the times in `plttrange` (all if None) --
        `snu:817-821,1091-1098,1169-1196`, files `<prfx>__timestep.pvd`"""
        from . import data_output_utils as dou
        tfilter = None if plttrange is None else [float(t) for t in plttrange]
        vfile = dou.PvdFile(vfileprfx + '__timestep.pvd', V, name='v')
        pfile = dou.PvdFile(pfileprfx + '__timestep.pvd', Q, name='p')
        for t in sorted(vdict.keys()):
            dou.output_paraview(V=V, Q=Q, vc=vdict[t], pc=pdict.get(t), t=t,
                                tfilter=tfilter, ppin=ppin, vfile=vfile,
                                pfile=pfile)

    vgroups = (invinds//2, invinds % 2)
    krydict = dict(krylov='gmres', vgroups=vgroups,
                   mass_diag=sps.csr_matrix(M).diagonal(),
                   krpslvprms=dict(tol=lin_tol*1e-1, maxiter=2000))
    # ---- initial value (`snu:836-940`) --------------------------------------
    if iniv is None:
        if not start_ssstokes:
            raise ValueError('No initial value given')
        logging.info('computing the Stokes-Solution for initial value')
        vp_stokes = lau.solve_sadpnt_smw(amat=A, jmat=J, jmatT=JT,
                                         rhsv=fv + fvss, rhsp=fp, **krydict)
        iniv = vp_stokes[:cnv].reshape((-1, 1))
    else:
        iniv = np.asarray(iniv).reshape(-1, 1)[invinds]
    if inip is None:
        # NB: the reference passes the mass matrix as `A` here (`snu:934`)
        inip = get_pfromv(v=iniv, V=V, M=M, A=M, J=J, fv=fv + fvss, fp=fp,
                          stokes_flow=stokes_flow, dbcinds=dbcinds,
                          dbcvals=dbcvals, invinds=invinds)
    if stokes_flow:
        vel_nwtn_stps, vel_pcrd_stps = 1, 0

    if lin_vel_point is None:       # ---- semi-explicit integration ---------
        if fvtd is None:
            f_tdp = None
            fvc = fv
        else:
            def f_tdp(t):
                return fv + np.asarray(fvtd(t)).reshape(cnv, 1)
            fvc = None
        vp_dict = {}
        want_traj = (return_vp_dict or return_dictofvelstrs or return_y_list
                     or paraviewoutput)

        def _svpplz(vvec, pvec, time=None):
            vp_dict.update({float(time): dict(p=pvec, v=vvec)})
        # ---- opaque callbacks: per-step host hop (`snu:1129-1140,1207-1247`) --
        if fvtvd is not None or use_custom_nonlinearity or closed_loop:
            if use_custom_nonlinearity:
                def nonlvfunc(vfull):      # minus: it goes to the rhs (`snu:1131`)
                    return -custom_nonlinear_vel_function(vfull)
            else:
                def nonlvfunc(vfull):
                    return -dts.get_convvec(V=V, u0_vec=vfull, invinds=invinds)
            dynamic_rhs = None
            if closed_loop and dynamic_feedback:
                if dyn_fb_disc != 'AB2':
                    # `snu:1212-1222` builds `implicit_dynamic_rhs` for the
                    # trapezoidal observer but never hands it to the integrator
                    raise NotImplementedError(
                        "dyn_fb_disc={0!r}: only 'AB2' reaches the integrator "
                        "at HEAD (`snu:1224-1234`)".format(dyn_fb_disc))
                dfb = dyn_fb_dict
                observer = tiu.get_heunab_lti(hb=dfb['hb'], ha=dfb['ha'],
                                              hc=dfb['hc'], inihx=dfb['inihx'],
                                              drift=dfb['drift'])

                def dynamic_rhs(t, vc=None, memory={}, mode=None):
                    curu, memory = observer(t, vc=cv_mat.dot(vc),
                                            memory=memory, mode=mode)
                    return b_mat.dot(curu), memory
            elif closed_loop and static_feedback:
                raise NotImplementedError('static feedback enters through the '
                                          'Newton sweeps (`snu:1367-1383`)')
            timintsc = dict(cnab=tiu.cnab, sbdf2=tiu.sbdftwo)[time_int_scheme]
            v_end, p_end, ffflag = timintsc(
                trange=trange, inivel=iniv, inip=inip, M=M, A=A, J=J,
                scalep=-1., f_vdp=nonlvfunc, f_tvdp=fvtvd,
                f_tdp=(lambda t: fv) if fvtd is None else
                (lambda t: fv + np.asarray(fvtd(t)).reshape(cnv, 1)),
                g_tdp=lambda t: fp, dynamic_rhs=dynamic_rhs,
                appndbcs=lambda vvec, bcs: _appbcs(vvec), savevp=_svpplz,
                check_ff_maxv=check_ff_maxv, tol=lin_tol)
            f_tdp = fvc = None
            hosthop_done = True
        else:
            hosthop_done = False
        scheme = time_int_scheme
        if hosthop_done:
            pass
        elif f_tdp is None:
            # constant rhs: hand it over as `fv` (no sampling needed)
            integ = tiu.DeviceImex(M, A, J, V, invinds, dbcinds, dbcvals,
                                   trange[1] - trange[0], scheme=scheme,
                                   nus=(1.,), fv=fvc, fp=fp,
                                   cheb_steps=cheb_steps)
            integ.set_state(iniv, inip)
            ffflag = integ.run(trange.size - 1,
                               snap_stride=1 if want_traj else 0,
                               tol=lin_tol, guess=guess,
                               check_ff_maxv=check_ff_maxv)
            v_end, p_end = integ.state()
            if want_traj:
                vs, ps = integ.snapshots()
                for k in range(vs.shape[0]):
                    _svpplz(_appbcs(vs[k, :, :1]), ps[k, :, :1],
                            time=trange[k])
            integ.close()
        else:
            v_end, p_end, ffflag = tiu._run_imex(
                scheme, trange=trange, inivel=iniv, inip=inip, M=M, A=A, J=J,
                f_tdp=f_tdp, g_tdp=lambda t: fp, V=V, invinds=invinds,
                dbcinds=dbcinds, dbcvals=dbcvals,
                savevp=_svpplz if want_traj else None,
                check_ff_maxv=check_ff_maxv, tol=lin_tol, guess=guess,
                cheb_steps=cheb_steps)

        if paraviewoutput:
            _paraview_trajectory({t: d['v'] for t, d in vp_dict.items()},
                                 {t: d['p'] for t, d in vp_dict.items()})

        def _flag(thing):
            return (thing, ffflag) if check_ff else thing
        if return_vp_dict:
            return _flag(vp_dict)
        elif return_final_vp:
            return _flag((v_end, p_end))
        elif return_dictofvelstrs:
            return _flag({t: d['v'] for t, d in vp_dict.items()})
        elif return_y_list:
            ylist = [d['v'] if cv_mat is None else cv_mat.dot(d['v'][invinds])
                     for t, d in sorted(vp_dict.items())]
            return _flag(ylist)
        return

    # ---- Picard / Newton sweeps with the trapezoidal rule (`snu:1304-1587`) --
    # The reference builds `M + dt/2 (A + N_n)` and factorises it anew every
    # step (`snu:1484-1512`).  Here all step matrices live on ONE pattern (the
    # union of M, A and the condensed P2 convection pattern): per step the
    # convection values come from the device kernel K1b, are gathered into that
    # pattern, and only the VALUES of the device matrix are replaced
    # (`dnsb_solver_update_fvalues`); the preconditioner (Schur approximation,
    # Chebyshev bounds) is set up once per sweep matrix family.
    from . import _lib
    from .time_int_utils import _on_pattern, _union_pattern
    cur_linvel_point = lin_vel_point
    newtk, norm_nwtnupd = 0, 1
    nwtnupd_norms = []
    M, A = sps.csr_matrix(M), sps.csr_matrix(A)
    dev = _lib.device_for(V)
    cip, cix = dev.pattern
    nvf = V.dim()
    # condensed convection pattern; its data are the (1-based) slots of the
    # full device pattern
    Pc = sps.csr_matrix((np.arange(1, cix.size + 1, dtype=float), cix, cip),
                        shape=(nvf, nvf))[invinds, :][:, invinds].tocsr()
    Pc.sort_indices()
    src = Pc.data.astype(np.int64) - 1
    pat = _union_pattern([M, A, Pc])
    pos = _on_pattern(sps.csr_matrix((np.arange(1, Pc.nnz + 1, dtype=float),
                                      Pc.indices, Pc.indptr), shape=Pc.shape),
                      pat)
    convpos = np.nonzero(pos.data)[0]
    assert np.array_equal(pos.data[convpos], np.arange(1, Pc.nnz + 1))
    Mv, Av = _on_pattern(M, pat).data, _on_pattern(A, pat).data
    ubc = np.zeros((nvf, 1))
    if len(dbcinds) > 0:
        ubc[np.asarray(dbcinds), 0] = dbcvals
    zerov = np.zeros((cnv, 1))

    def _csr(vals):
        return sps.csr_matrix((vals, pat.indices, pat.indptr), shape=pat.shape)

    def _convconts(vfull, picard):
        """values of the condensed convection matrix on `pat`, N(v)v[inv] (0
        for Picard), -(N u_bc)[inv]  -- `snu:40-133`"""
        if stokes_flow:
            return np.zeros(pat.nnz), zerov, zerov
        n1, n2, f3 = dev.convmats(np.asarray(vfull, dtype=float).reshape(-1))
        nfull = n1 if picard else n1 + n2
        vals = np.zeros(pat.nnz)
        vals[convpos] = nfull[src]
        Nfull = sps.csr_matrix((nfull, cix, cip), shape=(nvf, nvf))
        rbc = -(Nfull@ubc)[invinds, :]
        rc = zerov if picard else f3.reshape(-1, 1)[invinds, :]
        return vals, rc, rbc

    def _lookup(pnt, t):
        try:
            return pnt[t]
        except KeyError:
            return pnt[None]

    # ---- the sweeps: device resident (`dnsb_cnsweep_run`) ---------------------
    # per sweep the host uploads the linearisation trajectory and receives the
    # new one; assembly (K1b), matrix update, right-hand side, FGMRES and the
    # update norm of every step stay on the device.
    dts_all = np.diff(trange)
    nsteps = dts_all.size
    uniq = {}
    for i, v in zip(dbcinds, dbcvals):
        uniq[int(i)] = float(v)
    bci = np.array(sorted(uniq.keys()), dtype=np.int32)
    bcv = np.array([uniq[i] for i in bci], dtype=float)
    op = sweep = None
    if not stokes_flow:
        dt0 = float(dts_all[0])
        nv0, _, _ = _convconts(_appbcs(iniv), vel_pcrd_stps > 0)
        op = lau.SadpntOperator(_csr(Mv + 0.5*dt0*(Av + nv0)), J, JT, ncols=1,
                                velocity_amg=False, schur='lumped',
                                cheb_steps=min(cheb_steps, 3))
        mmat = op.ctx.csr(M)
        dev.bind()
        sweep = _lib.CnSweep(op.solver, mmat, Mv, Av, src, convpos, invinds,
                             bci, bcv, fv, fp, nvf)
        sweep.set_guess(krpslvprms.get('krylovini', None))
    while newtk < vel_nwtn_stps and norm_nwtnupd > vel_nwtn_tol:
        if vel_pcrd_stps > 0:
            vel_pcrd_stps -= 1
            pcrd_anyone = True
        else:
            pcrd_anyone = False
            newtk += 1
        if stokes_flow:
            # linear problem: one "sweep" without convection, host-driven
            v_old, p_old = iniv, inip
            dictofvelstrs = {float(trange[0]): _appbcs(iniv)}
            dictofpstrs = {float(trange[0]): inip}
            x0 = None
            for tk, t in enumerate(trange[1:]):
                cts = t - trange[tk]
                solvvals = Mv + 0.5*cts*Av
                rhsv = M@v_old + 0.5*cts*(2*fv - A@v_old)
                if op is None:
                    op = lau.SadpntOperator(_csr(solvvals), J, JT, ncols=1,
                                            velocity_amg=False, schur='lumped',
                                            cheb_steps=min(cheb_steps, 3))
                else:
                    op.update_values(solvvals)
                x0 = np.vstack([v_old, -cts*p_old]) if x0 is None else x0
                x0 = op.solve(rhsv, fp, x0=x0, tol=lin_tol, maxit=2000)
                v_old, p_old = x0[:cnv, ], -1/cts*x0[cnv:, ]
                dictofvelstrs[float(t)] = _appbcs(v_old)
                dictofpstrs[float(t)] = p_old
            norm_nwtnupd = None
            nwtnupd_norms.append(norm_nwtnupd)
            break
        lin = np.vstack([np.asarray(_lookup(cur_linvel_point, float(t))).
                         reshape(1, -1) if
                         np.asarray(_lookup(cur_linvel_point, float(t))).size
                         == nvf else
                         _appbcs(_lookup(cur_linvel_point, float(t))).
                         reshape(1, -1) for t in trange])
        dev.bind()
        vtraj, ptraj, norm_nwtnupd, its = sweep.run(
            dts_all, lin, iniv, inip, pcrd_anyone, tol=lin_tol, maxit=2000)
        if 'convstatsl' in krpslvprms:
            krpslvprms['convstatsl'].append(its/float(nsteps))
        dictofvelstrs = {float(t): vtraj[k].reshape(-1, 1)
                         for k, t in enumerate(trange)}
        dictofpstrs = {float(t): ptraj[k].reshape(-1, 1)
                       for k, t in enumerate(trange)}
        dictofpstrs[float(trange[0])] = inip
        v_old = vtraj[-1].reshape(-1, 1)[invinds, :]
        p_old = ptraj[-1].reshape(-1, 1)
        nwtnupd_norms.append(norm_nwtnupd)
        if verbose:
            print('norm of current Newton update: {}'.format(norm_nwtnupd))
        cur_linvel_point = dictofvelstrs
    if sweep is not None:
        sweep.close()
    if op is not None:
        op.close()
    if paraviewoutput:            # the last sweep (`snu:1563-1566`)
        _paraview_trajectory(dictofvelstrs, dictofpstrs)

    if return_final_vp:
        return (_appbcs(v_old), p_old)
    elif return_dictofvelstrs:
        if return_dictofpstrs:
            return dictofvelstrs, dictofpstrs
        return dictofvelstrs
    return
