"""Oracle: IMEX integrators of `dolfin_navier_scipy/time_int_utils.py`.

TEST INFRASTRUCTURE (see `oracle/__init__.py`).

Restates `cnab` (`tiu:23-145`), `sbdftwo` (`tiu:260-355`), `_onestepheun`
(`tiu:366-477`, default scheme 'IMEX-Euler', corrector solved with
``amat=M``), `_inittimegrid` (`tiu:480-489`) and `semi_implicit_euler`
(`tiu:566-635`) for the case exercised by the BASELINE configs: no
time-varying Dirichlet control (``applybcs == 0``, which is also what the
reference computes when controls *are* present because `snu:1112` is
commented out) and no dynamic (observer) right hand side.  Formulas:
SURVEY.md Appendix A.4/A.5.
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

from .lau import solve_sadpnt_smw


def inittimegrid(trange, ntimeslices=10):
    """uniform-grid check + 10 slices of ``trange[2:]`` (`tiu:358-363,480-489`)"""
    tr = np.array(trange)
    dtvec = tr[1:] - tr[:-1]
    if not np.allclose(np.linalg.norm(dtvec[1:] - dtvec[:-1]), 0):
        raise NotImplementedError()
    dt = trange[1] - trange[0]
    lltr = np.array(trange[2:])
    lenofts = np.floor(lltr.size/ntimeslices).astype(np.int32)
    listofts = [lltr[k*lenofts: (k+1)*lenofts].tolist()
                for k in range(ntimeslices)]
    listofts.append(lltr[ntimeslices*lenofts:].tolist())
    return dt, listofts


def onestepheun(vc, tc, tn, M, A, J, scalep, f_tdp, f_vdp, g_tdp, appndbcs):
    """start-up step, `tiu:366-477` with ``scheme='IMEX-Euler'``"""
    NP, NV = J.shape
    dt = tn - tc
    fv_c = f_tdp(tc)
    nfc_c = f_vdp(appndbcs(vc))
    fv_n, fp_n = f_tdp(tn), g_tdp(tn)
    # predictor: implicit Euler for diffusion, explicit for convection
    tfv = M@vc + dt*fv_n + dt*nfc_c                                # tiu:399-401
    tvp_n = solve_sadpnt_smw(amat=M + dt*A, jmat=J, jmatT=J.T,
                             rhsv=tfv, rhsp=fp_n)
    tv_n = tvp_n[:NV, :]
    # corrector: trapezoidal everything, mass-matrix solve
    tnfc_n = f_vdp(appndbcs(tv_n))
    rhs_n = M@vc - .5*dt*(A@(vc + tv_n)) \
        + .5*dt*(fv_c + fv_n + nfc_c + tnfc_n)                     # tiu:459-461
    vp_n = solve_sadpnt_smw(amat=M, jmat=J, jmatT=J.T,
                            rhsv=rhs_n, rhsp=fp_n)
    v_n = vp_n[:NV].reshape((NV, 1))
    p_n = 1./dt*scalep*vp_n[NV:].reshape((NP, 1))
    nfc_n = f_vdp(appndbcs(v_n))
    return v_n, p_n, fv_n, nfc_c, nfc_n


def _saddle_lu(F, J):
    NP = J.shape[0]
    K = sps.vstack([sps.hstack([F, J.T]),
                    sps.hstack([J, sps.csr_matrix((NP, NP))])], format='csc')
    return spsla.factorized(K)


def cnab(trange=None, inivel=None, inip=None, M=None, A=None, J=None,
         f_vdp=None, f_tdp=None, g_tdp=None, scalep=-1.,
         appndbcs=None, savevp=None, check_ff_maxv=1e8, ntimeslices=10):
    """Crank-Nicolson / Adams-Bashforth-2 -- `tiu:23-145`, SURVEY A.5"""
    dt, listofts = inittimegrid(trange, ntimeslices=ntimeslices)
    NP, NV = J.shape
    zerorhs = np.zeros((NV, 1))
    ffflag = 0
    if f_vdp is None:
        def f_vdp(vvec):
            return zerorhs
    savevp(appndbcs(inivel), inip, time=trange[0])
    v_n, p_n, fv_n, nfc_c, nfc_n = \
        onestepheun(inivel, trange[0], trange[1], M, A, J, scalep,
                    f_tdp, f_vdp, g_tdp, appndbcs)
    savevp(appndbcs(v_n), p_n, time=trange[1])
    coeffmatlu = _saddle_lu(M + .5*dt*A, J)                        # tiu:89-91
    for kck, ctrange in enumerate(listofts):
        nrmvc = np.linalg.norm(v_n)
        if nrmvc > check_ff_maxv or np.isnan(nrmvc):               # tiu:99-103
            ffflag = 1
            break
        for ctime in ctrange:
            v_c, fv_c = v_n, fv_n
            nfc_o = nfc_c
            nfc_c = f_vdp(appndbcs(v_c))
            fv_n, fp_n = f_tdp(ctime), g_tdp(ctime)
            rhs_n = M@v_c - .5*dt*(A@v_c) \
                + .5*dt*(3*nfc_c - nfc_o) + .5*dt*(fv_c + fv_n)    # tiu:125-128
            vp_n = coeffmatlu(np.vstack([rhs_n, fp_n]).flatten())
            v_n = vp_n[:NV].reshape((NV, 1))
            p_n = 1./dt*scalep*vp_n[NV:].reshape((NP, 1))
            savevp(appndbcs(v_n), p_n, time=ctime)
    return v_n, p_n, ffflag


def sbdftwo(trange=None, inivel=None, inip=None, M=None, A=None, J=None,
            f_vdp=None, f_tdp=None, g_tdp=None, scalep=-1.,
            appndbcs=None, savevp=None, check_ff_maxv=1e8, ntimeslices=10):
    """SBDF2 -- `tiu:260-355`"""
    dt, listofts = inittimegrid(trange, ntimeslices=ntimeslices)
    NP, NV = J.shape
    zerorhs = np.zeros((NV, 1))
    if f_vdp is None:
        def f_vdp(vvec):
            return zerorhs
    savevp(appndbcs(inivel), inip, time=trange[0])
    v_c = inivel
    v_n, p_n, fv_n, nfc_c, nfc_n = \
        onestepheun(v_c, trange[0], trange[1], M, A, J, scalep,
                    f_tdp, f_vdp, g_tdp, appndbcs)
    savevp(appndbcs(v_n), p_n, time=trange[1])
    coeffmatlu = _saddle_lu(M + 2./3*dt*A, J)                      # tiu:304-306
    ffflag = 0
    for kck, ctrange in enumerate(listofts):
        nrmvc = np.linalg.norm(v_c)                                # tiu:311 (v_c!)
        if nrmvc > check_ff_maxv or np.isnan(nrmvc):
            ffflag = 1
            break
        for ctime in ctrange:
            v_p = v_c
            v_c = v_n
            nfc_p = nfc_c
            nfc_c = f_vdp(appndbcs(v_c))
            fv_n, fp_n = f_tdp(ctime), g_tdp(ctime)
            rhs_n = 1/3*M@(4*v_c - v_p) \
                + 2/3*dt*(2*nfc_c - nfc_p) + 2/3*dt*fv_n           # tiu:342-346
            vp_n = coeffmatlu(np.vstack([rhs_n, fp_n]).flatten())
            v_n = vp_n[:NV].reshape((NV, 1))
            p_n = 1./dt*scalep*vp_n[NV:].reshape((NP, 1))
            savevp(appndbcs(v_n), p_n, time=ctime)
    return v_n, p_n, ffflag


def semi_implicit_euler(iniv=None, jmat=None, mmat=None, amat=None, rhsv=None,
                        trange=None, data_trange=None, fp=None):
    """IMEX Euler with a prefactorised ``M+dt*A`` -- `tiu:566-635`

    ``(M+dt A) v+ + J.T q = M v + dt*rhsv(t+, v)``, ``J v+ = fp`` (`tiu:608-615`);
    returns the velocities at ``data_trange`` (default: all of ``trange``).
    """
    dtpt = trange if data_trange is None else data_trange
    pending = np.copy(dtpt).tolist()
    pending.pop(0)
    NP, NV = jmat.shape
    fpz = np.zeros((NP, 1)) if fp is None else fp
    dt = trange[1] - trange[0]
    lu = _saddle_lu(mmat + dt*amat, jmat)
    ievlist = [iniv]
    cvn = iniv
    for ct in trange[1:]:
        rhs = (mmat@cvn).reshape((-1, 1)) + dt*rhsv(ct, cvn)
        cvn = lu(np.vstack([rhs, fpz]).flatten())[:NV].reshape((NV, 1))
        if len(pending) > 0 and ct == pending[0]:
            ievlist.append(cvn)
            pending.pop(0)
    return ievlist
