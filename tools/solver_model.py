#!/usr/bin/env python
"""numpy model of the DEVICE saddle-point solver (development tool, CPU only).

Right-preconditioned FGMRES with the block-triangular preconditioner of
`libdnsb200` (`apply_prec` in csrc/dnsb_api.cu) rebuilt from the SAME host-side
hierarchies (`hostsetup.sa_amg_hierarchy`, `lumped_schur`): Jacobi-Chebyshev
smoothing, smoothed-aggregation V-cycle on the velocity block, least-squares
commutator / PCD Schur approximations.  Used to study iteration counts of
preconditioner variants without a GPU (the device kernels are checked against
the same restatement in tests/test_gpu_parity.py::
test_preconditioner_matches_numpy_restatement).

    python tools/solver_model.py [--mesh 1] [--Re 60] [--schur lsc|pcd] ...
"""
import argparse
import os
import sys

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dolfin_navier_scipy_b200 import hostsetup as hs          # noqa: E402


def cheb(F, dinv, r, k, lmin, lmax):
    th, de = .5*(lmax + lmin), .5*(lmax - lmin)
    sigma = th/de
    rho = 1./sigma
    z = np.zeros_like(r)
    res = r.copy()
    d = dinv*res/th
    for i in range(k):
        z = z + d
        if i == k - 1:
            break
        res = res - F@d
        rho_n = 1./(2*sigma - rho)
        d = rho_n*rho*d + 2*rho_n/de*(dinv*res)
        rho = rho_n
    return z


class VCycle(object):
    """V-cycle on `levels` (dicts A, P, R, lmin, lmax) + dense coarse inverse"""

    def __init__(self, levels, dense, nsmooth=2, smoother='cheb', omega=0.6):
        self.levels, self.dense = levels, dense
        self.nsmooth, self.smoother, self.omega = nsmooth, smoother, omega
        self.dinv = [1./lv['A'].diagonal() for lv in levels]

    def smooth(self, l, b, x0=None):
        lv = self.levels[l]
        A, dinv = lv['A'], self.dinv[l]
        if self.smoother == 'cheb':
            if x0 is None:
                return cheb(A, dinv, b, self.nsmooth, lv['lmin'], lv['lmax'])
            return x0 + cheb(A, dinv, b - A@x0, self.nsmooth, lv['lmin'],
                             lv['lmax'])
        x = np.zeros_like(b) if x0 is None else x0.copy()
        for _ in range(self.nsmooth):          # damped Jacobi
            x = x + self.omega*dinv*(b - A@x)
        return x

    def __call__(self, b, l=0):
        if l == len(self.levels):
            return self.dense@b
        lv = self.levels[l]
        x = self.smooth(l, b)
        rc = lv['R']@(b - lv['A']@x)
        x = x + lv['P']@self(rc, l + 1)
        return self.smooth(l, b, x)


def fgmres(K, b, prec, tol=1e-12, maxit=400, restart=250, x0=None):
    n = b.size
    x = np.zeros(n) if x0 is None else x0.copy()
    bn = np.linalg.norm(b)
    hist = []
    total = 0
    while total < maxit:
        r = b - K@x
        beta = np.linalg.norm(r)
        hist.append(beta/bn)
        if beta <= tol*bn:
            break
        V = [r/beta]
        Z = []
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        j = 0
        for j in range(restart):
            z = prec(V[j])
            w = K@z
            for _ in range(2):                  # CGS2
                h = np.array([v@w for v in V])
                w = w - sum(hi*v for hi, v in zip(h, V))
                H[:j+1, j] += h
            hn = np.linalg.norm(w)
            H[j+1, j] = hn
            Z.append(z)
            V.append(w/hn if hn > 0 else w)
            for i in range(j):
                t = cs[i]*H[i, j] + sn[i]*H[i+1, j]
                H[i+1, j] = -sn[i]*H[i, j] + cs[i]*H[i+1, j]
                H[i, j] = t
            dd = np.hypot(H[j, j], H[j+1, j])
            cs[j], sn[j] = H[j, j]/dd, H[j+1, j]/dd
            H[j, j] = dd
            g[j+1] = -sn[j]*g[j]
            g[j] = cs[j]*g[j]
            total += 1
            hist.append(abs(g[j+1])/bn)
            if abs(g[j+1]) <= tol*bn or total >= maxit:
                break
        y = np.linalg.solve(np.triu(H[:j+1, :j+1]), g[:j+1])
        x = x + sum(yi*zi for yi, zi in zip(y, Z))
        if abs(g[j+1]) <= tol*bn:
            r = b - K@x
            hist.append(np.linalg.norm(r)/bn)
            break
    return x, total, hist


class BlockTriPrec(object):
    """z = P^-1 r,  P = [Fh JT; 0 -Sh]"""

    def __init__(self, F, J, velocity, schur):
        self.F, self.J, self.JT = F, J, J.T.tocsr()
        self.velocity, self.schur = velocity, schur
        self.NV = F.shape[0]

    def __call__(self, r):
        rv, rp = r[:self.NV], r[self.NV:]
        zp = -self.schur(rp)
        zv = self.velocity(rv - self.JT@zp)
        return np.concatenate([zv, zp])


def lsc_schur(F, J, du, coarse_max=4096):
    """Sh^-1 = L^-1 (J Du^-1 F Du^-1 JT) L^-1, L = J Du^-1 JT (dense inverse)"""
    S = hs.lumped_schur(du, J)
    levels, dense = hs.sa_amg_hierarchy(S, coarse_max=coarse_max)
    Lsolve = VCycle(levels, dense) if levels else (lambda b: dense@b)
    JT = J.T.tocsr()
    dui = 1./du

    def apply(rp):
        t = Lsolve(rp)
        t = dui*(JT@t)
        t = dui*(F@t)
        return Lsolve(J@t)
    return apply


def pcd_schur(Fp, Ap_inv, Mp_dinv):
    """pressure convection-diffusion (Kay, Loghin, Wathen 2002):
    Sh^-1 = Mp^-1 Fp Ap^-1 (lumped pressure mass)"""
    def apply(rp):
        return Mp_dinv*(Fp@(Ap_inv(rp)))
    return apply


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mesh', type=int, default=1)
    ap.add_argument('--Re', type=float, default=60.)
    ap.add_argument('--cheb', type=int, default=2)
    ap.add_argument('--nsmooth', type=int, default=2)
    ap.add_argument('--smoother', default='cheb')
    ap.add_argument('--omega', type=float, default=.6)
    ap.add_argument('--maxit', type=int, default=600)
    ap.add_argument('--picard', type=int, default=1)
    args = ap.parse_args()
    from dolfin_navier_scipy_b200 import problem_setups as dnsps
    from oracle import convection as oconv
    from oracle import snu as osnu
    from oracle.lau import solve_sadpnt_smw as olu
    femp, sm, rhsd = dnsps.get_sysmats(
        problem='cylinderwake', Re=args.Re, scheme='TH', mergerhs=True,
        meshparams=dict(refinement_level=args.mesh))
    A, M, J = sm['A'].tocsr(), sm['M'].tocsr(), sm['J'].tocsr()
    NP, NV = J.shape
    inv = np.asarray(femp['invinds'])
    vp = olu(amat=A, jmat=J, jmatT=J.T, rhsv=rhsd['fv'], rhsp=rhsd['fp'])
    for it in range(args.picard):
        vfull = osnu.append_bcs_vec(vp[:NV], femp['V'].dim(), inv,
                                    femp['dbcinds'], femp['dbcvals'])
        N1, _, _ = oconv.convmats(femp['V'], vfull.ravel())
        Nc, rbc = osnu.condense_velmat(N1, inv, femp['dbcinds'],
                                       femp['dbcvals'])
        F = (A + Nc).tocsr()
        K = sps.bmat([[F, J.T], [J, None]], format='csr')
        b = np.concatenate([(rhsd['fv'] + rbc).ravel(), rhsd['fp'].ravel()])
        Fsym = (.5*(F + F.T)).tocsr()
        vlevels, vdense = hs.sa_amg_hierarchy(
            F, coarse_max=2048, groups=(inv//2, inv % 2), Asym=Fsym)
        lmin, lmax = hs.jacobi_spectrum(F)
        print('levels', len(vlevels), 'jacobi spectrum', lmin, lmax,
              'coarse', vdense.shape)
        vc = VCycle(vlevels, vdense, nsmooth=args.cheb,
                    smoother=args.smoother, omega=args.omega)
        if vlevels:
            vlevels[0]['lmin'], vlevels[0]['lmax'] = lmax/10., lmax
        schur = lsc_schur(F, J, M.diagonal())
        prec = BlockTriPrec(F, J, vc, schur)
        x, its, hist = fgmres(K, b, prec, maxit=args.maxit)
        ref = spsla.spsolve(K.tocsc(), b)
        print('picard', it, 'iters', its, 'final relres', hist[-1],
              'err', np.linalg.norm(x - ref)/np.linalg.norm(ref))
        print('  history', ['%.1e' % h for h in hist[::max(1, len(hist)//12)]])
        vp = ref.reshape(-1, 1)


if __name__ == '__main__':
    main()
