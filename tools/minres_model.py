#!/usr/bin/env python
"""CPU experiment behind DESIGN.md section 8 ("MINRES"): iterations of MINRES
with the block-DIAGONAL (symmetric positive definite) form of the IMEX
preconditioner against FGMRES with the block-triangular form the library uses,
on the CNAB matrix [[M + dt/2 A, J.T], [J, 0]] of the cylinder wake.

    python tools/minres_model.py [--mesh 1] [--nts 2048]
"""
import argparse
import os
import sys

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

from dolfin_navier_scipy_b200 import hostsetup as hs          # noqa: E402
from solver_model import cheb, fgmres, BlockTriPrec           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mesh', type=int, default=1)
    ap.add_argument('--nts', type=int, default=2048)
    ap.add_argument('--cheb', type=int, default=4)
    args = ap.parse_args()
    from dolfin_navier_scipy_b200 import problem_setups as dnsps
    femp, sm, rhsd = dnsps.get_sysmats(
        problem='cylinderwake', Re=100., scheme='TH', mergerhs=True,
        meshparams=dict(refinement_level=args.mesh))
    A, M, J = sm['A'].tocsr(), sm['M'].tocsr(), sm['J'].tocsr()
    NP, NV = J.shape
    dt = 1./args.nts
    F = (M + .5*dt*A).tocsr()
    K = sps.bmat([[F, J.T], [J, None]], format='csr')
    rng = np.random.default_rng(0)
    b = np.concatenate([M@rng.standard_normal(NV) + dt*rhsd['fv'].ravel(),
                        rhsd['fp'].ravel()])
    dinv = 1./F.diagonal()
    lmin, lmax = hs.jacobi_spectrum(F)

    def vel(r):
        return cheb(F, dinv, r, args.cheb, lmin, lmax)

    # Schur complement of the two-step polynomial, dense (as on the device)
    Z2JT = np.column_stack([cheb(F, dinv, c, 2, lmin, lmax)
                            for c in J.T.toarray().T])
    S = J@Z2JT
    S = .5*(S + S.T)
    Sinv = np.linalg.inv(S)
    tri = BlockTriPrec(F, J, vel, lambda rp: Sinv@rp)
    x, its, hist = fgmres(K, b, tri, tol=1e-12, maxit=200)
    print('FGMRES, block-triangular preconditioner: %d iterations, relres %.1e'
          % (its, hist[-1]))

    def diag(r):
        return np.concatenate([vel(r[:NV]), Sinv@r[NV:]])
    count = [0]

    def cb(xk):
        count[0] += 1
    P = spsla.LinearOperator(K.shape, matvec=diag)
    bn = np.linalg.norm(b)
    for rtol in (1e-12,):
        count[0] = 0
        xm, info = spsla.minres(K, b, M=P, rtol=rtol, maxiter=2000, callback=cb)
        print('MINRES, block-diagonal preconditioner: %d iterations (info %d), '
              'true relres %.1e' % (count[0], info,
                                    np.linalg.norm(b - K@xm)/bn))


if __name__ == '__main__':
    main()
