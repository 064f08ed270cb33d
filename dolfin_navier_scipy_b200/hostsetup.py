"""Host-side, once-per-run setup of the device solvers.

Everything here is set-up work (spectral bounds for the Chebyshev smoothers,
the Schur-complement approximation, smoothed-aggregation AMG hierarchies,
dense coarse inverses) -- the per-step arithmetic runs in ``libdnsb200``.
The reference has no counterpart: it factorises the saddle-point matrix with
SuperLU (`time_int_utils.py:89-91`, `stokes_navier_utils.py:1505-1512`).
"""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spsla

from . import _lib

__all__ = ['jacobi_spectrum', 'lumped_schur', 'sa_amg_hierarchy',
           'make_saddle_solver']


def jacobi_spectrum(F, its=30, seed=0, ratio=None):
    """bounds ``(lmin, lmax)`` of the spectrum of ``D^-1 F`` (F ~ SPD)

    ``its`` Lanczos steps (full reorthogonalisation) on the symmetric part of
    ``D^-1/2 F D^-1/2``: ``lmax`` = largest Ritz value (+5 %), ``lmin`` = 0.9 x
    smallest Ritz value (an upper bound of the true one: fine for the
    well-conditioned, mass dominated matrices; for ill-conditioned ones the
    caller switches to multigrid and passes ``ratio``: ``lmin = lmax/ratio``).
    """
    F = sps.csr_matrix(F)
    d = np.abs(F.diagonal())
    ds = 1./np.sqrt(d)
    Fs = .5*(F + F.T) if ratio is None else F

    def op(x):
        return ds*(Fs@(ds*x))
    n = F.shape[0]
    rng = np.random.default_rng(seed)
    q = rng.standard_normal(n)
    q /= np.linalg.norm(q)
    its = min(its, n)
    Q = np.zeros((its, n))
    al, be = np.zeros(its), np.zeros(its)
    k = 0
    for k in range(its):
        Q[k] = q
        w = op(q)
        al[k] = q@w
        w -= Q[:k+1].T@(Q[:k+1]@w)
        w -= Q[:k+1].T@(Q[:k+1]@w)
        be[k] = np.linalg.norm(w)
        if be[k] < 1e-12*abs(al[0]):
            break
        q = w/be[k]
    T = np.diag(al[:k+1]) + np.diag(be[:k], 1) + np.diag(be[:k], -1)
    ev = np.linalg.eigvalsh(T)
    lmax = 1.05*float(ev[-1])
    if ratio is not None:
        return lmax/ratio, lmax
    return 0.9*float(max(ev[0], 1e-12*lmax)), lmax


def lumped_schur(fdiag, J):
    """``S = J diag(F)^-1 J.T`` -- the sparse Schur-complement approximation"""
    J = sps.csr_matrix(J)
    S = (J@sps.diags(1./np.asarray(fdiag))@J.T).tocsr()
    S.sort_indices()
    return S


def _strength(A, theta):
    """symmetric strength graph |a_ij| >= theta*sqrt(a_ii a_jj), i != j"""
    A = sps.coo_matrix(A)
    d = np.abs(sps.csr_matrix(A).diagonal())
    mask = (A.row != A.col) & \
        (np.abs(A.data) >= theta*np.sqrt(d[A.row]*d[A.col]))
    G = sps.csr_matrix((np.ones(int(mask.sum())),
                        (A.row[mask], A.col[mask])), shape=A.shape)
    return G


def _aggregate(G):
    """greedy aggregation (root + neighbours, then attach the leftovers)"""
    n = G.shape[0]
    indptr, indices = G.indptr, G.indices
    agg = -np.ones(n, dtype=np.int64)
    nagg = 0
    for i in range(n):
        if agg[i] >= 0:
            continue
        nbrs = indices[indptr[i]:indptr[i+1]]
        if np.all(agg[nbrs] < 0):
            agg[i] = nagg
            agg[nbrs] = nagg
            nagg += 1
    for i in range(n):
        if agg[i] >= 0:
            continue
        nbrs = indices[indptr[i]:indptr[i+1]]
        nbrs = nbrs[agg[nbrs] >= 0]
        if nbrs.size > 0:
            agg[i] = agg[nbrs[0]]
        else:
            agg[i] = nagg
            nagg += 1
    return agg, nagg


def sa_amg_hierarchy(A, coarse_max=4096, theta=0.08, max_levels=10,
                     groups=None, Asym=None):
    """smoothed-aggregation AMG hierarchy (set up on the host)

    ``groups`` (optional, per row): ``(node, comp)`` arrays for vector
    problems -- rows of one node are aggregated together and every component
    gets its own coarse column (unknown-based SA).  ``Asym``: matrix used for
    strength/aggregation and prolongator smoothing (default ``A``; pass the
    symmetric part for convection-diffusion operators), Galerkin products use
    ``A`` itself.  Returns ``(levels, dense_inv)``: ``levels`` = list of dicts
    ``A, P, R, lmin, lmax`` for all but the coarsest level.
    """
    A = sps.csr_matrix(A)
    As = A if Asym is None else sps.csr_matrix(Asym)
    levels = []
    while A.shape[0] > coarse_max and len(levels) < max_levels:
        n = A.shape[0]
        if groups is None:
            node = np.arange(n)
            comp = np.zeros(n, dtype=np.int64)
            ncomp = 1
        else:
            node, comp = groups
            ncomp = int(comp.max()) + 1
        unodes, nidx = np.unique(node, return_inverse=True)
        nn = unodes.size
        # node-condensed strength matrix
        Q = sps.csr_matrix((np.ones(n), (np.arange(n), nidx)), shape=(n, nn))
        An = (Q.T@abs(As)@Q).tocsr()
        G = _strength(An, theta)
        agg, nagg = _aggregate(G)
        if nagg >= 0.8*nn:
            break
        cols = agg[nidx]*ncomp + comp
        T = sps.csr_matrix((np.ones(n), (np.arange(n), cols)),
                           shape=(n, nagg*ncomp))
        keep = np.asarray(T.sum(axis=0)).ravel() > 0
        T = T[:, keep]
        cn = np.sqrt(np.asarray(T.multiply(T).sum(axis=0)).ravel())
        T = (T@sps.diags(1./cn)).tocsr()
        dinv = 1./As.diagonal()
        _, lmax_s = jacobi_spectrum(As, ratio=10.)
        P = (T - (4./(3.*lmax_s/1.05))*(sps.diags(dinv)@(As@T))).tocsr()
        R = P.T.tocsr()
        for mm in (P, R):
            mm.sort_indices()
        _, lmax = jacobi_spectrum(A, ratio=10.)
        levels.append(dict(A=A, P=P, R=R, lmin=lmax/10., lmax=lmax))
        A = (R@A@P).tocsr()
        A.sort_indices()
        As = A if Asym is None else (R@As@P).tocsr()
        if groups is not None:
            newcols = np.nonzero(keep)[0]
            groups = (newcols // ncomp, newcols % ncomp)
    Ad = A.toarray()
    sym = np.allclose(Ad, Ad.T, rtol=1e-10, atol=1e-14*np.abs(Ad).max())
    dense_inv = np.linalg.pinv(Ad, hermitian=True) if sym \
        else np.linalg.pinv(Ad)
    return levels, dense_inv


def make_saddle_solver(ctx, F1, J, JT=None, F2=None, coef=None, nb=1,
                       restart=40, cheb_steps=3, schur='lumped',
                       schur_diag=None, coarse_max=4096, mp_diag=None,
                       mp_scale=None, spectrum=None, hierarchy=None,
                       nsmooth=2, velocity_amg='auto', vgroups=None,
                       vhierarchy=None, Fsym=None, vcoarse_max=2048):
    """build a device ``SaddleSolver`` for ``[[F1 + coef_m*F2, JT], [J, 0]]``

    Schur approximation: ``schur='lumped'``: AMG/dense inverse of
    ``J diag^-1 JT`` (``schur_diag`` defaults to the diagonal of the mean
    member matrix) plus, optionally, the Cahouet-Chabard mass term
    ``mp_scale_m * diag(mp_diag)^-1``; ``schur='mass'``: that mass term alone
    (Stokes/Oseen).  ``velocity_amg``: smoothed-aggregation V-cycle for the
    velocity block (needed when F is not mass dominated; ``'auto'`` decides by
    the condition number of the Jacobi-scaled block).  Host-side
    hierarchies can be shared between solvers via ``hierarchy``/``vhierarchy``.
    Returns ``(solver, info)``.
    """
    F1 = sps.csr_matrix(F1)
    F1.sort_indices()
    J = sps.csr_matrix(J)
    JT = J.T.tocsr() if JT is None else sps.csr_matrix(JT)
    cm = 0. if coef is None else float(np.mean(coef))
    if F2 is not None:
        F2 = sps.csr_matrix(F2)
        F2.sort_indices()
        if not (np.array_equal(F1.indptr, F2.indptr) and
                np.array_equal(F1.indices, F2.indices)):
            raise ValueError('F1 and F2 must share one CSR pattern')
        Fmean = sps.csr_matrix((F1.data + cm*F2.data, F1.indices, F1.indptr),
                               shape=F1.shape)
        cmax = float(np.max(coef))
        Fext = sps.csr_matrix((F1.data + cmax*F2.data, F1.indices, F1.indptr),
                              shape=F1.shape)
    else:
        Fmean = Fext = F1
    if spectrum is None or velocity_amg == 'auto':
        lmin, lmax = jacobi_spectrum(Fext)
        lmin2, lmax2 = jacobi_spectrum(Fmean)
        spec = (min(lmin, lmin2), max(lmax, lmax2))
        if velocity_amg == 'auto':
            # mass dominated matrices have cond(D^-1 F) ~ 5; Stokes/Oseen
            # blocks are Laplace-like and need the V-cycle
            velocity_amg = spec[1]/spec[0] > 50.
        if spectrum is None:
            spectrum = (spec[1]/10., spec[1]) if velocity_amg else spec
    fmat = ctx.csr(F1, None if F2 is None else F2.data)
    jmat, jtmat = ctx.csr(J), ctx.csr(JT)
    solver = _lib.SaddleSolver(ctx, fmat, jmat, jtmat, coef=coef, nb=nb,
                               restart=restart, cheb_steps=cheb_steps,
                               lmin=spectrum[0], lmax=spectrum[1])
    keep = [fmat, jmat, jtmat]
    nlev = 0
    if schur == 'lumped':
        if hierarchy is None:
            sd = Fmean.diagonal() if schur_diag is None else schur_diag
            S = lumped_schur(sd, J)
            hierarchy = sa_amg_hierarchy(S, coarse_max=coarse_max)
        levels, dense_inv = hierarchy
        for lv in levels:
            a, p, r = ctx.csr(lv['A']), ctx.csr(lv['P']), ctx.csr(lv['R'])
            keep += [a, p, r]
            solver.add_schur_level(a, p, r, nsmooth=nsmooth, lmin=lv['lmin'],
                                   lmax=lv['lmax'])
        solver.add_schur_level(dense_inv=dense_inv)
        nlev = len(levels) + 1
    elif schur != 'mass':
        raise ValueError('schur must be `lumped` or `mass`')
    if mp_diag is not None and mp_scale is not None:
        solver.set_schur_mass(1./np.asarray(mp_diag),
                              np.broadcast_to(np.asarray(mp_scale, float),
                                              (nb,)))
    elif schur == 'mass':
        raise ValueError('schur=`mass` needs mp_diag and mp_scale')
    if velocity_amg:
        if vhierarchy is None:
            vhierarchy = sa_amg_hierarchy(Fmean, coarse_max=vcoarse_max,
                                          groups=vgroups, Asym=Fsym)
        vlevels, vdense = vhierarchy
        if len(vlevels) > 0:
            p0, r0 = ctx.csr(vlevels[0]['P']), ctx.csr(vlevels[0]['R'])
            keep += [p0, r0]
            solver.set_velocity_transfer(p0, r0)
            for lv in vlevels[1:]:
                a, p, r = ctx.csr(lv['A']), ctx.csr(lv['P']), ctx.csr(lv['R'])
                keep += [a, p, r]
                solver.add_velocity_level(a, p, r, nsmooth=nsmooth,
                                          lmin=lv['lmin'], lmax=lv['lmax'])
            solver.add_velocity_level(dense_inv=vdense)
    solver._keepalive = keep
    info = dict(spectrum=spectrum, hierarchy=hierarchy, vhierarchy=vhierarchy,
                schur_levels=nlev, velocity_amg=bool(velocity_amg))
    return solver, info
