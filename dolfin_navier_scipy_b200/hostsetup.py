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

__all__ = ['jacobi_spectrum', 'lumped_schur', 'poly_schur', 'sa_amg_hierarchy',
           'make_saddle_solver', 'hilbert_key', 'locality_perm',
           'pad_row_pairs']


def hilbert_key(xy, order=16):
    """index of 2D points along a Hilbert curve (vectorised)"""
    xy = np.asarray(xy, dtype=float)
    span = max(np.ptp(xy[:, 0]), np.ptp(xy[:, 1]), 1e-300)
    top = 2**order - 1
    x = ((xy[:, 0] - xy[:, 0].min())/span*top).astype(np.int64)
    y = ((xy[:, 1] - xy[:, 1].min())/span*top).astype(np.int64)
    d = np.zeros_like(x)
    s = 2**(order - 1)
    while s > 0:
        rx = (x & s) > 0
        ry = (y & s) > 0
        d += s*s*((3*rx) ^ ry)
        flip = (~ry) & rx
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        x, y = np.where(~ry, y, x), np.where(~ry, x, y)
        x &= (s - 1)
        y &= (s - 1)
        s //= 2
    return d


def locality_perm(coords=None, comp=None, pattern=None):
    """ordering of the unknowns that keeps mesh neighbours close in memory

    The batched SpMM kernels give every CTA a chunk of consecutive rows and
    rely on the gathered ``x`` rows being shared inside the chunk (L1/shared
    memory reuse); a space-filling-curve numbering makes a chunk a compact
    patch of the mesh.  ``coords`` (n, 2): Hilbert order, ties broken by
    ``comp`` (keeps the components of a node adjacent); without coordinates:
    reverse Cuthill-McKee of ``pattern``.  Returns ``perm`` with
    ``new[i] = old[perm[i]]``.
    """
    if coords is not None:
        key = hilbert_key(coords)
        comp = np.zeros(len(key), dtype=np.int64) if comp is None else comp
        # by curve index, then by node (two nodes may share a curve cell), then
        # by component: the components of a node stay adjacent, which the
        # row-pair SpMM kernels rely on
        _, node = np.unique(np.round(np.asarray(coords), 14), axis=0,
                            return_inverse=True)
        return np.lexsort((np.asarray(comp), node.ravel(), key)).astype(np.int64)
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    pat = sps.csr_matrix(pattern)
    return np.asarray(reverse_cuthill_mckee(pat, symmetric_mode=True),
                      dtype=np.int64)


def pad_row_pairs(P):
    """pattern of ``P`` with the rows (2k, 2k+1) given the union of their
    column lists (explicit zeros)

    In the device numbering the two velocity components of a node are
    adjacent rows; structurally they couple to the same columns, but assembly
    drops entries that happen to be exactly zero.  With identical lists the
    row-pair SpMM kernels load every gathered ``x`` once for both rows.
    """
    P = sps.csr_matrix(P)
    n = P.shape[0]
    ones = sps.csr_matrix((np.ones(P.nnz), P.indices, P.indptr), shape=P.shape)
    if n % 2:
        return ones
    swap = np.arange(n).reshape(-1, 2)[:, ::-1].ravel()
    out = (ones + ones[swap, :]).tocsr()
    out.data[:] = 1.
    out.sort_indices()
    return out


def jacobi_spectrum(F, its=30, seed=0, ratio=None):
    """bounds ``(lmin, lmax)`` of the spectrum of ``D^-1 F`` (F ~ SPD)

    ``its`` Lanczos steps (full reorthogonalisation) on the symmetric part of
    ``D^-1/2 F D^-1/2``: ``lmax`` = largest Ritz value (+5 %), ``lmin`` = 1.6 x
    smallest Ritz value (see the note at the end of the function; for
    ill-conditioned matrices the caller switches to multigrid and passes
    ``ratio``: ``lmin = lmax/ratio``).
    """
    F = sps.csr_matrix(F)
    d = np.abs(F.diagonal())
    ds = 1./np.sqrt(d)
    nonsym = abs(F - F.T).max() > 1e-12*abs(F).max()
    Fs = (.5*(F + F.T)).tocsr() if nonsym else F

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
    if nonsym:
        # convection: the Chebyshev smoother must stay stable on the field of
        # values of D^-1 F, not only on its eigenvalues -- with the bound of
        # the symmetric part the V-cycle stalls FGMRES on the Oseen systems
        # (DFG 2D-1: 2.6 fails, 3.7 and the Gershgorin bound 6.8 converge)
        gersh = float((abs(sps.diags(1./d)@F)).sum(axis=1).max())
        lmax = min(gersh, 1.5*lmax)
    if ratio is not None:
        return lmax/ratio, lmax
    # The Chebyshev iteration is a PRECONDITIONER inside FGMRES, not a solver:
    # a lower bound ABOVE the smallest eigenvalue concentrates the polynomial
    # on the bulk of the spectrum and leaves the few lowest modes to the Krylov
    # method.  Measured on the ensemble bench (cylinder_4, dt = 1/2048):
    # factor 0.4 / 0.9 / 1.0 / 1.7 / 3.0 -> 8.9 / 6.5 / 6.2 / 5.7 / 6.2
    # FGMRES iterations per step.
    import os
    fac = float(os.environ.get('DNSB_LMIN_FACTOR', '1.6'))
    return fac*float(max(ev[0], 1e-12*lmax)), lmax


def lumped_schur(fdiag, J):
    """``S = J diag(F)^-1 J.T`` -- the sparse Schur-complement approximation"""
    J = sps.csr_matrix(J)
    S = (J@sps.diags(1./np.asarray(fdiag))@J.T).tocsr()
    S.sort_indices()
    return S


def poly_schur(F, J, k=2, spectrum=None):
    """``S = J Z_k J.T`` with ``Z_k`` the k-term Jacobi-Chebyshev polynomial
    approximation of ``F^-1`` (``k=1``: the lumped ``J diag(F)^-1 J.T``)

    ``Z_k`` is built explicitly (sparse, pattern of ``F^(k-1)``); ``S`` stays
    sparse (35 / 57 entries per row for k = 2 / 3 on the cylinder meshes) and is
    a much closer approximation of the Schur complement ``J F^-1 J.T`` of the
    mass dominated time-stepping matrices than the lumped one: with 5 Chebyshev
    steps on the velocity block the FGMRES needs 16 instead of 22 iterations
    (cylinder_2, dt = 1/2048), the exact Schur complement 14.
    """
    F = sps.csr_matrix(F)
    J = sps.csr_matrix(J)
    n = F.shape[0]
    dinv = sps.diags(1./F.diagonal())
    if k <= 1:
        Z = dinv
    else:
        lmin, lmax = jacobi_spectrum(F) if spectrum is None else spectrum
        th, de = .5*(lmax + lmin), .5*(lmax - lmin)
        sigma = th/de
        rho = 1./sigma
        R = sps.identity(n, format='csr')
        Dd = (dinv@R/th).tocsr()
        Z = sps.csr_matrix((n, n))
        for i in range(k):
            Z = Z + Dd
            if i == k - 1:
                break
            R = (R - F@Dd).tocsr()
            rho_n = 1./(2*sigma - rho)
            Dd = (rho_n*rho*Dd + 2*rho_n/de*(dinv@R)).tocsr()
            rho = rho_n
    S = (J@Z@J.T).tocsr()
    S = .5*(S + S.T)
    S.sort_indices()
    return S.tocsr()


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
    # symmetry is decided on the sparse matrix (a dense allclose of a few
    # thousand rows costs as much as the factorisation itself)
    amax = abs(A).max() if A.nnz else 0.
    sym = A.nnz == 0 or abs(A - A.T).max() <= 1e-10*amax
    dense_inv = None
    import scipy.linalg as sla
    if sym:
        # SPD (the usual case: Schur approximations with an outflow boundary,
        # Galerkin coarse operators): Cholesky inverse, 4x faster than the
        # eigenvalue-based pseudo-inverse
        try:
            cf = sla.cho_factor(.5*(Ad + Ad.T), check_finite=True)
            dg = np.abs(np.diag(cf[0]))
            if dg.min() > 1e-7*dg.max():
                # inverse from the factor (dpotri fills one triangle)
                tri, info = sla.lapack.dpotri(cf[0], lower=cf[1])
                if info == 0:
                    tri = np.tril(tri) if cf[1] else np.triu(tri)
                    dense_inv = tri + tri.T
                    dense_inv[np.diag_indices_from(dense_inv)] *= .5
        except (sla.LinAlgError, ValueError):
            dense_inv = None
    if dense_inv is None and not sym and Ad.size:
        # nonsymmetric Galerkin operator of a convection-diffusion block: LU
        # inverse if LAPACK's condition estimate says it is safe (an SVD-based
        # pseudo-inverse of 2048 rows takes 8x as long)
        lu, piv = sla.lu_factor(Ad, check_finite=False)
        rcond, info = sla.lapack.dgecon(lu, np.abs(Ad).sum(axis=0).max(),
                                        norm='1')
        if info == 0 and rcond > 1e-12:
            dense_inv = sla.lu_solve((lu, piv), np.eye(Ad.shape[0]),
                                     check_finite=False)
    if dense_inv is None:
        # singular (enclosed flow: constant pressure mode)
        dense_inv = np.linalg.pinv(Ad, hermitian=True) if sym \
            else np.linalg.pinv(Ad)
    return levels, dense_inv


def make_saddle_solver(ctx, F1, J, JT=None, F2=None, coef=None, nb=1,
                       restart=40, cheb_steps=3, schur='auto',
                       schur_diag=None, coarse_max=4096, mp_diag=None,
                       mp_scale=None, spectrum=None, hierarchy=None,
                       nsmooth=2, velocity_amg='auto', vgroups=None,
                       vhierarchy=None, Fsym=None, vcoarse_max=2048,
                       mass_diag=None):
    """build a device ``SaddleSolver`` for ``[[F1 + coef_m*F2, JT], [J, 0]]``

    Schur approximation: ``schur='lumped'``: AMG/dense inverse of
    ``J diag^-1 JT`` (``schur_diag`` defaults to the diagonal of the mean
    member matrix) plus, optionally, the Cahouet-Chabard mass term
    ``mp_scale_m * diag(mp_diag)^-1``; ``schur='mass'``: that mass term alone
    (Stokes); ``schur='lsc'``: least-squares commutator
    ``L^-1 (J Du^-1 F Du^-1 JT) L^-1`` with ``L = J Du^-1 JT`` and ``Du`` =
    ``mass_diag`` (diagonal of the velocity mass matrix; default ``diag(F)``)
    -- the choice for the stiffness dominated Stokes / Picard / Newton
    systems of the steady solver; ``'auto'``: ``lsc`` if the velocity block
    needs multigrid, else ``lumped``.  ``velocity_amg``: smoothed-aggregation V-cycle for the
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
        lmin2, lmax2 = (lmin, lmax) if Fmean is Fext else \
            jacobi_spectrum(Fmean)
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
    if schur == 'auto':
        schur = 'lsc' if velocity_amg else 'lumped'
    if schur == 'lsc':
        du = np.abs(Fmean.diagonal()) if mass_diag is None \
            else np.asarray(mass_diag, dtype=float).ravel()
    if schur in ('lumped', 'lsc'):
        if hierarchy is None:
            if schur == 'lsc':
                sd = du
            else:
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
        if schur == 'lsc':
            solver.set_schur_lsc(1./du)
    elif schur != 'mass':
        raise ValueError('schur must be `auto`, `lumped`, `lsc` or `mass`')
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
