"""FEM <-> SciPy bridge with the reference's names and argument meaning.

Mirrors `dolfin_navier_scipy/dolfin_to_sparrays.py` (`dts`):
 * host-side, once: `get_stokessysmats` (`dts:167-322`),
   `condense_sysmatsbybcs` (`dts:475-573`), `condense_velmatsbybcs`
   (`dts:576-642`), `append_bcs_vec` (`dts:49-64`), `unroll_dlfn_dbcs`
   (`dts:27-46`) -- numpy/scipy, same outputs (``(N,1)`` arrays,
   ``csr_matrix``);
 * per-step: `get_convvec` (`dts:427-472`) and `get_convmats`
   (`dts:325-376`) run the CUDA convection kernels (K1a/K1b) through the
   C-ABI; there is no CPU fallback -- without the library they raise.
"""
import numpy as np
import scipy.sparse as sps

from . import fem

__all__ = ['unroll_dlfn_dbcs', 'get_stokessysmats', 'get_convmats',
           'ass_convmat_asmatquad',
           'get_convvec', 'condense_sysmatsbybcs', 'condense_velmatsbybcs',
           'append_bcs_vec', 'expand_vp', 'setget_rhs']


class _Dirichlet(object):
    """index maps of one Dirichlet condensation

    ``n`` full unknowns split into prescribed positions (``idx``, value
    ``val``; a position listed twice keeps its LAST value, `dts:534-535`) and
    the free ones (``free``, ascending).  The three things every helper below
    needs follow from that: the lifting ``g`` (boundary values, zero elsewhere),
    restriction to the free rows/columns and the load ``-(A g)[free]``.
    """

    def __init__(self, n, idx=(), val=(), free=None):
        self.n = int(n)
        self.idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        self.val = np.asarray(val, dtype=float).reshape(-1)
        if free is None:
            mask = np.ones(self.n, dtype=bool)
            mask[self.idx] = False
            free = np.flatnonzero(mask).astype(np.int32)
        self.free = free

    @staticmethod
    def flat(groups_idx, groups_val):
        """``[[i..], [j..]], [[a..], [b..]]`` -> one index and one value list"""
        if groups_idx is None or len(groups_idx) == 0:
            return [], []
        if isinstance(groups_idx[0], (list, np.ndarray)):
            return ([i for grp in groups_idx for i in grp],
                    [v for grp in groups_val for v in grp])
        return groups_idx, groups_val

    def lifting(self):
        g = np.zeros((self.n, 1))
        g[self.idx, 0] = self.val          # fancy assignment: last write wins
        return g

    def load(self, mat, rows=True):
        """``-(mat g)``, restricted to the free rows if ``rows``"""
        f = -(mat@self.lifting())
        return f[self.free, :] if rows else f


def unroll_dlfn_dbcs(diribclist, bcinds=None, bcvals=None):
    """flatten (lists of) Dirichlet indices/values (`dts:27-46`)"""
    if diribclist is not None:
        raise NotImplementedError('dolfin DirichletBC objects need dolfin; '
                                  'pass `dbcinds`/`dbcvals`')
    return _Dirichlet.flat(bcinds, bcvals)


def append_bcs_vec(vvec, V=None, vdim=None,
                   bcinds=None, bcvals=None,
                   invinds=None, diribcs=None, **kwargs):
    """append boundary values to a vector of inner nodes (`dts:49-64`)

    Positions that are neither inner nor prescribed come back as NaN, like in
    the reference."""
    idx, val = unroll_dlfn_dbcs(diribcs, bcinds=bcinds, bcvals=bcvals)
    full = np.full((V.dim() if vdim is None else vdim, 1), np.nan)
    full[invinds] = np.asarray(vvec).reshape(-1, 1)
    if len(idx) > 0:
        full[np.asarray(idx), 0] = val
    return full


def expand_vp(vc=None, pc=None, V=None, Q=None, invinds=None,
              dbcinds=[], dbcvals=[], ppin=None, **kwargs):
    """full coefficient vectors (the dolfin-free part of `dts:645-740`)"""
    v = append_bcs_vec(vc, V=V, invinds=invinds, bcinds=dbcinds,
                       bcvals=dbcvals)
    if pc is None:
        return v, None
    pc = np.asarray(pc).reshape(-1, 1)
    if ppin is not None:
        pc = np.vstack([pc, [[0.]]]) if ppin == -1 else pc
    return v, pc


def get_stokessysmats(V, Q, nu=None, bccontrol=False, gradvsymmtrc=True,
                      outflowds=None,
                      cbclist=None, cbds=None, cbshapefuns=None, device=False):
    """M, A, JT, J, MP [, amatrob, bmatrob] -- `dts:167-322`

    ``outflowds`` / ``cbds`` are boolean masks over ``V.mesh().bnd_edge``
    (the shim's stand-in for dolfin `ds` measures).  ``device=True`` (an
    extension) assembles the cell integrals on the GPU (`dnsb_assemble_stokes`:
    the per-cell machinery of the convection matrices); boundary integrals
    stay on the host.
    """
    if nu is None:
        nu = 1
        print('No viscosity provided -- we set `nu=1`')
    if outflowds is None and gradvsymmtrc:
        print('Note: The symmetric gradient is not corrected in the outflow')
    elif not gradvsymmtrc:
        print('we use the nonsymmetric velocity gradient')
    if device:
        from . import _lib
        stokesmats = _lib.device_for(V).assemble_stokes(
            Q, nu=nu, gradvsymmtrc=gradvsymmtrc)
        if outflowds is not None and gradvsymmtrc and np.any(outflowds):
            A = stokesmats['A'] - \
                nu*fem._assemble_outflow_correction(V, outflowds)
            A.sort_indices()
            stokesmats['A'] = A.tocsr()
    else:
        stokesmats = fem.assemble_stokes_operators(V, Q, nu=nu,
                                                   gradvsymmtrc=gradvsymmtrc,
                                                   outflow_mask=outflowds)
    for key in ('M', 'A', 'J', 'JT', 'MP'):
        stokesmats[key].eliminate_zeros()      # `mat_dolfin2sparse`, dts:80
    if bccontrol:
        amatrobl, bmatrobl = [], []
        xn = V.node_coords()
        for ncb, bcfun in enumerate(cbshapefuns):
            amatrob = fem.assemble_boundary_mass(V, cbds[ncb])
            # `brob = inner(v, bcfun)*cds` with `bcfun` interpolated into V
            # (`problem_setups.py:557-558`): edge mass times nodal values
            gk = np.zeros((V.dim(), 1))
            nodes = V.mesh().facet_nodes(cbds[ncb])
            vals = np.asarray(bcfun(xn[nodes]))
            gk[2*nodes, 0] = vals[:, 0]
            gk[2*nodes + 1, 0] = vals[:, 1]
            amatrobl.append(amatrob)
            bmatrobl.append(amatrob.dot(gk))
        amatrob = amatrobl[0]
        for amatadd in amatrobl[1:]:
            amatrob = amatrob + amatadd
        stokesmats.update({'amatrob': amatrob.tocsr(),
                           'bmatrob': np.hstack(bmatrobl)})
    return stokesmats


def setget_rhs(V, Q, fv, fp, t=None):
    """constant zero body force of all shipped setups (`dts:379-405`)"""
    return {'fv': np.array(fv, dtype=float).reshape(-1, 1),
            'fp': np.array(fp, dtype=float).reshape(-1, 1)}


def condense_sysmatsbybcs(stms, velbcs=None, dbcinds=None, dbcvals=None,
                          invinds=None,
                          mergerhs=False, rhsdict=None, ret_unrolled=False,
                          get_rhs_only=False):
    """resolve the Dirichlet BCs, condense to the inner nodes (`dts:475-573`)

    Returns ``(stokesmatsc, rhsvecsbc, invinds, bcinds, bcvals)`` [or the
    unrolled tuple / only the right-hand sides] with the reference's meaning.
    """
    if velbcs is not None:
        raise NotImplementedError('dolfin DirichletBC objects need dolfin')
    dbc = _Dirichlet(stms['A'].shape[0], dbcinds, dbcvals, free=invinds)
    rhs = {'fv': dbc.load(stms['A']), 'fp': dbc.load(stms['J'], rows=False)}
    if mergerhs:
        rhs = {'fv': rhsdict['fv'][dbc.free, :] + rhs['fv'],
               'fp': rhsdict['fp'] + rhs['fp']}
    if get_rhs_only:
        return rhs
    fr = dbc.free
    mats = {'M': stms['M'][fr, :][:, fr], 'A': stms['A'][fr, :][:, fr],
            'JT': stms['JT'][fr, :], 'J': stms['J'][:, fr], 'MP': stms['MP']}
    if ret_unrolled:
        return (mats['M'], mats['A'], mats['JT'], mats['J'], mats['MP'],
                rhs['fv'], rhs['fp'], fr)
    return mats, rhs, fr, dbcinds, dbc.lifting()[dbcinds]


def condense_velmatsbybcs(A, velbcs=None, return_bcinfo=False,
                          invinds=None, dbcinds=None, dbcvals=None,
                          vwithbcs=None, get_rhs_only=False,
                          columnsonly=False):
    """condense a velocity matrix, rhs contribution of the BCs (`dts:576-642`)

    The boundary data come either as index/value lists or as a full vector
    ``vwithbcs`` whose inner part is ignored.  ``columnsonly``: ``A`` acts on
    full velocities but its rows are something else (e.g. ``J``).
    """
    ncols = A.shape[1] if columnsonly else A.shape[0]
    if vwithbcs is not None:
        lift = np.array(vwithbcs, dtype=float).reshape(ncols, -1)
        lift[invinds] = 0.
        dbc = _Dirichlet(ncols, free=invinds)
        load_full = -(A@lift)
        bnd = None
    else:
        bnd, val = unroll_dlfn_dbcs(velbcs, bcinds=dbcinds, bcvals=dbcvals)
        dbc = _Dirichlet(ncols, bnd, val, free=invinds)
        load_full = dbc.load(A, rows=False)
    fr = dbc.free
    if get_rhs_only:
        return load_full[fr, :]
    if columnsonly:
        Ac, load = A[:, fr], load_full
    else:
        Ac, load = A[fr, :][:, fr], load_full[fr, :]
    if return_bcinfo:
        return Ac, load, dict(ininds=fr, bcinds=bnd)
    return Ac, load


def ass_convmat_asmatquad(W=None, invindsw=None):
    """the convection as a quadratic form: ``H`` with ``N(v)v = H (v kron v)``
    on the inner nodes (`dts:86-164`; 2D, as there)

    The reference assembles, per inner scalar basis function ``phi_i``, the two
    matrices ``int phi_j d_x phi_i phi_k`` and ``int phi_j d_y phi_i phi_k``
    with dolfin and shuffles them into place; here the third-order element
    tensor ``T[k, j, i, d] = int phi_k phi_j d_d phi_i`` is computed for all
    cells at once (7-point rule, exact for the degree-5 integrand) and
    scattered into the same layout:

        H[pos(2k + c),  b*NVi + pos(2j + d)] += T[k, j, i, d],   b = pos(2i + c)

    for both components ``c`` (``pos`` = position in ``invindsw``; as in the
    reference ``invindsw`` must list the x and y dof of every inner node next
    to each other).  Shape ``(NVi, NVi**2)``, csc -- a set-up routine for
    reduced-order models on small meshes (864 entries per cell before merging).
    """
    inv = np.asarray(invindsw, dtype=np.int64)
    nvi = inv.size
    if nvi % 2 or not (np.all(inv[0::2] % 2 == 0) and
                       np.all(inv[1::2] == inv[0::2] + 1)):
        raise ValueError('`invindsw` must hold the (x, y) dof pair of every '
                         'inner node (`dts:102-103`)')
    pos = -np.ones(W.dim(), dtype=np.int64)
    pos[inv] = np.arange(nvi)
    mesh = W.mesh()
    cn = mesh.p2_cell_nodes().astype(np.int64)                 # (nc, 6)
    gl, detj = mesh.geometry()                                 # grad lambda (nc, 3, 2)
    phi = fem.p2_basis(fem.TRI_QP)                             # (nq, 6)
    dphi = np.einsum('qar,crd->cqad', fem.p2_dbasis(fem.TRI_QP), gl)
    wq = fem.TRI_QW[None, :]*(.5*np.abs(detj))[:, None]
    # T[c, k, j, i, d] = sum_q w phi_k phi_j d_d phi_i
    T = np.einsum('cq,qk,qj,cqid->ckjid', wq, phi, phi, dphi)
    rows, cols, vals = [], [], []
    k_, j_, i_ = (cn[:, :, None, None], cn[:, None, :, None],
                  cn[:, None, None, :])
    for c in (0, 1):
        prow = pos[2*k_ + c]
        pb = pos[2*i_ + c]
        for d in (0, 1):
            pcol = pos[2*j_ + d]
            keep = (prow >= 0) & (pb >= 0) & (pcol >= 0)
            keep = np.broadcast_to(keep, T.shape[:4])
            rr = np.broadcast_to(prow, T.shape[:4])[keep]
            cc = (np.broadcast_to(pb, T.shape[:4])[keep]*nvi
                  + np.broadcast_to(pcol, T.shape[:4])[keep])
            rows.append(rr)
            cols.append(cc)
            vals.append(T[..., d][keep])
    hmat = sps.coo_matrix((np.concatenate(vals),
                           (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nvi, nvi*nvi)).tocsc()
    hmat.eliminate_zeros()
    return hmat


def _full_velocity(V, u0_vec, invinds, dbcinds, dbcvals):
    u0 = np.asarray(u0_vec, dtype=np.float64).reshape(-1)
    if u0.size == V.dim():
        return u0
    return append_bcs_vec(u0, V=V, invinds=invinds, bcinds=dbcinds,
                          bcvals=dbcvals).reshape(-1)


def get_convvec(u0_dolfun=None, V=None, u0_vec=None, femp=None,
                uone_utwo_same=True, utwo_dolfun=None, utwo_vec=None,
                dbcvals=None, dbcinds=None,
                diribcs=None, invinds=None):
    """convection vector ``int (grad(u1)*u2).phi dx`` -- `dts:427-472`

    Assembled on the device by the coloured per-cell kernel K1a
    (``dnsb_convvec``).  ``u0_vec`` is the full vector (``V.dim()``) or the
    inner-node vector (then ``invinds``/``dbcinds``/``dbcvals`` expand it).
    """
    from . import _lib
    if femp is not None:
        invinds = femp['invinds']
        dbcinds, dbcvals = femp['dbcinds'], femp['dbcvals']
    if u0_vec is None:
        u0_vec = u0_dolfun
    uone = _full_velocity(V, u0_vec, invinds, dbcinds, dbcvals)
    if uone_utwo_same:
        utwo = None
    else:
        utwo = _full_velocity(V, utwo_vec if utwo_vec is not None
                              else utwo_dolfun, invinds, dbcinds, dbcvals)
    cvec = _lib.device_for(V).convvec(uone, utwo)
    if invinds is not None:
        cvec = cvec[invinds]
    return cvec.reshape(-1, 1)


def get_convmats(u0_dolfun=None, u0_vec=None, V=None, invinds=None,
                 dbcvals=None, dbcinds=None, diribcs=None):
    """``N1 ~ (u0.grad)u``, ``N2 ~ (u.grad)u0``, ``fv = (u0.grad)u0``

    as in `dts:325-376`; assembled on the device by K1b into the fixed CSR
    pattern of the P2 vector space (explicit zeros are removed on return like
    `dts:368-371`).
    """
    from . import _lib
    if u0_vec is None:
        u0_vec = u0_dolfun
    u0 = _full_velocity(V, u0_vec, invinds, dbcinds, dbcvals)
    dev = _lib.device_for(V)
    n1d, n2d, fv = dev.convmats(u0)
    indptr, indices = dev.pattern
    NV = V.dim()
    # copies: `eliminate_zeros` compacts the index arrays in place
    N1 = sps.csr_matrix((n1d, indices.copy(), indptr.copy()), shape=(NV, NV))
    N2 = sps.csr_matrix((n2d, indices.copy(), indptr.copy()), shape=(NV, NV))
    N1.eliminate_zeros()
    N2.eliminate_zeros()
    return N1, N2, fv.reshape(-1, 1)
