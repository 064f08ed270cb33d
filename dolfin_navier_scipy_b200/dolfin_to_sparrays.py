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
           'get_convvec', 'condense_sysmatsbybcs', 'condense_velmatsbybcs',
           'append_bcs_vec', 'expand_vp', 'setget_rhs']


def unroll_dlfn_dbcs(diribclist, bcinds=None, bcvals=None):
    """flatten (lists of) Dirichlet indices/values (`dts:27-46`)"""
    if diribclist is not None:
        raise NotImplementedError('dolfin DirichletBC objects need dolfin; '
                                  'pass `dbcinds`/`dbcvals`')
    urbcinds, urbcvals = [], []
    if bcinds is None or len(bcinds) == 0:
        pass
    elif not isinstance(bcinds[0], (list, np.ndarray)):
        urbcinds, urbcvals = bcinds, bcvals
    else:
        for k, cbci in enumerate(bcinds):
            urbcinds.extend(cbci)
            urbcvals.extend(bcvals[k])
    return urbcinds, urbcvals


def append_bcs_vec(vvec, V=None, vdim=None,
                   bcinds=None, bcvals=None,
                   invinds=None, diribcs=None, **kwargs):
    """append boundary values to a vector of inner nodes (`dts:49-64`)"""
    if vdim is None:
        vdim = V.dim()
    vwbcs = np.full((vdim, 1), np.nan)
    cbcinds, cbcvals = unroll_dlfn_dbcs(diribcs, bcinds=bcinds, bcvals=bcvals)
    vwbcs[invinds] = np.asarray(vvec).reshape(-1, 1)
    if len(cbcinds) > 0:
        vwbcs[cbcinds, 0] = cbcvals
    return vwbcs


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
    """resolve the Dirichlet BCs, condense to the inner nodes (`dts:475-573`)"""
    if velbcs is not None:
        raise NotImplementedError('dolfin DirichletBC objects need dolfin')
    bcinds, bcvals = dbcinds, dbcvals
    nv = stms['A'].shape[0]
    if invinds is None:
        invinds = np.setdiff1d(np.arange(nv), bcinds).astype(np.int32)
    auxu = np.zeros((nv, 1))
    if len(bcinds) > 0:
        auxu[bcinds, 0] = bcvals
    fvbc = - stms['A'] * auxu
    fpbc = - stms['J'] * auxu
    fvbc = fvbc[invinds, :]
    if get_rhs_only:
        if mergerhs:
            return {'fv': rhsdict['fv'][invinds, :] + fvbc,
                    'fp': rhsdict['fp'] + fpbc}
        else:
            return {'fv': fvbc, 'fp': fpbc}
    Mc = stms['M'][invinds, :][:, invinds]
    Ac = stms['A'][invinds, :][:, invinds]
    Jc = stms['J'][:, invinds]
    JTc = stms['JT'][invinds, :]
    bcvals = auxu[bcinds]
    stokesmatsc = {'M': Mc, 'A': Ac, 'JT': JTc, 'J': Jc, 'MP': stms['MP']}
    if mergerhs:
        rhsvecsbc = {'fv': rhsdict['fv'][invinds, :] + fvbc,
                     'fp': rhsdict['fp'] + fpbc}
    else:
        rhsvecsbc = {'fv': fvbc, 'fp': fpbc}
    if ret_unrolled:
        return (Mc, Ac, JTc, Jc, stms['MP'], rhsvecsbc['fv'], rhsvecsbc['fp'],
                invinds)
    else:
        return stokesmatsc, rhsvecsbc, invinds, bcinds, bcvals


def condense_velmatsbybcs(A, velbcs=None, return_bcinfo=False,
                          invinds=None, dbcinds=None, dbcvals=None,
                          vwithbcs=None, get_rhs_only=False,
                          columnsonly=False):
    """condense a velocity matrix, rhs contribution of the BCs (`dts:576-642`)"""
    bcinds = None
    if vwithbcs is not None:
        bcsv = np.copy(vwithbcs)
        bcsv[invinds] = 0
    else:
        nv = A.shape[0] if not columnsonly else A.shape[1]
        bcinds, bcvals = unroll_dlfn_dbcs(velbcs, bcinds=dbcinds,
                                          bcvals=dbcvals)
        bcsv = np.zeros((nv, 1))
        if len(bcinds) > 0:
            bcsv[bcinds, 0] = bcvals
    fvbc = - A * bcsv
    if invinds is None:
        ininds = np.setdiff1d(np.arange(nv), bcinds).astype(np.int32)
    else:
        ininds = invinds
    if get_rhs_only:
        return fvbc[ininds, :]
    if columnsonly:
        Ac = A[:, ininds]
    else:
        Ac = A[ininds, :][:, ininds]
        fvbc = fvbc[ininds, :]
    if return_bcinfo:
        return Ac, fvbc, dict(ininds=ininds, bcinds=bcinds)
    else:
        return Ac, fvbc


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
