"""Oracle: convection forms of `dolfin_to_sparrays.py:325-376,427-472`.

TEST INFRASTRUCTURE (see `oracle/__init__.py`).

Independent of the product's tabulations on purpose: the P2 basis comes from a
numerically inverted Vandermonde matrix on the reference triangle and the
integrals use a collapsed 4x4 Gauss-Legendre rule (exact to total degree 6 >= 5
= degree of ``(grad(u) u) . phi`` on affine triangles), whereas the kernels use
closed-form barycentric shape functions and the 7-point Radon rule.
"""
import numpy as np
import scipy.sparse as sps

# reference nodes: vertices (0,0),(1,0),(0,1), then midpoints of the edges
# opposite to vertex 0,1,2 -- the local P2 order of SURVEY.md A.1
_NODES = np.array([[0., 0.], [1., 0.], [0., 1.],
                   [.5, .5], [0., .5], [.5, 0.]])


def _monomials(x, y):
    return np.stack([np.ones_like(x), x, y, x*x, x*y, y*y], axis=-1)


def _dmonomials(x, y):
    z, o = np.zeros_like(x), np.ones_like(x)
    dx = np.stack([z, o, z, 2*x, y, z], axis=-1)
    dy = np.stack([z, z, o, z, x, 2*y], axis=-1)
    return dx, dy


_COEF = np.linalg.inv(_monomials(_NODES[:, 0], _NODES[:, 1]))  # (6 mono, 6 fn)


def _collapsed_rule(n=4):
    g, w = np.polynomial.legendre.leggauss(n)
    g, w = .5*(g + 1.), .5*w
    U, Vv = np.meshgrid(g, g, indexing='ij')
    W = np.outer(w, w)*(1. - U)
    return U.ravel(), (Vv*(1. - U)).ravel(), W.ravel()   # x, y, w (sum 1/2)


_QX, _QY, _QW = _collapsed_rule(4)
_PHI = _monomials(_QX, _QY).dot(_COEF)                    # (nq, 6)
_DX, _DY = _dmonomials(_QX, _QY)
_DPHI_REF = np.stack([_DX.dot(_COEF), _DY.dot(_COEF)], axis=2)   # (nq, 6, 2)


def _cell_data(V):
    mesh = V.mesh()
    x = mesh.coords[mesh.cells]
    # jac[c, d, r] = d x_d / d xi_r
    jac = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]], axis=2)
    det = jac[:, 0, 0]*jac[:, 1, 1] - jac[:, 0, 1]*jac[:, 1, 0]
    jinv = np.linalg.inv(jac)                              # [c, r, d]
    # d phi / d x_d = sum_r d phi / d xi_r * d xi_r / d x_d
    dphi = np.einsum('qar,crd->cqad', _DPHI_REF, jinv)
    wq = _QW[None, :]*np.abs(det)[:, None]
    return V.cell_nodes.astype(np.int64), dphi, wq


def convvec(V, u1, u2=None):
    """``c_i = int (grad(u1) u2) . phi_i dx`` -- `dts:463`; full vector"""
    cn, dphi, wq = _cell_data(V)
    u1 = np.asarray(u1, float).reshape(-1, 2)[cn]            # (nc, 6, 2)
    u2 = u1 if u2 is None else np.asarray(u2, float).reshape(-1, 2)[cn]
    gu = np.einsum('cqmd,cma->cqad', dphi, u1)               # d_d u1_a
    uq = np.einsum('qm,cmb->cqb', _PHI, u2)
    adv = np.einsum('cqad,cqd->cqa', gu, uq)
    cl = np.einsum('cq,cqa,qn->cna', wq, adv, _PHI)          # (nc, 6, 2)
    out = np.zeros(V.dim())
    np.add.at(out, (2*cn[:, :, None] + np.arange(2)[None, None, :]).ravel(),
              cl.ravel())
    return out


def convmats(V, u0):
    """N1, N2 (csr, zeros eliminated) and ``f3`` (N,1) -- `dts:358-376`"""
    cn, dphi, wq = _cell_data(V)
    nc = cn.shape[0]
    u = np.asarray(u0, float).reshape(-1, 2)[cn]
    uq = np.einsum('qm,cmb->cqb', _PHI, u)
    gu = np.einsum('cqmd,cma->cqad', dphi, u)
    # N1[(n,a),(m,a)] = int (u0 . grad phi_m) phi_n
    ugp = np.einsum('cqmd,cqd->cqm', dphi, uq)
    n1s = np.einsum('cq,cqm,qn->cnm', wq, ugp, _PHI)
    n1 = np.zeros((nc, 6, 2, 6, 2))
    for a in range(2):
        n1[:, :, a, :, a] = n1s
    # N2[(n,a),(m,b)] = int d_b u0_a phi_m phi_n
    n2 = np.einsum('cq,cqab,qm,qn->cnamb', wq, gu, _PHI, _PHI)
    vd = (2*cn[:, :, None] + np.arange(2)[None, None, :]).reshape(nc, 12)
    rows = np.repeat(vd[:, :, None], 12, axis=2).ravel()
    cols = np.repeat(vd[:, None, :], 12, axis=1).ravel()
    NV = V.dim()
    N1 = sps.coo_matrix((n1.ravel(), (rows, cols)), shape=(NV, NV)).tocsr()
    N2 = sps.coo_matrix((n2.ravel(), (rows, cols)), shape=(NV, NV)).tocsr()
    N1.eliminate_zeros()
    N2.eliminate_zeros()
    return N1, N2, convvec(V, u0).reshape(-1, 1)
