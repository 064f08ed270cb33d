"""Host-side Taylor-Hood (P2-P1) finite-element shim -- no dolfin required.

The reference builds mesh, dof maps and the constant operators with FEniCS
(`dolfin_navier_scipy/dolfin_to_sparrays.py:167-322`,
`dolfin_navier_scipy/problem_setups.py:321-627,773-987`).  FEniCS is not
available on the B200 boxes, so this module provides the same *host-side, once*
setup in plain numpy/scipy: dolfin-XML mesh reader, P2 vector / P1 scalar dof
maps, affine geometry, facet marking and the assembled CSR operators.  All
per-step arithmetic happens in the CUDA library; nothing here is on the hot
path.

Conventions:
 * scalar P2 node numbering: vertices ``0..nv-1`` then edge midpoints
   ``nv..nv+ne-1``; vector dof = ``2*node + comp`` (dolfin interleaves x/y);
 * local P2 order: 3 vertices, then edge ``i`` opposite vertex ``i``;
 * P1 dof = vertex index.
"""
import gzip
import os
import re

import numpy as np
import scipy.sparse as sps

__all__ = ['Mesh', 'VectorP2Space', 'P1Space', 'read_dolfin_xml', 'load_mesh',
           'unit_square_mesh', 'refine_uniform', 'TRI_QP', 'TRI_QW',
           'p2_basis', 'p2_dbasis', 'assemble_stokes_operators',
           'assemble_boundary_mass', 'colour_cells']

# --- 7-point, degree-5 Gauss rule on the triangle (barycentric coordinates) --
_a1 = (6. - np.sqrt(15.)) / 21.
_a2 = (6. + np.sqrt(15.)) / 21.
_w1 = (155. - np.sqrt(15.)) / 1200.
_w2 = (155. + np.sqrt(15.)) / 1200.
TRI_QP = np.array([[1./3, 1./3, 1./3],
                   [1 - 2*_a1, _a1, _a1], [_a1, 1 - 2*_a1, _a1],
                   [_a1, _a1, 1 - 2*_a1],
                   [1 - 2*_a2, _a2, _a2], [_a2, 1 - 2*_a2, _a2],
                   [_a2, _a2, 1 - 2*_a2]])
TRI_QW = np.array([9./40, _w1, _w1, _w1, _w2, _w2, _w2])  # sums to 1

# 3-point Gauss-Legendre on [0, 1] (exact to degree 5)
_g = np.sqrt(3./5)
EDGE_QP = np.array([.5 - .5*_g, .5, .5 + .5*_g])
EDGE_QW = np.array([5./18, 8./18, 5./18])


def p2_basis(lam):
    """P2 shape functions at barycentric points ``lam`` (nq, 3) -> (nq, 6)"""
    l0, l1, l2 = lam[:, 0], lam[:, 1], lam[:, 2]
    return np.stack([l0*(2*l0 - 1), l1*(2*l1 - 1), l2*(2*l2 - 1),
                     4*l1*l2, 4*l0*l2, 4*l0*l1], axis=1)


def p2_dbasis(lam):
    """d phi_a / d lambda_i at ``lam`` (nq, 3) -> (nq, 6, 3)"""
    l0, l1, l2 = lam[:, 0], lam[:, 1], lam[:, 2]
    z = np.zeros_like(l0)
    return np.stack([
        np.stack([4*l0 - 1, z, z], axis=1),
        np.stack([z, 4*l1 - 1, z], axis=1),
        np.stack([z, z, 4*l2 - 1], axis=1),
        np.stack([z, 4*l2, 4*l1], axis=1),
        np.stack([4*l2, z, 4*l0], axis=1),
        np.stack([4*l1, 4*l0, z], axis=1)], axis=1)


class Mesh(object):
    """2D triangle mesh with edge numbering and boundary facets"""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        nv = self.coords.shape[0]
        c = self.cells.astype(np.int64)
        # local edge i is opposite local vertex i
        lo = np.stack([np.minimum(c[:, 1], c[:, 2]),
                       np.minimum(c[:, 0], c[:, 2]),
                       np.minimum(c[:, 0], c[:, 1])], axis=1)
        hi = np.stack([np.maximum(c[:, 1], c[:, 2]),
                       np.maximum(c[:, 0], c[:, 2]),
                       np.maximum(c[:, 0], c[:, 1])], axis=1)
        keys = (lo*nv + hi).ravel()
        ukeys, inv, counts = np.unique(keys, return_inverse=True,
                                       return_counts=True)
        self.edges = np.stack([ukeys // nv, ukeys % nv],
                              axis=1).astype(np.int32)
        self.cell_edges = inv.reshape(-1, 3).astype(np.int32)
        # boundary facets = edges with exactly one adjacent cell
        isb = counts[inv] == 1
        flat = np.nonzero(isb)[0]
        self.bnd_cell = (flat // 3).astype(np.int32)
        self.bnd_local = (flat % 3).astype(np.int32)
        self.bnd_edge = inv[flat].astype(np.int32)

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    @property
    def num_edges(self):
        return self.edges.shape[0]

    def geometry(self):
        """per-cell ``grad(lambda_i)`` (nc, 3, 2) and signed detJ (nc,)"""
        x = self.coords[self.cells]               # (nc, 3, 2)
        x0, x1, x2 = x[:, 0], x[:, 1], x[:, 2]
        detj = (x1[:, 0]-x0[:, 0])*(x2[:, 1]-x0[:, 1]) \
            - (x2[:, 0]-x0[:, 0])*(x1[:, 1]-x0[:, 1])
        gl = np.empty((self.num_cells, 3, 2))
        gl[:, 0, 0] = x1[:, 1] - x2[:, 1]
        gl[:, 0, 1] = x2[:, 0] - x1[:, 0]
        gl[:, 1, 0] = x2[:, 1] - x0[:, 1]
        gl[:, 1, 1] = x0[:, 0] - x2[:, 0]
        gl[:, 2, 0] = x0[:, 1] - x1[:, 1]
        gl[:, 2, 1] = x1[:, 0] - x0[:, 0]
        gl /= detj[:, None, None]
        return gl, detj

    def p2_cell_nodes(self):
        """scalar P2 node ids per cell (nc, 6)"""
        return np.hstack([self.cells,
                          self.cell_edges + self.num_vertices]).astype(np.int32)

    def p2_node_coords(self):
        mid = .5*(self.coords[self.edges[:, 0]] + self.coords[self.edges[:, 1]])
        return np.vstack([self.coords, mid])

    def mark_facets(self, pred):
        """boolean mask over the boundary facets (``self.bnd_edge``)

        A facet is marked iff ``pred`` holds at both vertices and at the
        midpoint -- the rule of ``dolfin.SubDomain.mark`` with
        ``check_midpoint=True`` that `DirichletBC` uses for the subdomains of
        `problem_setups.py:445-468`.  ``pred`` maps (n, 2) -> (n,) bool.
        """
        e = self.edges[self.bnd_edge]
        xa, xb = self.coords[e[:, 0]], self.coords[e[:, 1]]
        return pred(xa) & pred(xb) & pred(.5*(xa + xb))

    def facet_nodes(self, mask):
        """scalar P2 nodes (2 vertices + midpoint) of the marked facets"""
        be = self.bnd_edge[mask]
        e = self.edges[be]
        return np.unique(np.concatenate([e[:, 0], e[:, 1],
                                         be + self.num_vertices]))


class _Space(object):
    def __init__(self, mesh):
        self._mesh = mesh

    def mesh(self):
        return self._mesh


class VectorP2Space(_Space):
    """stand-in for ``dolfin.VectorFunctionSpace(mesh, 'CG', 2)``

    (`problem_setups.py:486`); exposes what `solve_nse(V=...)` needs:
    ``dim()``, cell dof map, dof coordinates.
    """

    def __init__(self, mesh):
        super().__init__(mesh)
        self.cell_nodes = mesh.p2_cell_nodes()
        self.num_nodes = mesh.num_vertices + mesh.num_edges
        self._geom = None
        self._colouring = None

    def dim(self):
        return 2*self.num_nodes

    def node_coords(self):
        return self._mesh.p2_node_coords()

    def tabulate_dof_coordinates(self):
        return np.repeat(self.node_coords(), 2, axis=0)

    def geometry(self):
        if self._geom is None:
            self._geom = self._mesh.geometry()
        return self._geom

    def interpolate(self, fun):
        """nodal interpolation of ``fun: (n,2) -> (n,2)`` into a dof vector"""
        vals = np.asarray(fun(self.node_coords()))
        return vals.reshape(-1)

    def colouring(self):
        if self._colouring is None:
            self._colouring = colour_cells(self._mesh)
        return self._colouring


class P1Space(_Space):
    """stand-in for ``dolfin.FunctionSpace(mesh, 'CG', 1)``"""

    def dim(self):
        return self._mesh.num_vertices

    def tabulate_dof_coordinates(self):
        return self._mesh.coords

    def eval_at(self, pvec, point):
        """point evaluation of a P1 function (for the Delta-p check)"""
        m = self._mesh
        gl, _ = m.geometry()
        x0 = m.coords[m.cells[:, 0]]
        d = np.asarray(point)[None, :] - x0
        l1 = np.einsum('cd,cd->c', gl[:, 1], d)
        l2 = np.einsum('cd,cd->c', gl[:, 2], d)
        l0 = 1. - l1 - l2
        lam = np.stack([l0, l1, l2], axis=1)
        k = np.argmax(lam.min(axis=1))
        if lam[k].min() < -1e-10:
            raise ValueError('point outside the mesh')
        return float(np.dot(lam[k], np.asarray(pvec).ravel()[m.cells[k]]))


# ---------------------------------------------------------------- mesh I/O ---
_vre = re.compile(r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)"')
_tre = re.compile(r'<triangle index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)"')


def read_dolfin_xml(path):
    """read a legacy dolfin-XML triangle mesh (``.xml`` or ``.xml.gz``)"""
    opener = gzip.open if path.endswith('.gz') else open
    with opener(path, 'rt') as f:
        txt = f.read()
    vs = _vre.findall(txt)
    ts = _tre.findall(txt)
    if not vs or not ts:
        raise ValueError('no vertices/triangles found in ' + path)
    coords = np.zeros((len(vs), 2))
    for i, x, y in vs:
        coords[int(i)] = (float(x), float(y))
    cells = np.zeros((len(ts), 3), dtype=np.int32)
    for i, a, b, c in ts:
        cells[int(i)] = (int(a), int(b), int(c))
    return Mesh(coords, cells)


_MESHDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mesh')


def load_mesh(name, meshdir=None):
    """load a mesh by base name, e.g. ``cylinder_1``, ``karman2D-rotcyl_lvl1``

    Looks for ``<name>.npz`` (compact fixtures shipped with the package, made
    by ``tools/convert_meshes.py`` from the reference's ``tests/mesh``), then
    for dolfin XML in ``meshdir``, ``$DNSB_MESH_DIR`` and ``./mesh``.
    """
    if os.path.isfile(name):
        if name.endswith('.npz'):
            d = np.load(name)
            return Mesh(d['coords'], d['cells'])
        return read_dolfin_xml(name)
    base = os.path.basename(name)
    for suf in ('.xml.gz', '.xml', '.npz'):
        if base.endswith(suf):
            base = base[:-len(suf)]
    dirs = [d for d in (meshdir, os.environ.get('DNSB_MESH_DIR'), _MESHDIR,
                        'mesh', os.path.dirname(name)) if d]
    for d in dirs:
        p = os.path.join(d, base + '.npz')
        if os.path.isfile(p):
            dd = np.load(p)
            return Mesh(dd['coords'], dd['cells'])
        for suf in ('.xml.gz', '.xml'):
            p = os.path.join(d, base + suf)
            if os.path.isfile(p):
                return read_dolfin_xml(p)
    raise IOError('mesh `{0}` not found in {1}'.format(name, dirs))


def unit_square_mesh(n, m=None):
    """``dolfin.UnitSquareMesh(n, m)`` look-alike ('right' diagonals)"""
    m = n if m is None else m
    xs, ys = np.linspace(0, 1, n+1), np.linspace(0, 1, m+1)
    X, Y = np.meshgrid(xs, ys)
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    cells = []
    for j in range(m):
        for i in range(n):
            v0 = j*(n+1) + i
            v1, v2, v3 = v0 + 1, v0 + n + 1, v0 + n + 2
            cells.append((v0, v1, v3))
            cells.append((v0, v2, v3))
    return Mesh(coords, np.array(cells, dtype=np.int32))


def refine_uniform(mesh, times=1):
    """uniform red refinement (no boundary snapping) -- SURVEY.md 8(d).6"""
    for _ in range(times):
        nv = mesh.num_vertices
        mid = .5*(mesh.coords[mesh.edges[:, 0]] + mesh.coords[mesh.edges[:, 1]])
        coords = np.vstack([mesh.coords, mid])
        c, e = mesh.cells, mesh.cell_edges + nv
        cells = np.vstack([
            np.stack([c[:, 0], e[:, 2], e[:, 1]], axis=1),
            np.stack([c[:, 1], e[:, 0], e[:, 2]], axis=1),
            np.stack([c[:, 2], e[:, 1], e[:, 0]], axis=1),
            np.stack([e[:, 0], e[:, 1], e[:, 2]], axis=1)])
        mesh = Mesh(coords, cells)
    return mesh


# ------------------------------------------------------------- colouring ----
def colour_cells(mesh, balance=True):
    """greedy colouring of the cell conflict graph (cells sharing a vertex)

    Two cells that share a vertex share at least one P2 dof; cells of one
    colour therefore write disjoint dofs / CSR slots and can scatter without
    atomics (deterministic).  Returns ``(ncolours, colour_of_cell)``; with
    ``balance`` cells are moved from over-full to under-full colours where no
    conflict arises (hard parts).
    """
    nc, nv = mesh.num_cells, mesh.num_vertices
    cells = mesh.cells
    # vertex -> cells adjacency
    order = np.argsort(cells.ravel(), kind='stable')
    vcells = (order // 3).astype(np.int64)
    vptr = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(vptr, cells.ravel().astype(np.int64) + 1, 1)
    vptr = np.cumsum(vptr)
    colour = -np.ones(nc, dtype=np.int32)
    # per-vertex bit mask of the colours already used around it
    vmask = np.zeros(nv, dtype=np.int64)
    for c in range(nc):
        used = vmask[cells[c, 0]] | vmask[cells[c, 1]] | vmask[cells[c, 2]]
        k = 0
        while (used >> k) & 1:
            k += 1
        colour[c] = k
        bit = np.int64(1) << k
        vmask[cells[c]] |= bit
    ncol = int(colour.max()) + 1
    if balance:
        target = int(np.ceil(nc / float(ncol)))
        counts = np.bincount(colour, minlength=ncol)
        for c in range(nc):
            k = colour[c]
            if counts[k] <= target:
                continue
            # colours used by the neighbours (excluding c itself)
            used = 0
            for v in cells[c]:
                for cc in vcells[vptr[v]:vptr[v+1]]:
                    if cc != c:
                        used |= 1 << int(colour[cc])
            cand = [j for j in range(ncol)
                    if not (used >> j) & 1 and counts[j] < target]
            if cand:
                j = min(cand, key=lambda jj: counts[jj])
                colour[c] = j
                counts[k] -= 1
                counts[j] += 1
    return ncol, colour


# --------------------------------------------------------------- assembly ---
def _coo_to_csr(rows, cols, vals, shape):
    mat = sps.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())),
                         shape=shape).tocsr()
    mat.sum_duplicates()
    mat.sort_indices()
    return mat


def _vector_block_rc(cn):
    """row/col index arrays (nc,12,12) of the interleaved vector dofs"""
    vd = np.stack([2*cn, 2*cn + 1], axis=2).reshape(cn.shape[0], 12)
    rows = np.repeat(vd[:, :, None], 12, axis=2)
    cols = np.repeat(vd[:, None, :], 12, axis=1)
    return vd, rows, cols


def assemble_stokes_operators(V, Q, nu=1., gradvsymmtrc=True,
                              outflow_mask=None):
    """M, A, J, JT, MP as in `dolfin_to_sparrays.py:243-275`

    ``A = nu*int 2*eps(u):grad(v)`` (`dts:245`), minus
    ``nu*int_out (grad(u).T n).v ds`` on the facets in ``outflow_mask``
    (`dts:246-248`); ``J = int q div(u)``, ``JT = J.T`` (`dts:254-255`).
    Zero entries are *kept* (fixed pattern for the device); the API layer
    applies ``eliminate_zeros`` like `dts:80`.
    """
    mesh = V.mesh()
    nc = mesh.num_cells
    cn = V.cell_nodes.astype(np.int64)
    gl, detj = V.geometry()
    area = .5*np.abs(detj)
    phi = p2_basis(TRI_QP)                       # (nq, 6)
    dphi_l = p2_dbasis(TRI_QP)                   # (nq, 6, 3)
    # physical gradients (nc, nq, 6, 2)
    dphi = np.einsum('qai,cid->cqad', dphi_l, gl)
    wq = TRI_QW[None, :]*area[:, None]           # (nc, nq)

    # scalar local matrices
    mloc = np.einsum('cq,qa,qb->cab', wq, phi, phi)
    kloc = np.einsum('cq,cqad,cqbd->cab', wq, dphi, dphi)        # grad.grad
    gloc = np.einsum('cq,cqad,cqbe->cadbe', wq, dphi, dphi)      # d_d phi_a d_e phi_b

    vd, rows, cols = _vector_block_rc(cn)
    NV = V.dim()
    mvec = np.zeros((nc, 6, 2, 6, 2))
    avec = np.zeros((nc, 6, 2, 6, 2))
    for a in range(2):
        mvec[:, :, a, :, a] = mloc
        avec[:, :, a, :, a] += kloc
    if gradvsymmtrc:
        # + int d_a phi_m d_b phi_n  for test (n,a), trial (m,b)
        avec += np.einsum('cmanb->cnamb', gloc)
    else:
        pass  # `epsilon = grad`: 2*eps -> `dts:240-241` gives 2*grad(u)
    if not gradvsymmtrc:
        avec *= 2.
    avec *= nu
    M = _coo_to_csr(rows, cols, mvec.reshape(nc, 12, 12), (NV, NV))
    A = _coo_to_csr(rows, cols, avec.reshape(nc, 12, 12), (NV, NV))

    if outflow_mask is not None and gradvsymmtrc and np.any(outflow_mask):
        A = A - nu*_assemble_outflow_correction(V, outflow_mask)
        A.sort_indices()

    # J[k, (m,b)] = int lambda_k d_b phi_m
    jloc = np.einsum('cq,qk,cqmb->ckmb', wq, TRI_QP, dphi)       # (nc,3,6,2)
    prow = np.repeat(mesh.cells.astype(np.int64)[:, :, None], 12, axis=2)
    pcol = np.repeat(vd[:, None, :], 3, axis=1)
    J = _coo_to_csr(prow, pcol, jloc.reshape(nc, 3, 12), (Q.dim(), NV))
    JT = J.T.tocsr()
    JT.sort_indices()
    mploc = np.einsum('cq,qk,ql->ckl', wq, TRI_QP, TRI_QP)
    c3 = mesh.cells.astype(np.int64)
    MP = _coo_to_csr(np.repeat(c3[:, :, None], 3, axis=2),
                     np.repeat(c3[:, None, :], 3, axis=1), mploc,
                     (Q.dim(), Q.dim()))
    return dict(M=M, A=A, J=J, JT=JT, MP=MP)


def _edge_lambda(local, s):
    """barycentric coordinates along local edge ``local`` at parameters s"""
    j, k = [(1, 2), (0, 2), (0, 1)][local]
    lam = np.zeros((s.size, 3))
    lam[:, j] = 1. - s
    lam[:, k] = s
    return lam


def _facet_iter(V, mask):
    mesh = V.mesh()
    for loc in range(3):
        sel = mask & (mesh.bnd_local == loc)
        if not np.any(sel):
            continue
        yield loc, mesh.bnd_cell[sel]


def _assemble_outflow_correction(V, mask):
    """C[(n,a),(m,b)] = int_Gamma d_a phi_m n_b phi_n ds  (`dts:246-248`)"""
    mesh = V.mesh()
    gl, detj = V.geometry()
    NV = V.dim()
    mats = []
    for loc, bc in _facet_iter(V, mask):
        j, k = [(1, 2), (0, 2), (0, 1)][loc]
        xa = mesh.coords[mesh.cells[bc, j]]
        xb = mesh.coords[mesh.cells[bc, k]]
        length = np.linalg.norm(xb - xa, axis=1)
        # outward normal = -grad(lambda_loc)/|grad(lambda_loc)|
        nrm = -gl[bc, loc, :]
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        lam = _edge_lambda(loc, EDGE_QP)
        phi = p2_basis(lam)
        dphi = np.einsum('qai,cid->cqad', p2_dbasis(lam), gl[bc])
        wq = EDGE_QW[None, :]*length[:, None]
        # test (n,a), trial (m,b)
        cl = np.einsum('cq,qn,cqma,cb->cnamb', wq, phi, dphi, nrm)
        cn = V.cell_nodes[bc].astype(np.int64)
        _, rows, cols = _vector_block_rc(cn)
        mats.append(_coo_to_csr(rows, cols, cl.reshape(-1, 12, 12), (NV, NV)))
    out = mats[0]
    for mm in mats[1:]:
        out = out + mm
    return out.tocsr()


def assemble_boundary_mass(V, mask):
    """``int_Gamma u.v ds`` over the marked facets (`dts:304`) as CSR"""
    mesh = V.mesh()
    NV = V.dim()
    out = sps.csr_matrix((NV, NV))
    for loc, bc in _facet_iter(V, mask):
        j, k = [(1, 2), (0, 2), (0, 1)][loc]
        xa = mesh.coords[mesh.cells[bc, j]]
        xb = mesh.coords[mesh.cells[bc, k]]
        length = np.linalg.norm(xb - xa, axis=1)
        phi = p2_basis(_edge_lambda(loc, EDGE_QP))
        wq = EDGE_QW[None, :]*length[:, None]
        ml = np.einsum('cq,qa,qb->cab', wq, phi, phi)
        mv = np.zeros((bc.size, 6, 2, 6, 2))
        for a in range(2):
            mv[:, :, a, :, a] = ml
        cn = V.cell_nodes[bc].astype(np.int64)
        _, rows, cols = _vector_block_rc(cn)
        out = out + _coo_to_csr(rows, cols, mv.reshape(-1, 12, 12), (NV, NV))
    out = out.tocsr()
    out.eliminate_zeros()
    out.sort_indices()
    return out
