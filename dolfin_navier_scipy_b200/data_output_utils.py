"""paraview output and array I/O without dolfin -- mirrors the interface of
`dolfin_navier_scipy/data_output_utils.py` (`output_paraview` at `dou:14-71`,
`save_npa/load_npa` at `dou:74-89`), SURVEY.md 8f-2.

The reference streams dolfin functions into `dolfin.File('..._vel.pvd')`
objects (`vfile << v, t`).  Here `PvdFile` plays that role: every `<<` writes
one VTK XML unstructured-grid piece (`<stem>NNNNNN.vtu`, ASCII) and rewrites
the `.pvd` collection.  Velocities live on the P2 nodes (quadratic triangles,
VTK cell type 22), pressures on the vertices (linear triangles, type 5).
Host-only code: nothing here touches the device.
"""
import os

import numpy as np

from .dolfin_to_sparrays import expand_vp

__all__ = ['output_paraview', 'PvdFile', 'save_npa', 'load_npa']

# VTK quadratic triangle: 3 vertices, then the midpoints of the edges (0,1),
# (1,2), (2,0); the P2 space numbers the midpoint OPPOSITE vertex i as 3+i
_VTK_P2_ORDER = [0, 1, 2, 5, 3, 4]


def _fmt(arr):
    return ' '.join(repr(float(x)) for x in np.asarray(arr).ravel())


class PvdFile(object):
    """stand-in for `dolfin.File(name + '.pvd')`: ``pvd << (values, t)``"""

    def __init__(self, path, space, name='f'):
        self.path = path if path.endswith('.pvd') else path + '.pvd'
        self.stem = self.path[:-4]
        self.space, self.name = space, name
        self.entries = []
        mesh = space.mesh()
        if hasattr(space, 'cell_nodes'):        # vector P2 space
            xy = np.asarray(space.node_coords())
            self.conn = np.asarray(space.cell_nodes)[:, _VTK_P2_ORDER]
            self.celltype, self.ncomp = 22, 3
        else:                                   # P1 space: dofs = vertices
            xy = np.asarray(mesh.coords)
            self.conn = np.asarray(mesh.cells)
            self.celltype, self.ncomp = 5, 1
        self.points = np.column_stack([xy, np.zeros(xy.shape[0])])

    def __lshift__(self, item):
        vals, t = item if isinstance(item, tuple) else (item, len(self.entries))
        self.write(vals, t)
        return self

    def write(self, vals, t):
        vals = np.asarray(vals, dtype=float).reshape(-1)
        npts, ncell = self.points.shape[0], self.conn.shape[0]
        if self.ncomp == 3:
            data = np.column_stack([vals[0::2], vals[1::2], np.zeros(npts)])
            attr = 'Vectors="{0}"'.format(self.name)
        else:
            data = vals[:npts]
            attr = 'Scalars="{0}"'.format(self.name)
        nper = self.conn.shape[1]
        piece = '{0}{1:06d}.vtu'.format(self.stem, len(self.entries))
        with open(piece, 'w') as fh:
            fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" '
                     'version="0.1" byte_order="LittleEndian">\n'
                     '<UnstructuredGrid>\n')
            fh.write('<Piece NumberOfPoints="{0}" NumberOfCells="{1}">\n'
                     .format(npts, ncell))
            fh.write('<Points>\n<DataArray type="Float64" '
                     'NumberOfComponents="3" format="ascii">\n')
            fh.write(_fmt(self.points) + '\n</DataArray>\n</Points>\n')
            fh.write('<Cells>\n<DataArray type="Int32" Name="connectivity" '
                     'format="ascii">\n')
            fh.write(' '.join(str(int(i)) for i in self.conn.ravel()))
            fh.write('\n</DataArray>\n<DataArray type="Int32" Name="offsets" '
                     'format="ascii">\n')
            fh.write(' '.join(str(nper*(k + 1)) for k in range(ncell)))
            fh.write('\n</DataArray>\n<DataArray type="UInt8" Name="types" '
                     'format="ascii">\n')
            fh.write(' '.join([str(self.celltype)]*ncell))
            fh.write('\n</DataArray>\n</Cells>\n')
            fh.write('<PointData {0}>\n<DataArray type="Float64" '
                     'Name="{1}" NumberOfComponents="{2}" format="ascii">\n'
                     .format(attr, self.name, self.ncomp))
            fh.write(_fmt(data) + '\n</DataArray>\n</PointData>\n')
            fh.write('</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
        self.entries.append((float(t), os.path.basename(piece)))
        with open(self.path, 'w') as fh:
            fh.write('<?xml version="1.0"?>\n<VTKFile type="Collection" '
                     'version="0.1">\n<Collection>\n')
            for tt, name in self.entries:
                fh.write('<DataSet timestep="{0!r}" part="0" file="{1}" />\n'
                         .format(tt, name))
            fh.write('</Collection>\n</VTKFile>\n')


def output_paraview(V=None, Q=None, VS=None, fstring='nn',
                    invinds=None, diribcs=None,
                    dbcinds=None, dbcvals=None,
                    vp=None, vc=None, pc=None, sc=None,
                    sname='nn',
                    ppin=None, t=None, tfilter=None, writeoutput=True,
                    vfile=None, pfile=None, sfile=None):
    """write the paraview output for a solution ``(v, p)`` given as
    coefficients -- same arguments as `dou.output_paraview` (`dou:14-71`);
    ``vfile`` / ``pfile`` are `PvdFile` objects (created from ``fstring`` when
    None, like the reference creates `dolfin.File`s).  Returns the files."""
    if not writeoutput:
        return vfile, pfile
    if tfilter is not None:
        if tfilter == []:
            return vfile, pfile
        if not t == tfilter[0]:
            return vfile, pfile
        tfilter.pop(0)
    if diribcs is not None:
        raise NotImplementedError('dolfin DirichletBC objects need dolfin')
    if sc is not None:
        raise NotImplementedError('scalar fields on a separate space `VS`')
    if vp is not None:
        nv = np.asarray(invinds).size if invinds is not None else V.dim()
        vp = np.asarray(vp, dtype=float).reshape(-1, 1)
        vc, pc = vp[:nv], vp[nv:]
    if vc is None:
        return vfile, pfile
    if invinds is None or np.asarray(vc).size == V.dim():
        v, p = np.asarray(vc, dtype=float).reshape(-1, 1), pc
    else:
        v, p = expand_vp(vc=vc, pc=pc, V=V, Q=Q, invinds=invinds,
                         dbcinds=[] if dbcinds is None else dbcinds,
                         dbcvals=[] if dbcvals is None else dbcvals, ppin=ppin)
    if vfile is None:
        vfile = PvdFile(fstring + '_vel.pvd', V, name='v')
    vfile << (v, t)
    if p is not None and Q is not None:
        if pfile is None:
            pfile = PvdFile(fstring + '_p.pvd', Q, name='p')
        pfile << (p, t)
    return vfile, pfile


def save_npa(v, fstring='notspecified'):
    """`dou:74-79`"""
    np.save(fstring, v)


def load_npa(fstring):
    """`dou:82-89`: arrays are stored without the `.npy` suffix in the name"""
    if not fstring[-4:] == '.npy':
        return np.load(fstring + '.npy')
    return np.load(fstring)
