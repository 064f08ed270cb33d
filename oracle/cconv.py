"""Oracle (TEST INFRASTRUCTURE): ctypes wrapper of the compiled cell loop
``oracle/csrc/convvec.c`` -- the CPU baseline's convection assembly, so that the
baseline is not handicapped by numpy (the reference assembles in FFC-generated
C++, `dolfin_to_sparrays.py:462-470`).  Checked against `oracle.convection`."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle_conv.so')
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            subprocess.check_call(['make', '-s', '-C', _HERE])
        _lib = ctypes.CDLL(_SO)
        ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
        _lib.oracle_convvec.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip,
                                        dp, dp, dp]
        _lib.oracle_convvec.restype = None
    return _lib


class CConv(object):
    def __init__(self, V):
        m = V.mesh()
        self.cells = np.ascontiguousarray(V.cell_nodes, dtype=np.int32)
        self.cverts = np.ascontiguousarray(m.cells, dtype=np.int32)
        self.coords = np.ascontiguousarray(m.coords, dtype=np.float64)
        self.nnodes = V.num_nodes
        self.ncell = m.num_cells
        self.lib = _load()

    def convvec(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        out = np.empty(2*self.nnodes)
        ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
        self.lib.oracle_convvec(self.ncell, self.nnodes,
                                self.cells.ctypes.data_as(ip),
                                self.cverts.ctypes.data_as(ip),
                                self.coords.ctypes.data_as(dp),
                                u.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out
