"""one refined-mesh SpMV for an ncu capture:  ncu -k regex:k_spmv_tma -c 2 python tools/ncu_spmv.py [r]"""
import sys
import numpy as np
import scipy.sparse as sps
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import _lib, fem, hostsetup

r = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mesh = fem.refine_uniform(fem.load_mesh('cylinder_4'), r)
V = fem.VectorP2Space(mesh)
nvf = V.dim()
ctx = _lib.default_context(0)
indptr, indices = _lib.ConvDevice(V, ctx).pattern
rng = np.random.default_rng(0)
A = sps.csr_matrix((rng.standard_normal(indices.size), indices, indptr), shape=(nvf, nvf))
perm = hostsetup.locality_perm(np.asarray(V.tabulate_dof_coordinates()), comp=np.arange(nvf) % 2)
A = A[perm][:, perm].tocsr()
mat = ctx.csr(A)
x = rng.standard_normal(nvf)
for _ in range(4):
    y = mat.spmm(x)
print(np.linalg.norm(y - A@x)/np.linalg.norm(y))
