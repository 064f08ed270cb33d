import numpy as np, sys
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, dolfin_to_sparrays as dts, _lib
from oracle import convection as oconv
femp = dnsps.cyl_fems(refinement_level=1)
V = femp['V']
rng = np.random.default_rng(2)
u = rng.standard_normal(V.dim())
dev = _lib.device_for(V)
n1, n2, f3 = dev.convmats(u)
indptr, indices = dev.pattern
import scipy.sparse as sps
NV = V.dim()
N2 = sps.csr_matrix((n2, indices, indptr), shape=(NV, NV))
O1, O2, o3 = oconv.convmats(V, u)
for a in range(2):
    for b in range(2):
        D = N2[a::2, :][:, b::2]; O = O2[a::2, :][:, b::2]
        print('block', a, b, 'dev max', abs(D).max(), 'oracle max', abs(O).max(), 'diff', abs(D-O).max())
print('f3 diff', np.abs(f3 - o3.ravel()).max())
