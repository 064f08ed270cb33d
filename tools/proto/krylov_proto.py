"""numpy prototype: iteration counts of block-preconditioned FGMRES on the CNAB matrix"""
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spsla, time, sys
from dolfin_navier_scipy_b200 import problem_setups as dnsps

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Re = float(sys.argv[2]) if len(sys.argv) > 2 else 60
Nts = int(sys.argv[3]) if len(sys.argv) > 3 else 512
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=Re, scheme='TH', mergerhs=True, meshparams=dict(refinement_level=N))
M, A, J, MP = sm['M'].tocsr(), sm['A'].tocsr(), sm['J'].tocsr(), sm['MP'].tocsr()
NP, NV = J.shape
dt = 1.0/Nts; theta = .5; nu = femp['nu']
F = (M + theta*dt*A).tocsr()
K = sps.bmat([[F, J.T], [J, None]], format='csr')
lu = spsla.splu(K.tocsc())
rng = np.random.default_rng(0)
# realistic rhs: Stokes solution advanced
b = np.concatenate([M@rng.standard_normal(NV), np.zeros(NP)])
xex = lu.solve(b)

D = F.diagonal()
def lam_max(Fm, Dm, its=50):
    x = rng.standard_normal(Fm.shape[0])
    for _ in range(its):
        x = (Fm@x)/Dm; l = np.linalg.norm(x); x /= l
    return l
lmax = lam_max(F, D)*1.05
# lmin estimate via inverse iteration (dense-free): use eigsh on small
lmin_true = spsla.eigsh(sps.diags(1/np.sqrt(D))@F@sps.diags(1/np.sqrt(D)), k=1, which='SA', return_eigenvectors=False)[0]
print('N', N, 'NV', NV, 'NP', NP, 'dt', dt, 'nu', nu, 'eig(D^-1 F) in', lmin_true, lmax/1.05)
lmin = lmin_true*0.95

def cheb_apply(Fm, Dinv, r, k, lmin, lmax):
    """k-step Chebyshev iteration for Fm z = r, zero initial guess, Jacobi-preconditioned"""
    th = .5*(lmax+lmin); de = .5*(lmax-lmin)
    sigma = th/de; rho = 1./sigma
    z = np.zeros_like(r); res = r.copy()
    d = Dinv*res/th
    for i in range(k):
        z = z + d
        if i == k-1: break
        res = res - Fm@d
        rho_n = 1./(2*sigma - rho)
        d = rho_n*rho*d + 2*rho_n/de*(Dinv*res)
        rho = rho_n
    return z

Dinv = 1./D
def schur_variant(name):
    if name == 'lump':
        S = (J@sps.diags(Dinv)@J.T).tocsc(); slu = spsla.splu(S); return lambda r: slu.solve(r)
    if name == 'exact':
        flu = spsla.splu(F.tocsc()); S = J@flu.solve(J.T.toarray()); Sinv = np.linalg.inv(S); return lambda r: Sinv@r
    if name == 'ccM':
        mlu = spsla.splu(M.tocsc()); S = J@mlu.solve(J.T.toarray()); Sinv = np.linalg.inv(S)
        mplu = spsla.splu(MP.tocsc())
        return lambda r: Sinv@r + theta*dt*nu*mplu.solve(r)
    if name.startswith('poly'):
        k = int(name[4:])
        # H = polynomial approx of F^-1 of degree k-1 via Chebyshev applied to identity columns (sparse): build explicitly
        I = sps.identity(NV, format='csr')
        th = .5*(lmax+lmin); de = .5*(lmax-lmin); sigma = th/de; rho = 1./sigma
        Di = sps.diags(Dinv)
        Z = sps.csr_matrix((NV,NV)); R = I.copy(); Dd = Di@R/th
        for i in range(k):
            Z = Z + Dd
            if i == k-1: break
            R = R - F@Dd
            rho_n = 1./(2*sigma-rho)
            Dd = rho_n*rho*Dd + 2*rho_n/de*(Di@R)
            rho = rho_n
        S = (J@Z@J.T).tocsc(); print('  nnz(S_poly)/row', S.nnz/NP)
        slu = spsla.splu(S); return lambda r: slu.solve(r)

def fgmres(b, x0, prec, tol=1e-11, maxit=200):
    r = b - K@x0; beta = np.linalg.norm(r); bn = np.linalg.norm(b)
    Vb = [r/beta]; Z = []; H = np.zeros((maxit+1, maxit)); 
    for j in range(maxit):
        z = prec(Vb[j]); Z.append(z)
        w = K@z
        for i in range(j+1):
            H[i,j] = Vb[i]@w; w = w - H[i,j]*Vb[i]
        H[j+1,j] = np.linalg.norm(w); Vb.append(w/H[j+1,j])
        e1 = np.zeros(j+2); e1[0] = beta
        y, *_ = np.linalg.lstsq(H[:j+2,:j+1], e1, rcond=None)
        res = np.linalg.norm(H[:j+2,:j+1]@y - e1)
        if res <= tol*bn: break
    x = x0 + sum(yi*zi for yi, zi in zip(y, Z))
    return x, j+1, np.linalg.norm(b-K@x)/bn

for sname in ['lump', 'poly2', 'poly3', 'ccM', 'exact']:
    t = time.time(); Sinv = schur_variant(sname); ts = time.time()-t
    for kF in (2, 3, 5):
        def prec(r, kF=kF):
            rv, rp = r[:NV], r[NV:]
            zp = -Sinv(rp)
            zv = cheb_apply(F, Dinv, rv - J.T@zp, kF, lmin, lmax)
            return np.concatenate([zv, zp])
        x, its, rr = fgmres(b, np.zeros_like(b), prec)
        err_v = np.linalg.norm(x[:NV]-xex[:NV])/np.linalg.norm(xex[:NV]); err_p = np.linalg.norm(x[NV:]-xex[NV:])/np.linalg.norm(xex[NV:])
        print(f'{sname:6s} kF={kF}: its={its:3d} relres={rr:.1e} err_v={err_v:.1e} err_p={err_p:.1e} (setup {ts:.1f}s)')
