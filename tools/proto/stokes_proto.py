"""numpy emulation of the device preconditioner for Stokes/Oseen systems:
velocity SA-AMG V-cycle (Chebyshev smoothing) + scaled pressure-mass Schur."""
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spsla, sys, time
from dolfin_navier_scipy_b200 import problem_setups as dnsps, hostsetup as hs

N = int(sys.argv[1]); Re = float(sys.argv[2]); kF = int(sys.argv[3]); nsm = int(sys.argv[4])
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=Re, scheme='TH', mergerhs=True, meshparams=dict(refinement_level=N))
A, J, MP = sm['A'].tocsr(), sm['J'].tocsr(), sm['MP'].tocsr(); nu = femp['nu']
NP, NV = J.shape
inv = femp['invinds']
groups = (inv//2, inv % 2)
t = time.time()
levels, dense = hs.sa_amg_hierarchy(A, coarse_max=2048, groups=groups)
print('amg setup', time.time()-t, [l['A'].shape[0] for l in levels], dense.shape)
lmin, lmax = hs.jacobi_spectrum(A, ratio=10.)
def cheb(Am, dinv, r, k, lmin, lmax):
    th=.5*(lmax+lmin); de=.5*(lmax-lmin); sg=th/de; rho=1/sg
    z=np.zeros_like(r); res=r.copy(); d=dinv*res/th
    for i in range(k):
        z+=d
        if i==k-1: break
        res-=Am@d; rn=1/(2*sg-rho); d=rn*rho*d+2*rn/de*(dinv*res); rho=rn
    return z
L = [dict(A=A, P=levels[0]['P'], R=levels[0]['R'], lmin=lmin, lmax=lmax, ns=kF)] + [dict(A=l['A'], P=l['P'], R=l['R'], lmin=l['lmin'], lmax=l['lmax'], ns=nsm) for l in levels[1:]] if levels else []
for l in L: l['dinv'] = 1/l['A'].diagonal()
def vcyc(l, b):
    if l == len(L): return dense@b
    lv = L[l]
    x = cheb(lv['A'], lv['dinv'], b, lv['ns'], lv['lmin'], lv['lmax'])
    r = b - lv['A']@x
    xc = vcyc(l+1, lv['R']@r)
    x = x + lv['P']@xc
    r = b - lv['A']@x
    return x + cheb(lv['A'], lv['dinv'], r, lv['ns'], lv['lmin'], lv['lmax'])
mpd = np.asarray(MP.sum(axis=1)).ravel()   # lumped P1 mass
K = sps.bmat([[A, J.T],[J, None]], format='csr')
b = np.concatenate([rhsd['fv'].ravel(), rhsd['fp'].ravel()])
xex = spsla.splu(K.tocsc()).solve(b)
def prec(r):
    zp = -nu*r[NV:]/mpd
    zv = vcyc(0, r[:NV] - J.T@zp)
    return np.concatenate([zv, zp])
def fgmres(b, tol=1e-12, maxit=300, mr=60):
    x = np.zeros_like(b); bn = np.linalg.norm(b); tot=0
    while tot < maxit:
        r = b-K@x; beta=np.linalg.norm(r)
        if beta <= tol*bn: break
        V=[r/beta]; Z=[]; H=np.zeros((mr+1,mr))
        for j in range(mr):
            Z.append(prec(V[j])); w=K@Z[j]
            h=np.array([v@w for v in V]); w=w-sum(hi*v for hi,v in zip(h,V))
            H[:j+1,j]=h; H[j+1,j]=np.linalg.norm(w); V.append(w/H[j+1,j]); tot+=1
            e1=np.zeros(j+2); e1[0]=beta
            y,*_=np.linalg.lstsq(H[:j+2,:j+1],e1,rcond=None)
            if np.linalg.norm(H[:j+2,:j+1]@y-e1)<=tol*bn: break
        x = x+sum(yi*zi for yi,zi in zip(y,Z))
    return x, tot
x, its = fgmres(b)
print(f'N={N} Re={Re} kF={kF} nsm={nsm}: its={its} relres={np.linalg.norm(b-K@x)/np.linalg.norm(b):.1e} errv={np.linalg.norm(x[:NV]-xex[:NV])/np.linalg.norm(xex[:NV]):.1e} errp={np.linalg.norm(x[NV:]-xex[NV:])/np.linalg.norm(xex[NV:]):.1e}')
