"""numpy prototype: LSC Schur approximation + SA-AMG velocity V-cycle for the
steady Stokes / Oseen / Newton systems of the DFG 2D-1 problem"""
import sys, time
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spsla
sys.path.insert(0, '.')
from dolfin_navier_scipy_b200 import problem_setups as dnsps, hostsetup
from oracle import convection as oconv, snu as osnu

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
femp, sm, rhsd = dnsps.get_sysmats(
    problem='gen_bccont', nu=1e-3, charvel=.2, scheme='TH', mergerhs=True,
    meshparams=dict(strtomeshfile='mesh/karman2D-rotcyl_lvl%d.xml.gz' % lvl, movingwallcntrl=False,
                    strtophysicalregions='mesh/karman2D-rotcyl_lvl%d_facet_region.xml.gz' % lvl,
                    strtobcsobs='mesh/karman2D-rotcyl-bm_geo_cntrlbc.json'))
A, M, J = sm['A'].tocsr(), sm['M'].tocsr(), sm['J'].tocsr()
NP, NV = J.shape
inv = np.asarray(femp['invinds'])
V = femp['V']
fv, fp = rhsd['fv'], rhsd['fp']
K0 = sps.bmat([[A, J.T], [J, None]], format='csc')
vp = spsla.splu(K0).solve(np.vstack([fv, fp]).ravel())
vfull = osnu.append_bcs_vec(vp[:NV], V.dim(), inv, femp['dbcinds'], femp['dbcvals'])
N1, N2, f3 = oconv.convmats(V, vfull.ravel())
N1c = N1[inv][:, inv]; N2c = N2[inv][:, inv]


def fgmres(K, b, prec, tol=1e-12, maxit=300):
    x0 = np.zeros_like(b)
    r = b - K@x0; beta = np.linalg.norm(r); bn = np.linalg.norm(b)
    Vb = [r/beta]; Z = []; H = np.zeros((maxit+1, maxit))
    for j in range(maxit):
        z = prec(Vb[j]); Z.append(z)
        w = K@z
        for i in range(j+1):
            H[i, j] = Vb[i]@w; w = w - H[i, j]*Vb[i]
        H[j+1, j] = np.linalg.norm(w); Vb.append(w/H[j+1, j])
        e1 = np.zeros(j+2); e1[0] = beta
        y, *_ = np.linalg.lstsq(H[:j+2, :j+1], e1, rcond=None)
        res = np.linalg.norm(H[:j+2, :j+1]@y - e1)
        if res <= tol*bn:
            break
    x = x0 + sum(yi*zi for yi, zi in zip(y, Z))
    return x, j+1, np.linalg.norm(b - K@x)/bn


def cheb(Fm, dinv, r, k, lmin, lmax):
    th = .5*(lmax+lmin); de = .5*(lmax-lmin); sigma = th/de; rho = 1./sigma
    z = np.zeros_like(r); res = r.copy(); d = dinv*res/th
    for i in range(k):
        z = z + d
        if i == k-1:
            break
        res = res - Fm@d
        rho_n = 1./(2*sigma - rho)
        d = rho_n*rho*d + 2*rho_n/de*(dinv*res); rho = rho_n
    return z


def vcycle(levels, dense_inv, l, b, nsmooth=2):
    if l == len(levels):
        return dense_inv@b
    L = levels[l]
    dinv = 1./L['A'].diagonal()
    x = cheb(L['A'], dinv, b, nsmooth, L['lmin'], L['lmax'])
    r = b - L['A']@x
    xc = vcycle(levels, dense_inv, l+1, L['R']@r, nsmooth)
    x = x + L['P']@xc
    r = b - L['A']@x
    return x + cheb(L['A'], dinv, r, nsmooth, L['lmin'], L['lmax'])


for name, F in (('stokes', A), ('picard', (A + N1c).tocsr()), ('newton', (A + N1c + N2c).tocsr())):
    K = sps.bmat([[F, J.T], [J, None]], format='csr')
    b = np.vstack([fv, fp]).ravel()
    du_inv = 1./M.diagonal()
    Lp = hostsetup.lumped_schur(M.diagonal(), J)
    Llu = spsla.splu(Lp.tocsc())
    Flu = spsla.splu(F.tocsc())

    def schur_lsc(rp):
        t = Llu.solve(rp)
        t = du_inv*(J.T@t); t = du_inv*(F@t)
        return Llu.solve(J@t)
    t0 = time.time()
    Fsym = .5*(F + F.T)
    levels, dense_inv = hostsetup.sa_amg_hierarchy(F, coarse_max=2048, groups=(inv//2, inv % 2), Asym=Fsym)
    tset = time.time() - t0
    for vname, vsolve in (('exactF', lambda r: Flu.solve(r)),
                          ('amg2', lambda r: vcycle(levels, dense_inv, 0, r, 2)),
                          ('amg3', lambda r: vcycle(levels, dense_inv, 0, r, 3))):
        def prec(r):
            rv, rp = r[:NV], r[NV:]
            zp = -schur_lsc(rp)
            zv = vsolve(rv - J.T@zp)
            return np.concatenate([zv, zp])
        x, its, rr = fgmres(K, b, prec)
        print(name, vname, 'levels', len(levels), [l['A'].shape[0] for l in levels], 'its', its, 'relres %.1e' % rr, 'amg setup %.1fs' % tset)
