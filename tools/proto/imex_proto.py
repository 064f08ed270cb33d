"""numpy prototype of the planned device algorithm: CNAB loop with FGMRES(CGS-1),
block-triangular preconditioner (Chebyshev-Jacobi on F, exact inverse of lumped Schur),
extrapolated initial guess.  Compares against the LU oracle step by step."""
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spsla, time, sys
from dolfin_navier_scipy_b200 import problem_setups as dnsps
from oracle import snu as osnu, convection as oconv

N = int(sys.argv[1]); Re = float(sys.argv[2]); Nts = int(sys.argv[3]); nsteps = int(sys.argv[4])
NRS = int(sys.argv[10]) if len(sys.argv) > 10 else 1
REORTH = int(sys.argv[9]) if len(sys.argv) > 9 else 1
EPS = float(sys.argv[8]) if len(sys.argv) > 8 else 1e-13
kF = int(sys.argv[5]); tol = float(sys.argv[6]); guess = sys.argv[7] if len(sys.argv) > 7 else 'extrap'
femp, sm, rhsd = dnsps.get_sysmats(problem='cylinderwake', Re=Re, scheme='TH', mergerhs=True, meshparams=dict(refinement_level=N))
M, A, J = sm['M'].tocsr(), sm['A'].tocsr(), sm['J'].tocsr()
NP, NV = J.shape; V = femp['V']; inv = femp['invinds']
trange = np.linspace(0, nsteps/Nts, nsteps+1); dt = trange[1]-trange[0]
sol = dict(A=A, J=J, JT=J.T, M=M, fv=rhsd['fv'], fp=rhsd['fp'], V=V, invinds=inv, dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
ref = osnu.solve_nse(trange=trange, start_ssstokes=True, return_vp_dict=True, **sol)
ts = sorted(ref.keys())

F = (M + .5*dt*A).tocsr(); K = sps.bmat([[F, J.T], [J, None]], format='csr')
D = F.diagonal(); Dinv = 1./D
rng = np.random.default_rng(0)
x = rng.standard_normal(NV)
for _ in range(40):
    x = (F@x)*Dinv; l = np.linalg.norm(x); x /= l
lmax = 1.05*l; lmin = lmax/6.0
S = (J@sps.diags(Dinv)@J.T).toarray(); Sinv = np.linalg.inv(S)
th = .5*(lmax+lmin); de = .5*(lmax-lmin); sigma = th/de
def cheb(r):
    rho = 1./sigma; z = np.zeros_like(r); res = r.copy(); d = Dinv*res/th
    for i in range(kF):
        z += d
        if i == kF-1: break
        res -= F@d; rho_n = 1./(2*sigma-rho); d = rho_n*rho*d + 2*rho_n/de*(Dinv*res); rho = rho_n
    return z
def prec(r):
    zp = -Sinv@r[NV:]; zv = cheb(r[:NV] - J.T@zp); return np.concatenate([zv, zp])
def fgmres(b, x0, tol, maxit=60):
    r = b - K@x0; beta = np.linalg.norm(r); bn = np.linalg.norm(b)
    if beta <= tol*bn: return x0, 0
    Vb = np.zeros((maxit+1, b.size)); Z = np.zeros((maxit, b.size)); H = np.zeros((maxit+1, maxit))
    cs = np.zeros(maxit); sn = np.zeros(maxit); g = np.zeros(maxit+1); g[0] = beta
    Vb[0] = r/beta
    for j in range(maxit):
        Z[j] = prec(Vb[j]); w = K@Z[j]
        h = Vb[:j+1]@w                      # CGS-1
        w = w - h@Vb[:j+1]
        hn = np.linalg.norm(w); Vb[j+1] = w/hn
        H[:j+1, j] = h; H[j+1, j] = hn
        for i in range(j):
            t = cs[i]*H[i,j] + sn[i]*H[i+1,j]; H[i+1,j] = -sn[i]*H[i,j] + cs[i]*H[i+1,j]; H[i,j] = t
        dd = np.hypot(H[j,j], H[j+1,j]); cs[j] = H[j,j]/dd; sn[j] = H[j+1,j]/dd
        H[j,j] = dd; H[j+1,j] = 0; g[j+1] = -sn[j]*g[j]; g[j] = cs[j]*g[j]
        if abs(g[j+1]) <= tol*bn: break
    y = np.linalg.solve(np.triu(H[:j+1,:j+1]), g[:j+1])
    return x0 + y@Z[:j+1], j+1

def appbc(v): return osnu.append_bcs_vec(v, V.dim(), inv, femp['dbcinds'], femp['dbcvals']).ravel()
def nfc(v): return -oconv.convvec(V, appbc(v))[inv]
fv = rhsd['fv'].ravel(); fp = rhsd['fp'].ravel()
# start from the oracle's v1 (Heun step uses different matrices; tested separately)
v0 = ref[ts[0]]['v'][inv].ravel(); v1 = ref[ts[1]]['v'][inv].ravel()
nfc_o = nfc(v0); v_c = v1; xprev = None; xcur = None
r0l = []; fx = []; fb = []; hist = []; itsl = []; errv = []; errp = []
for n in range(2, nsteps+1):
    nfc_c = nfc(v_c)
    rhs = M@v_c - .5*dt*(A@v_c) + .5*dt*(3*nfc_c - nfc_o) + dt*fv
    b = np.concatenate([rhs, fp])
    if guess.startswith('fre'):
        L = int(guess[3:])
        if len(fx) == 0: x0 = np.zeros_like(b) if xcur is None else None
        if len(fx) == 0 and xcur is not None:
            # (re)start: the last NRS full solutions (newest first), orthonormalised
            for (xh_, bh_) in hist[::-1][:NRS]:
                xn_, bn_ = xh_.copy(), bh_.copy()
                for _ in range(2):
                    if len(fb) > 0:
                        c = np.array(fb)@bn_; bn_ = bn_ - c@np.array(fb); xn_ = xn_ - c@np.array(fx)
                nn = np.linalg.norm(bn_); fx.append(xn_/nn); fb.append(bn_/nn)
        if len(fx) > 0:
            al = np.array(fb)@b; x0 = al@np.array(fx)
    elif guess.startswith('gram') and len(hist) >= 1:
        L = int(guess[4:]); Xh = np.array([h[0] for h in hist[-L:]]); Bh = np.array([h[1] for h in hist[-L:]])
        G = Bh@Bh.T; g = Bh@b
        lam, Uu = np.linalg.eigh(G)
        keep = lam > EPS*lam.max()
        c = Uu[:, keep]@((Uu[:, keep].T@g)/lam[keep]); x0 = c@Xh
    elif guess.startswith('fib'):
        L = int(guess[3:])
        if xprev is not None: xb = 2*xcur - xprev
        elif xcur is not None: xb = xcur.copy()
        else: xb = np.concatenate([v_c, np.zeros(NP)])
        if len(fx) == 0: x0 = xb
        else:
            rb = b - K@xb; al = np.array(fb)@rb; x0 = xb + al@np.array(fx)
    elif guess.startswith('fis'):
        L = int(guess[3:])
        if len(fx) == 0: x0 = np.concatenate([v_c, xcur[NV:] if xcur is not None else np.zeros(NP)])
        else:
            al = np.array(fb)@b; x0 = al@np.array(fx)
    elif guess.startswith('proj') and len(hist) >= 2:
        L = int(guess[4:]); Xh = np.array([h[0] for h in hist[-L:]]); Bh = np.array([h[1] for h in hist[-L:]])
        c, *_ = np.linalg.lstsq(Bh.T, b, rcond=None); x0 = c@Xh
    elif guess == 'extrap' and xprev is not None: x0 = 2*xcur - xprev
    elif guess != 'zero' and xcur is not None: x0 = xcur.copy()
    else: x0 = np.zeros_like(b)
    r0n = np.linalg.norm(b - K@x0)/np.linalg.norm(b); r0l.append(r0n)
    xs, its = fgmres(b, x0, tol)
    if guess.startswith('fre'):
        r0 = b - K@x0; beta = np.linalg.norm(r0)
        xn_, bn_ = (xs-x0)/beta, r0/beta
        if len(fb) > 0 and REORTH:
            for _ in range(REORTH):
                c = np.array(fb)@bn_; bn_ = bn_ - c@np.array(fb); xn_ = xn_ - c@np.array(fx)
            nn = np.linalg.norm(bn_); bn_ /= nn; xn_ /= nn
        fx.append(xn_); fb.append(bn_)
        FB = np.array(fb); FX = np.array(fx)
        if False: print('n', n, 'nvec', len(fx), 'orth err', np.abs(FB@FB.T - np.eye(len(fb))).max(), 'consistency', max(np.linalg.norm(K@FX[i]-FB[i]) for i in range(len(fx))), 'beta/|b|', beta/np.linalg.norm(b))
        if len(fx) >= L: fx.clear(); fb.clear()
    if guess.startswith('fis') or guess.startswith('fib'):
        r0 = b - K@x0; beta = np.linalg.norm(r0)
        if len(fx) >= L: fx.pop(0); fb.pop(0)
        fx.append((xs-x0)/beta); fb.append(r0/beta)
    xprev, xcur = xcur, xs
    hist.append((xs, b))
    v_c = xs[:NV]; p = -xs[NV:]/dt; nfc_o = nfc_c
    rv = ref[ts[n]]['v'][inv].ravel(); rp = ref[ts[n]]['p'].ravel()
    itsl.append(its); errv.append(np.linalg.norm(v_c-rv)/np.linalg.norm(rv)); errp.append(np.linalg.norm(p-rp)/np.linalg.norm(rp))
print('r0:', ' '.join('%.0e'%r for r in r0l[:30]))
print(f'N={N} kF={kF} tol={tol} guess={guess}: its mean {np.mean(itsl):.1f} max {max(itsl)} first {itsl[:6]} last {itsl[-4:]}; errv max {max(errv):.1e} errp max {max(errp):.1e}')
