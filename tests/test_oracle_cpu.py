"""CPU tests: the oracle against the reference's own identities / known answers
 and the host-side FEM shim."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from dolfin_navier_scipy_b200 import fem
from dolfin_navier_scipy_b200 import problem_setups as dnsps
from oracle import convection as oconv
from oracle import snu as osnu
from oracle import tiu as otiu
from oracle.lau import solve_sadpnt_smw

from conftest import soldict


def test_mesh_fixture_counts():
    # sizes and cylinder-facet counts of the fixtures
    expect = {1: (806, 1501, 2307, 18), 2: (1289, 2413, 3702, 31),
              4: (3836, 7364, 11200, 94)}
    for lvl, (nv, nc, ne, ncyl) in expect.items():
        f = dnsps.cyl_fems(refinement_level=lvl)
        m = f['mesh']
        assert (m.num_vertices, m.num_cells, m.num_edges) == (nv, nc, ne)
        assert int(f['facetmasks']['cylinder'].sum()) == ncyl
        masks = f['facetmasks']
        tot = sum(int(masks[k].sum()) for k in
                  ('inflow', 'walls', 'cylinder', 'outflow'))
        assert tot == m.bnd_edge.size          # 0 unclassified


def test_rotcyl_facet_classification_matches_marker_histogram():
    # marker histogram of karman2D-rotcyl_lvl1_facet_region: 18/7/49+49/134
    f = dnsps.gen_bccont_fems(
        scheme='TH', bccontrol=False,
        strtomeshfile='mesh/karman2D-rotcyl_lvl1.xml.gz',
        strtobcsobs='mesh/karman2D-rotcyl-bm_geo_cntrlbc.json',
        inflowvel=.2)
    fm = f['facetmasks']
    assert int(fm['pe1'].sum()) == 18 and int(fm['pe2'].sum()) == 7
    assert int(fm['walls'].sum()) == 98 and int(fm['pe5'].sum()) == 134
    d = np.load(fem._MESHDIR + '/karman2D-rotcyl_lvl1.npz')
    assert list(d['facet_hist'][1:]) == [18, 7, 49, 49, 134]


def test_operator_identities_unit_square():
    m = fem.unit_square_mesh(6)
    V, Q = fem.VectorP2Space(m), fem.P1Space(m)
    ops = fem.assemble_stokes_operators(V, Q, nu=1.)
    M, A, J, MP = ops['M'], ops['A'], ops['J'], ops['MP']
    xn = V.node_coords()
    one = np.ones(V.num_nodes)
    assert abs(one@M[::2, :][:, ::2]@one - 1.) < 1e-13
    assert abs(MP.sum() - 1.) < 1e-13
    rot = np.zeros(V.dim())
    rot[0::2], rot[1::2] = -xn[:, 1], xn[:, 0]
    assert np.abs(A@rot).max() < 1e-12          # rigid rotation: eps(u) = 0
    u = np.zeros(V.dim())
    u[0::2], u[1::2] = xn[:, 0]**2, xn[:, 1]
    assert abs((J@u).sum() - 2.) < 1e-13        # int div u = int 2x+1
    assert abs(u@A@u - 14./3) < 1e-12           # int 2 eps:grad = 2(4x^2+1)
    assert abs(JTnorm(ops)) < 1e-14


def JTnorm(ops):
    return abs(ops['JT'] - ops['J'].T).max()


def test_convection_identities_test_units_fenicsci():
    # `tests/test_units_fenicsci.py:69-85`: convvec(u) == N1(u) u == N2(u) u
    m = fem.unit_square_mesh(8)
    V = fem.VectorP2Space(m)
    xn = V.node_coords()
    x, y = xn[:, 0], xn[:, 1]
    u = np.zeros(V.dim())            # the div-free field of the reference test
    u[0::2] = x*x*(1 - x)*(1 - x)*2*y*(1 - y)*(2*y - 1)
    u[1::2] = y*y*(1 - y)*(1 - y)*2*x*(1 - x)*(1 - 2*x)
    N1, N2, f3 = oconv.convmats(V, u)
    cv = oconv.convvec(V, u)
    assert np.linalg.norm(N1@u - cv) < 1e-14
    assert np.linalg.norm(N2@u - cv) < 1e-14
    assert np.linalg.norm(f3.ravel() - cv) < 1e-14


def test_convection_exact_polynomial_integral():
    # sum_i w_i c_i = int (grad(u) u) . w  for P2 fields u, w: exact value by
    # sympy on the unit square (independent of any quadrature rule)
    import sympy as sp
    xs, ys = sp.symbols('x y')
    ue = sp.Matrix([xs**2 + xs*ys - ys, 1 + xs - 2*ys**2 + xs*ys])
    we = sp.Matrix([ys**2 - xs, xs*ys + 1])
    gu = ue.jacobian([xs, ys])
    integrand = ((gu*ue).T*we)[0, 0]
    exact = float(sp.integrate(sp.integrate(integrand, (xs, 0, 1)), (ys, 0, 1)))
    m = fem.unit_square_mesh(3)
    V = fem.VectorP2Space(m)
    fu = sp.lambdify((xs, ys), ue, 'numpy')
    fw = sp.lambdify((xs, ys), we, 'numpy')
    xn = V.node_coords()
    u = np.zeros(V.dim())
    w = np.zeros(V.dim())
    for k, (a, b) in enumerate(xn):
        uu, ww = np.asarray(fu(a, b)).ravel(), np.asarray(fw(a, b)).ravel()
        u[2*k:2*k+2], w[2*k:2*k+2] = uu, ww
    assert abs(w@oconv.convvec(V, u) - exact) < 1e-12*max(1., abs(exact))


def test_pfromv_identity_test_units_pfromv(cyl1):
    # `tests/test_units_pfromv.py:17-45`: get_pfromv(v_ss) == p_ss
    femp, sm, rhsd = cyl1
    sd = soldict(femp, sm, rhsd)
    sd.pop('JT')
    v, p = osnu.solve_steadystate_nse(return_vp=True, JT=sm['JT'], **sd)
    inv = femp['invinds']
    pfv = osnu.get_pfromv(v=v[inv, :], V=femp['V'], M=sm['M'], A=sm['A'],
                          J=sm['J'], fv=rhsd['fv'], invinds=inv,
                          dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
    assert np.allclose(pfv, p, rtol=1e-8, atol=1e-10)


def test_step_residuals_test_units_residuals(cyl1):
    # `tests/test_units_residuals.py:93-124`: scipy residuals of the Heun and
    # AB2 steps vanish (projected onto the divergence-free space via J)
    femp, sm, rhsd = cyl1
    M, A, J = sm['M'], sm['A'], sm['J']
    V, inv = femp['V'], femp['invinds']
    trange = np.linspace(0., 3./256, 4)
    dt = trange[1] - trange[0]
    vpd = osnu.solve_nse(trange=trange, start_ssstokes=True,
                         return_vp_dict=True, **soldict(femp, sm, rhsd))
    ts = sorted(vpd.keys())

    def nfc(vfull):
        return -oconv.convvec(V, vfull.ravel())[inv].reshape(-1, 1)
    v = [vpd[t]['v'][inv] for t in ts]
    p = [vpd[t]['p'] for t in ts]
    n = [nfc(vpd[t]['v']) for t in ts]
    fv = rhsd['fv']
    # AB2 step 1 -> 2 and 2 -> 3 (`test_units_residuals.py:121-124`)
    for k in (2, 3):
        res = 1./dt*M@(v[k] - v[k-1]) + .5*A@(v[k] + v[k-1]) \
            - (1.5*n[k-1] - .5*n[k-2]) - J.T@p[k] - fv
        assert np.linalg.norm(res) < 1e-9*np.linalg.norm(fv)
        assert np.linalg.norm(J@v[k] - rhsd['fp']) < 1e-10


def test_cnab_second_order_in_time(cyl1):
    # `tests/tdp_convcheck.py:115-138`: halving dt quarters the error
    femp, sm, rhsd = cyl1
    sd = soldict(femp, sm, rhsd)
    tE = 0.05
    ref = osnu.solve_nse(t0=0., tE=tE, Nts=128, start_ssstokes=True,
                         return_final_vp=True, **sd)[0]
    errs = []
    for nts in (8, 16, 32):
        v = osnu.solve_nse(t0=0., tE=tE, Nts=nts, start_ssstokes=True,
                           return_final_vp=True, **sd)[0]
        errs.append(np.linalg.norm(v - ref))
    rates = np.log2(np.array(errs[:-1])/np.array(errs[1:]))
    assert np.all(rates > 1.7), rates


def test_sbdf2_and_imex_euler_run(cyl1):
    femp, sm, rhsd = cyl1
    sd = soldict(femp, sm, rhsd)
    v1 = osnu.solve_nse(t0=0., tE=.02, Nts=16, start_ssstokes=True,
                        return_final_vp=True, time_int_scheme='sbdf2', **sd)[0]
    v2 = osnu.solve_nse(t0=0., tE=.02, Nts=16, start_ssstokes=True,
                        return_final_vp=True, time_int_scheme='cnab', **sd)[0]
    assert np.linalg.norm(v1 - v2) < 2e-3*np.linalg.norm(v2)
    inv = femp['invinds']
    vs = otiu.semi_implicit_euler(
        iniv=v2, jmat=sm['J'], mmat=sm['M'], amat=sm['A'],
        rhsv=lambda t, v: rhsd['fv'], trange=np.linspace(0, .01, 5))
    assert len(vs) == 5 and np.all(np.isfinite(vs[-1]))


def test_solve_sadpnt_smw_contract():
    rng = np.random.default_rng(1)
    A = sps.random(30, 30, .2, random_state=2) + 30*sps.identity(30)
    J = sps.random(6, 30, .5, random_state=3)
    b, g = rng.standard_normal((30, 2)), rng.standard_normal((6, 2))
    vp = solve_sadpnt_smw(amat=A.tocsr(), jmat=J.tocsr(), rhsv=b, rhsp=g)
    assert vp.shape == (36, 2)
    assert np.allclose(A@vp[:30] + J.T@vp[30:], b)
    assert np.allclose(J@vp[:30], g)


@pytest.mark.parametrize('lvl', [1])
def test_schaefer_turek_dfg_2d1(lvl):
    # `tests/steadystate_schaefer-turek_2D-1.py:112-114`: Cd, Cl, dP of DFG 2D-1
    femp, sm, rhsd = dnsps.get_sysmats(
        problem='gen_bccont', nu=1e-3, charvel=.2, scheme='TH', mergerhs=True,
        meshparams=dict(strtomeshfile='mesh/karman2D-rotcyl_lvl%d.xml.gz' % lvl,
                        movingwallcntrl=False,
                        strtophysicalregions='mesh/karman2D-rotcyl_lvl%d_'
                        'facet_region.xml.gz' % lvl,
                        strtobcsobs='mesh/karman2D-rotcyl-bm_geo_cntrlbc.json'))
    v, p = osnu.solve_steadystate_nse(return_vp=True,
                                      **soldict(femp, sm, rhsd))
    cd, cl = osnu.drag_lift(sm['Afull'], sm['JTfull'], femp['V'], v, p,
                            femp['ldsbcinds'])
    dp = femp['Q'].eval_at(p, (.15, .2)) - femp['Q'].eval_at(p, (.25, .2))
    # the tested residual is the force of the cylinder on the fluid (sign)
    assert abs(-cd - 5.57953523384) < 2e-3
    assert abs(-cl - 0.010618948146) < 2e-5
    assert abs(dp - 0.11752016697) < 5e-5


# ---------------------------------------------------------------------------
# golden fixtures (tests/golden/make_golden.py): the oracle must reproduce them
# ---------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b))/np.linalg.norm(b)


def test_golden_convection_oracle(cyl1):
    g = np.load(os.path.join(GOLD, 'convection_cyl1.npz'))
    V = cyl1[0]['V']
    rng = np.random.default_rng(int(g['seed']))
    u = rng.standard_normal(V.dim())
    w = rng.standard_normal(V.dim())
    assert _rel(oconv.convvec(V, u), g['c_uu']) < 1e-13
    assert _rel(oconv.convvec(V, u, w), g['c_uw']) < 1e-13
    N1, N2, f3 = oconv.convmats(V, u)
    assert _rel(N1@w, g['n1_times_w']) < 1e-13
    assert _rel(N2@w, g['n2_times_w']) < 1e-13
    assert _rel(np.ravel(f3), g['f3']) < 1e-13


def test_golden_cnab_oracle(cyl1):
    g = np.load(os.path.join(GOLD, 'cnab_cyl1_re60.npz'))
    femp, sm, rhsd = cyl1
    ref = osnu.solve_nse(t0=0., tE=16./512, Nts=16, start_ssstokes=True,
                         return_vp_dict=True, **soldict(femp, sm, rhsd))
    for k, t in enumerate(g['t']):
        assert _rel(ref[float(t)]['v'], g['v'][:, k:k+1]) < 1e-11
        assert _rel(ref[float(t)]['p'], g['p'][:, k:k+1]) < 1e-9


def test_golden_dfg_numbers():
    g = np.load(os.path.join(GOLD, 'dfg2d1_lvl1.npz'))
    lit = g['literature']
    assert abs(float(g['cd']) - lit[0]) < 2e-3
    assert abs(float(g['cl']) - lit[1]) < 2e-5
    assert abs(float(g['dp']) - lit[2]) < 5e-5


def test_paraview_writer_without_dolfin(tmp_path):
    """`data_output_utils.output_paraview` (`dou:14-71`): one .vtu piece per
    call plus the .pvd collection; quadratic triangles for the velocity, the
    values round-trip, `tfilter` selects the times like the reference"""
    import xml.etree.ElementTree as ET
    from dolfin_navier_scipy_b200 import data_output_utils as dou
    from dolfin_navier_scipy_b200 import fem
    mesh = fem.unit_square_mesh(3, 2)
    V, Q = fem.VectorP2Space(mesh), fem.P1Space(mesh)
    rng = np.random.default_rng(0)
    inv = np.arange(4, V.dim())
    bci, bcv = [0, 1, 2, 3], [.5, -1., 2., 0.]
    prfx = str(tmp_path / 'run')
    vfile = pfile = None
    tfilter = [0., 2.]
    written = {}
    for t in (0., 1., 2.):
        vc, pc = rng.standard_normal((inv.size, 1)), rng.standard_normal((Q.dim(), 1))
        vfile, pfile = dou.output_paraview(V=V, Q=Q, fstring=prfx, invinds=inv, dbcinds=bci, dbcvals=bcv,
                                           vc=vc, pc=pc, t=t, tfilter=tfilter, vfile=vfile, pfile=pfile)
        written[t] = (vc, pc)
    assert tfilter == []
    coll = ET.parse(prfx + '_vel.pvd').getroot().find('Collection').findall('DataSet')
    assert [float(d.get('timestep')) for d in coll] == [0., 2.]
    root = ET.parse(str(tmp_path / coll[1].get('file'))).getroot()
    piece = root.find('UnstructuredGrid').find('Piece')
    assert int(piece.get('NumberOfPoints')) == V.dim()//2
    assert int(piece.get('NumberOfCells')) == mesh.num_cells
    arrays = {a.get('Name'): a for a in piece.iter('DataArray')}
    assert set(arrays['types'].text.split()) == {'22'}
    conn = np.array(arrays['connectivity'].text.split(), dtype=int).reshape(-1, 6)
    xy = np.array(piece.find('Points').find('DataArray').text.split(), dtype=float).reshape(-1, 3)
    # VTK order: node 3 of a cell is the midpoint of the edge (0, 1)
    assert np.allclose(xy[conn[:, 3]], .5*(xy[conn[:, 0]] + xy[conn[:, 1]]))
    assert np.allclose(xy[conn[:, 4]], .5*(xy[conn[:, 1]] + xy[conn[:, 2]]))
    v = np.array(arrays['v'].text.split(), dtype=float).reshape(-1, 3)
    full = np.zeros(V.dim())
    full[inv] = written[2.][0][:, 0]
    full[bci] = bcv
    assert np.array_equal(v[:, 0], full[0::2]) and np.array_equal(v[:, 1], full[1::2])
    proot = ET.parse(str(tmp_path / ET.parse(prfx + '_p.pvd').getroot().find('Collection')
                         .findall('DataSet')[1].get('file'))).getroot()
    parr = {a.get('Name'): a for a in proot.iter('DataArray')}
    assert np.array_equal(np.array(parr['p'].text.split(), dtype=float), written[2.][1][:, 0])
    assert set(parr['types'].text.split()) == {'5'}


def test_convection_as_quadratic_form(cyl1):
    """`dts.ass_convmat_asmatquad` (`dts:86-164`): ``H (v kron v) = N(v) v`` on
    the inner nodes for fields that vanish on the Dirichlet boundary"""
    import scipy.sparse as sps
    from dolfin_navier_scipy_b200 import dolfin_to_sparrays as dts
    from dolfin_navier_scipy_b200 import fem
    from oracle import convection as oconv
    mesh = fem.unit_square_mesh(4, 3)
    V = fem.VectorP2Space(mesh)
    xy = np.asarray(V.tabulate_dof_coordinates())
    inner = (xy[:, 0] > 1e-12) & (xy[:, 0] < 1 - 1e-12) & \
        (xy[:, 1] > 1e-12) & (xy[:, 1] < 1 - 1e-12)
    inv = np.flatnonzero(inner).astype(np.int32)
    H = dts.ass_convmat_asmatquad(W=V, invindsw=inv)
    assert H.shape == (inv.size, inv.size**2) and sps.issparse(H)
    rng = np.random.default_rng(8)
    for _ in range(2):
        vi = rng.standard_normal(inv.size)
        vfull = np.zeros(V.dim())
        vfull[inv] = vi
        ref = oconv.convvec(V, vfull)[inv]
        got = H@np.kron(vi, vi)
        assert np.linalg.norm(got - ref) <= 1e-12*np.linalg.norm(ref)
    with pytest.raises(ValueError):
        dts.ass_convmat_asmatquad(W=V, invindsw=inv[1:-1])


def test_pod_from_gram_gives_orthonormal_modes():
    from dolfin_navier_scipy_b200 import ensemble as ens
    rng = np.random.default_rng(4)
    ns, nv, nb = 9, 40, 3
    A = rng.standard_normal((nv, nv))
    M = A@A.T + nv*np.eye(nv)
    low = rng.standard_normal((nv, 4, nb))
    X = np.einsum('nkm,sk->snm', low, rng.standard_normal((ns, 4)))   # rank 4
    G = sum(X[:, :, m]@M@X[:, :, m].T for m in range(nb))
    lam, W, r = ens.pod_from_gram(G, energy=1 - 1e-12)
    assert r == 4 and np.all(np.diff(lam) <= 1e-9*lam[0])
    Phi = ens.pod_modes(X, W)
    gram = sum(Phi[:, :, m].T@M@Phi[:, :, m] for m in range(nb))
    assert np.allclose(gram, np.eye(r), atol=1e-9)
    # the rank-r reconstruction reproduces the snapshots
    coeff = W.T@G                      # (r, ns)
    for m in range(nb):
        assert np.allclose(Phi[:, :, m]@coeff, X[:, :, m].T, atol=1e-8)
