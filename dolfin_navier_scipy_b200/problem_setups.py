"""Problem factory -- host side, once per run.

Same call signatures and return dictionaries as the reference's
`dolfin_navier_scipy/problem_setups.py` (`get_sysmats` :34-220, `cyl_fems`
:321-627, `gen_bccont_fems` :773-987), built on the dolfin-free shim in
`fem.py`.  Boundary parts are classified geometrically with the predicates of
`problem_setups.py:445-468` (cylinder wake) or from the geometry JSON
(`gen_bccont`); facets follow dolfin's rule "both vertices and the midpoint
inside".
"""
import json
import logging
import os

import numpy as np

from . import dolfin_to_sparrays as dts
from . import fem

__all__ = ['get_sysmats', 'cyl_fems', 'gen_bccont_fems', 'drivcav_fems']

DOLFIN_EPS = 3.0e-16


def _collect_dbcs(V, masks_and_funs):
    """concatenate `bc.get_boundary_values()` of several Dirichlet parts

    in the given order (`problem_setups.py:595-599`, `:913-917`); returns
    python lists like the reference (duplicates possible, last write wins
    in `dts.condense_sysmatsbybcs`, `dolfin_to_sparrays.py:534-535`).
    """
    mesh = V.mesh()
    xn = V.node_coords()
    dbcinds, dbcvals = [], []
    for mask, fun in masks_and_funs:
        nodes = mesh.facet_nodes(mask)
        vals = np.asarray(fun(xn[nodes])).reshape(-1, 2)
        for comp in range(2):
            dbcinds.extend((2*nodes + comp).tolist())
            dbcvals.extend(vals[:, comp].tolist())
    return dbcinds, dbcvals


def _zero2(x):
    return np.zeros((x.shape[0], 2))


def cyl_fems(refinement_level=2, vdgree=2, pdgree=1, scheme=None,
             inflowvel=1., bccontrol=False, verbose=False, meshdir=None,
             mesh=None):
    """FEM items for the cylinder wake (`problem_setups.py:321-627`)

    Channel 2.2 x 0.41, cylinder at (0.2, 0.2), r=0.05, inflow
    ``4 y (ymax-y)/ymax^2`` (`:576`), two Robin control outlets on the cylinder
    (`:384-411`).  ``mesh`` may be passed to use e.g. a refined mesh.
    """
    if scheme == 'CR':
        raise NotImplementedError('only Taylor-Hood (P2-P1) is supported')
    if vdgree != 2 or pdgree != 1:
        raise NotImplementedError('only P2-P1')
    bmarg = 1.e-3 + DOLFIN_EPS
    xmin, xmax, ymin, ymax = 0.0, 2.2, 0.0, 0.41
    xcenter, ycenter, radius = 0.2, 0.2, 0.05
    centerrad, extensrad = np.pi/3, np.pi/6

    b1xmin = xcenter + radius*np.cos(centerrad + extensrad/2)
    b1ymax = ycenter + radius*np.sin(centerrad + extensrad/2)
    b1xmax = xcenter + radius*np.cos(centerrad - extensrad/2)
    b1ymin = ycenter + radius*np.sin(centerrad - extensrad/2)
    b2xmin, b2xmax = b1xmin, b1xmax
    b2ymin = ycenter - radius*np.sin(centerrad + extensrad/2)
    b2ymax = ycenter - radius*np.sin(centerrad - extensrad/2)
    b1base = np.array([b1xmax - xcenter, b1ymin - ycenter])
    b2base = np.array([b2xmin - xcenter, b2ymin - ycenter])
    centvec = np.array([xcenter, ycenter])
    b1tang = np.array([b1xmax - b1xmin, b1ymin - b1ymax])
    b2tang = np.array([b2xmin - b2xmax, b2ymin - b2ymax])
    rotby90 = np.array([[0, -1.], [1., 0]])
    b1normal = rotby90.dot(b1tang) / np.linalg.norm(b1tang)
    b2normal = rotby90.dot(b2tang) / np.linalg.norm(b2tang)

    def inbb(x, which):
        one = ((x[:, 0] > b1xmin) & (x[:, 0] < b1xmax) &
               (x[:, 1] > b1ymin) & (x[:, 1] < b1ymax))
        two = ((x[:, 0] > b2xmin) & (x[:, 0] < b2xmax) &
               (x[:, 1] > b2ymin) & (x[:, 1] < b2ymax))
        return one if which == 1 else two if which == 2 else (one | two)

    def oncyl(x):
        return np.hypot(x[:, 0] - xcenter, x[:, 1] - ycenter) < radius + bmarg

    if mesh is None:
        if refinement_level > 9:
            raise RuntimeError("No mesh available for refinement level {0}".
                               format(refinement_level))
        mesh = fem.load_mesh('cylinder_%d' % refinement_level, meshdir=meshdir)
    V = fem.VectorP2Space(mesh)
    Q = fem.P1Space(mesh)

    inflow = mesh.mark_facets(lambda x: x[:, 0] < xmin + bmarg)
    walls = mesh.mark_facets(lambda x: (x[:, 1] < ymin + bmarg) |
                             (x[:, 1] > ymax - bmarg))
    if bccontrol:
        cyl = mesh.mark_facets(lambda x: oncyl(x) & ~inbb(x, None))
    else:
        cyl = mesh.mark_facets(oncyl)
    outflow = mesh.mark_facets(lambda x: x[:, 0] > xmax - bmarg)
    cylall = mesh.mark_facets(oncyl)

    def g0(x):
        return np.stack([4*(x[:, 1]*(ymax - x[:, 1]))/(ymax*ymax),
                         np.zeros(x.shape[0])], axis=1)

    dbcinds, dbcvals = _collect_dbcs(V, [(inflow, g0), (walls, _zero2),
                                         (cyl, _zero2)])

    if bccontrol:
        def _shape(base, nvec):
            def fun(x):
                xvec = x - centvec
                carg = xvec.dot(base)/(np.linalg.norm(xvec, axis=1) *
                                       np.linalg.norm(base))
                s = np.arccos(np.clip(carg, -1., 1.))/extensrad
                sf = 1. - 0.5*(1 + np.sin(s*2*np.pi + 0.5*np.pi))
                return sf[:, None]*nvec[None, :]
            return fun
        cmasks = [mesh.mark_facets(lambda x: oncyl(x) & inbb(x, 1)),
                  mesh.mark_facets(lambda x: oncyl(x) & inbb(x, 2))]
        cshapes = [_shape(b1base, b1normal), _shape(b2base, b2normal)]
    else:
        cmasks, cshapes = [None, None], [None, None]

    cylfems = dict(V=V, Q=Q,
                   dbcinds=dbcinds, dbcvals=dbcvals,
                   contrbcsmasks=cmasks,
                   contrbcsshapefuns=cshapes,
                   fv=np.zeros((V.dim(), 1)), fp=np.zeros((Q.dim(), 1)),
                   uspacedep=0,
                   charlen=0.1,
                   mesh=mesh,
                   facetmasks=dict(inflow=inflow, walls=walls, cylinder=cyl,
                                   outflow=outflow, cylinder_all=cylall),
                   ldsbcinds=_vdofs(mesh.facet_nodes(cylall)),
                   odcoo=dict(xmin=0.6, xmax=0.7, ymin=0.15, ymax=0.25),
                   cdcoo=dict(xmin=0.27, xmax=0.32, ymin=0.15, ymax=0.25))
    return cylfems


def _vdofs(nodes):
    return np.stack([2*nodes, 2*nodes + 1], axis=1).ravel().tolist()


def _find_geo(strtobcsobs):
    if os.path.isfile(strtobcsobs):
        return strtobcsobs
    base = os.path.basename(strtobcsobs).replace('_cntrlbc', '')
    cand = os.path.join(fem._MESHDIR, base)
    if os.path.isfile(cand):
        return cand
    raise IOError('geometry file `{0}` not found'.format(strtobcsobs))


def gen_bccont_fems(scheme='TH', bccontrol=True, verbose=False,
                    strtomeshfile='', strtophysicalregions='',
                    inflowvel=1., inflowprofile='parabola',
                    movingwallcntrl=False,
                    strtobcsobs='', meshdir=None, mesh=None):
    """FEM items for a general 2D setup (`problem_setups.py:773-987`)

    The facet-region file (``strtophysicalregions``) is indexed by dolfin's
    private edge numbering and cannot be used without dolfin; the physical
    entities are recovered geometrically from the JSON description instead
    (inflow segment, circles, control inlets, outflow = side opposite to the
    inflow, walls = the rest).  SURVEY.md section 4 lists the facet counts
    that pin this classification.
    """
    if scheme != 'TH':
        raise NotImplementedError('only Taylor-Hood (P2-P1) is supported')
    if mesh is None:
        mesh = fem.load_mesh(strtomeshfile, meshdir=meshdir)
    V = fem.VectorP2Space(mesh)
    Q = fem.P1Space(mesh)
    with open(_find_geo(strtobcsobs)) as f:
        cntbcsdata = json.load(f)

    nbf = mesh.bnd_edge.size
    unassigned = np.ones(nbf, dtype=bool)
    diam = np.ptp(mesh.coords, axis=0).max()
    tol = 1e-8*diam

    def onsegment(xi, xii):
        xi, xii = np.asarray(xi, float), np.asarray(xii, float)
        tvec = (xii - xi)/np.linalg.norm(xii - xi)
        lenb = np.linalg.norm(xii - xi)

        def pred(x):
            d = x - xi
            s = d.dot(tvec)
            dist = np.abs(d[:, 0]*tvec[1] - d[:, 1]*tvec[0])
            return (dist < tol) & (s > -tol) & (s < lenb + tol)
        return pred

    def oncircle(center, radius):
        center = np.asarray(center, float)

        def pred(x):
            return np.linalg.norm(x - center, axis=1) < radius*(1 + 1e-6) + tol
        return pred

    def take(pred):
        mask = mesh.mark_facets(pred) & unassigned
        unassigned[mask] = False
        return mask

    inflowgeodata = cntbcsdata['inflow']
    inflwin = np.array(inflowgeodata['inward normal'], float)
    inflwxi = np.array(inflowgeodata['xone'], float)
    inflwxii = np.array(inflowgeodata['xtwo'], float)
    leninflwb = np.linalg.norm(inflwxi - inflwxii)
    pemasks = {}
    pemasks[inflowgeodata['physical entity']] = take(onsegment(inflwxi,
                                                               inflwxii))
    mvwalls = cntbcsdata.get('moving walls', [])
    for mw in mvwalls:
        if mw['type'] != 'circle':
            raise NotImplementedError()
        pemasks[mw['physical entity']] = \
            take(oncircle(mw['geometry']['center'], mw['geometry']['radius']))
    cntrls = cntbcsdata.get('controlbcs', [])
    for cbc in cntrls:
        if cbc.get('type', 'inlet') == 'rotating circle':
            pemasks[cbc['physical entity']] = \
                take(oncircle(cbc['center'], cbc['radius']))
        else:
            pemasks[cbc['physical entity']] = \
                take(onsegment(cbc['xone'], cbc['xtwo']))
    # outflow: the straight side opposite to the inflow
    sproj = mesh.coords.dot(inflwin)
    smax = sproj.max()
    outflwpe = cntbcsdata['outflow']['physical entity']
    pemasks[outflwpe] = take(lambda x: x.dot(inflwin) > smax - tol)
    wallmask = unassigned.copy()     # everything else is a wall

    if inflowprofile == 'block':
        def inflwprfl(x):
            return np.repeat((inflowvel*inflwin)[None, :], x.shape[0], axis=0)
    elif inflowprofile == 'parabola':
        # `InflowParabola.eval` (`problem_setups.py:1034-1038`)
        def inflwprfl(x):
            curs = np.linalg.norm(x - inflwxi, axis=1)/leninflwb
            return (inflowvel*6*curs*(1 - curs))[:, None]*inflwin[None, :]
    parts = [(pemasks[inflowgeodata['physical entity']], inflwprfl),
             (wallmask, _zero2)]
    if not bccontrol:
        for cbc in cntrls:
            parts.append((pemasks[cbc['physical entity']], _zero2))

    def _rotcirc(center, radius, omega):
        center = np.asarray(center, float)

        def fun(x):
            curn = (x - center)/radius
            anglevel = radius*omega
            return np.stack([-anglevel*curn[:, 1], anglevel*curn[:, 0]],
                            axis=1)
        return fun

    mvwparts = []
    for mw in mvwalls:
        omega = 1. if movingwallcntrl else 0.
        mvwparts.append((pemasks[mw['physical entity']],
                         _rotcirc(mw['geometry']['center'],
                                  mw['geometry']['radius'], omega)))
    if not movingwallcntrl and len(mvwparts) > 0:
        parts.extend(mvwparts)
        mvwparts = []
    dbcinds, dbcvals = _collect_dbcs(V, parts)
    mvwbcinds, mvwbcvals = _collect_dbcs(V, mvwparts)

    bcpes, bcshapefuns, bcmasks = [], [], []
    if bccontrol:
        for cbc in cntrls:
            if cbc.get('type', 'inlet') == 'rotating circle':
                csf = _rotcirc(cbc['center'], cbc['radius'], 1.)
            else:
                cxi, cxii = np.array(cbc['xone']), np.array(cbc['xtwo'])
                lencb = np.linalg.norm(cxi - cxii)
                cbt = 1./lencb*(cxii - cxi)
                cbn = np.array([cbt[1], -cbt[0]])

                def csf(x, cxi=cxi, lencb=lencb, cbn=cbn):
                    curs = np.linalg.norm(x - cxi, axis=1)/lencb
                    return (6*curs*(1 - curs))[:, None]*cbn[None, :]
            bcshapefuns.append(csf)
            bcpes.append(cbc['physical entity'])
            bcmasks.append(pemasks[cbc['physical entity']])

    try:
        ldsurfpe = cntbcsdata['lift drag surface']['physical entity']
        ldsbcinds = _vdofs(mesh.facet_nodes(pemasks[ldsurfpe]))
        liftdragds = pemasks[ldsurfpe]
    except KeyError:
        liftdragds, ldsbcinds = None, None

    gbcfems = dict(V=V, Q=Q,
                   dbcinds=dbcinds, dbcvals=dbcvals,
                   mvwbcinds=mvwbcinds, mvwbcvals=mvwbcvals, mvwtvs=[],
                   outflowds=pemasks[outflwpe],
                   liftdragds=liftdragds, ldsbcinds=ldsbcinds,
                   contrbcspes=bcpes,
                   contrbcsshapefuns=bcshapefuns,
                   cntrbcsds=bcmasks,
                   facetmasks=dict(walls=wallmask, **{'pe%d' % k: v for k, v
                                                      in pemasks.items()}),
                   odcoo=cntbcsdata.get('observation-domain-coordinates'),
                   fv=np.zeros((V.dim(), 1)), fp=np.zeros((Q.dim(), 1)),
                   charlen=cntbcsdata['characteristic length'],
                   mesh=mesh)
    return gbcfems


def drivcav_fems(N=10, vdgree=2, pdgree=1, scheme=None, bccontrol=None):
    """driven cavity on the unit square (`problem_setups.py:223-318`)"""
    mesh = fem.unit_square_mesh(N)
    V = fem.VectorP2Space(mesh)
    Q = fem.P1Space(mesh)
    eps = 1e-12
    lid = mesh.mark_facets(lambda x: x[:, 1] > 1.0 - eps)
    noslip = mesh.mark_facets(lambda x: (x[:, 0] > 1.0 - eps) |
                              (x[:, 1] < eps) | (x[:, 0] < eps))

    def glid(x):
        return np.stack([np.ones(x.shape[0]), np.zeros(x.shape[0])], axis=1)
    # reference order: [noslip, lid] (`problem_setups.py:284`)
    dbcinds, dbcvals = _collect_dbcs(V, [(noslip, _zero2), (lid, glid)])
    return dict(V=V, Q=Q, dbcinds=dbcinds, dbcvals=dbcvals,
                fv=np.zeros((V.dim(), 1)), fp=np.zeros((Q.dim(), 1)),
                uspacedep=0, charlen=1.0, mesh=mesh,
                odcoo=dict(xmin=0.45, xmax=0.55, ymin=0.5, ymax=0.7),
                cdcoo=dict(xmin=0.4, xmax=0.6, ymin=0.2, ymax=0.3))


def get_sysmats(problem='gen_bccont', scheme=None, ppin=None,
                Re=None, nu=None, charvel=1., gradvsymmtrc=True,
                bccontrol=False, mergerhs=False,
                onlymesh=False, meshparams={}):
    """system matrices for Stokes flow -- mirrors `problem_setups.py:34-220`

    Returns ``femp, stokesmatsc, rhsd_vfrc, rhsd_stbc`` (or ``femp,
    stokesmatsc, rhsd`` with ``mergerhs``); matrices are condensed
    ``scipy.sparse.csr_matrix`` objects exactly as in the reference; the
    uncondensed ones are kept under ``stokesmatsc['Afull'|'Mfull'|'Jfull']``
    for the device path and the drag/lift functional.
    ``meshparams['assemble_on_device']`` (extension) assembles the cell
    integrals of M, A, J, MP on the GPU (`dts.get_stokessysmats(device=True)`).
    """
    # ---- which geometry, with which mesh arguments ---------------------------
    geo = dict(meshparams)
    if problem in ('cylinderwake', 'gen_bccont', 'cylinder_rot'):
        geo['inflowvel'] = charvel
    if problem == 'drivencavity':
        builder, geo = drivcav_fems, dict(N=geo['N'])
    elif problem == 'cylinder_rot':
        builder = gen_bccont_fems
        geo['movingwallcntrl'] = True
    elif problem == 'cylinderwake':
        builder = cyl_fems
    elif problem == 'gen_bccont':
        builder = gen_bccont_fems
    else:
        raise KeyError(problem)
    femp = builder(scheme=scheme, bccontrol=bccontrol, **geo)
    if onlymesh:
        return femp

    # ---- Reynolds number <-> viscosity (`dnsps:138-141`) ---------------------
    scale = charvel*femp['charlen']
    nu, Re = (scale/Re, Re) if Re is not None else (nu, scale/nu)

    # ---- uncondensed operators ----------------------------------------------
    full = dts.get_stokessysmats(
        femp['V'], femp['Q'], nu, gradvsymmtrc=gradvsymmtrc,
        outflowds=femp.get('outflowds', None), bccontrol=bccontrol,
        cbshapefuns=femp['contrbcsshapefuns'] if bccontrol else None,
        cbds=femp.get('contrbcsmasks', femp.get('cntrbcsds'))
        if bccontrol else None,
        device=bool(geo.get('assemble_on_device', False)))
    body = dict(fv=np.array(femp['fv'], dtype=float),
                fp=np.array(femp['fp'], dtype=float))

    # ---- pressure pin (`dnsps:171-184`) --------------------------------------
    if problem == 'cylinderwake':
        if ppin is not None:      # outflow boundary: p is fixed by the do-nothing condition
            raise UserWarning('pinning the p will give wrong results')
    elif ppin == -1:
        full['J'], full['JT'] = full['J'][:-1, :], full['JT'][:, :-1]
        body['fp'] = body['fp'][:-1, :]
        logging.info('pressure pinned at last dof `-1`')
    elif ppin is not None:
        raise NotImplementedError('Cannot pin `p` other than at `-1`')
    else:
        logging.debug('pressure is not pinned - `J` may be singular for '
                      'internal flow')

    # ---- Dirichlet condensation ----------------------------------------------
    inner_mats, bc_rhs, inner, _, _ = dts.condense_sysmatsbybcs(
        full, dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
    for key in ('J', 'A', 'M', 'JT'):
        inner_mats[key + 'full'] = full[key]
    if bccontrol:
        arob, arob_load = dts.condense_velmatsbybcs(
            full['amatrob'], dbcinds=femp['dbcinds'], dbcvals=femp['dbcvals'])
        if np.linalg.norm(arob_load) > 1e-15:
            raise UserWarning('diri and control bc must not intersect')
        inner_mats['Arob'], inner_mats['Brob'] = arob, full['bmatrob'][inner, :]
    femp.update(invinds=inner, ppin=ppin, nu=nu, Re=Re)

    body_inner = dict(fv=body['fv'][inner, ], fp=body['fp'])
    if mergerhs:
        return femp, inner_mats, dict(fv=body_inner['fv'] + bc_rhs['fv'],
                                      fp=body_inner['fp'] + bc_rhs['fp'])
    return femp, inner_mats, body_inner, bc_rhs
