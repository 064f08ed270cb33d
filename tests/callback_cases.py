"""Callback set shared by the reference-run fixture of the callback interface
(`tests/golden/make_reference_golden.py::golden_callbacks`) and its GPU test:
a custom nonlinearity ``f_vdp``, a state- and time-dependent forcing
``f_tvdp(t, v)`` and an output-feedback observer as ``dynamic_rhs``
(`tiu:23-58,148-198`, `snu:1129-1140,1224-1234`)."""
import numpy as np


def observer_matrices(NV, M, seed=5):
    rng = np.random.default_rng(seed)
    cmat = rng.standard_normal((3, NV))/np.sqrt(NV)          # y = C v
    bmat = np.asarray(M@rng.standard_normal((NV, 2)))         # B u
    ha = -6.*np.eye(3) + .5*rng.standard_normal((3, 3))
    hb = rng.standard_normal((3, 3))
    hc = .3*rng.standard_normal((2, 3))
    inihx = .1*rng.standard_normal((3, 1))
    return dict(cmat=cmat, bmat=bmat, ha=ha, hb=hb, hc=hc, inihx=inihx,
                drift=lambda t: np.array([[np.cos(30*t)], [0.], [0.]]))


def callback_kwargs(M, inv, convvec_inner, heunab_lti, NV):
    """``convvec_inner(vfull) -> c(v, v)[inv]`` as (NV, 1); ``heunab_lti`` = the
    `get_heunab_lti` of the module under test"""
    om = observer_matrices(NV, M)
    observer = heunab_lti(hb=om['hb'], ha=om['ha'], hc=om['hc'],
                          inihx=om['inihx'], drift=om['drift'])

    def f_vdp(vfull):
        vfull = np.asarray(vfull).reshape(-1, 1)
        return -(convvec_inner(vfull) + .05*(M@vfull[inv]))

    def f_tvdp(t, vc):
        return .2*np.sin(40.*t)*(M@np.asarray(vc).reshape(-1, 1))

    def dynamic_rhs(t, vc=None, memory={}, mode=None):
        u, memory = observer(t, vc=om['cmat']@vc, memory=memory, mode=mode)
        return om['bmat']@u, memory
    return dict(f_vdp=f_vdp, f_tvdp=f_tvdp, dynamic_rhs=dynamic_rhs)
