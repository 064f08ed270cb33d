"""CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement of the reference's scipy path for the `solve_nse` /
`time_int_utils` hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it -- as the checker, never as the thing
measured or shipped.  The package `dolfin_navier_scipy_b200` never imports
this module.

Parity pinning: the reference (`dolfin`, `sadptprj_riclyap_adi`, `krypy`)
cannot be imported in the build container and ships no golden vectors for
this path.  The oracle is pinned instead on
 (i) the identities of the reference's own unit tests
     (`tests/test_units_fenicsci.py:84-85`, `tests/test_units_pfromv.py:45`,
     `tests/test_units_residuals.py:93-124`),
 (ii) exact polynomial integrals (sympy) for the convection forms,
 (iii) the DFG 2D-1 benchmark values printed at
     `tests/steadystate_schaefer-turek_2D-1.py:112-114`,
 (iv) 2nd-order convergence in time (`tests/tdp_convcheck.py:115-138`).
The third-party saddle-point solver `sadptprj_riclyap_adi.lin_alg_utils`
(un-vendored, version unpinned in `requirements.txt:6`) is restated from its
call sites as an exact sparse-LU solve of ``[[A, J.T], [J, 0]]``.
"""
