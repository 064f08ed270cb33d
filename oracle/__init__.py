"""CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement of the reference's scipy path for the `solve_nse` /
`time_int_utils` hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it -- as the checker, never as the thing
measured or shipped.  The package `dolfin_navier_scipy_b200` never imports
this module.

Parity pinning: PINNED to reference-run outputs.  `tests/refharness.py` loads
the UNMODIFIED reference modules from /root/reference (`time_int_utils.py`,
`stokes_navier_utils.py`, the condensation helpers of
`dolfin_to_sparrays.py`, `data_output_utils.py`) with stub carrier objects for
`dolfin`, and `tests/golden/make_reference_golden.py` stores what they return
(`tests/golden/ref_*.npz`: `tiu.cnab`, `tiu.sbdftwo`, `tiu.
semi_implicit_euler`, `snu.solve_nse` IMEX branch with and without Robin
control, the Picard/Newton + trapezoidal sweeps, `snu.solve_steadystate_nse`
on DFG 2D-1, `snu.get_pfromv`, `snu.get_v_conv_conts`).
`tests/test_reference_pin.py` checks the oracle against those fixtures (max
rel. error measured 0.0, tolerance 1e-13) and, in the build container, against
the live reference on fresh inputs.
Two pieces cannot be executed here and stay restatements inside that run:
 * `dolfin.assemble` of the two convection forms (`dts:325-376,427-472`,
   FFC-generated C++): `oracle/convection.py`, pinned on the identities of
   the reference's unit tests (`tests/test_units_fenicsci.py:84-85`), exact
   polynomial integrals, and the DFG 2D-1 literature values printed at
   `tests/steadystate_schaefer-turek_2D-1.py:112-114`;
 * the un-vendored `sadptprj_riclyap_adi.lin_alg_utils.solve_sadpnt_smw`
   (bare name in `requirements.txt:6`): an exact sparse-LU solve of
   ``[[A, J.T], [J, 0]]`` as fixed by its call sites (`oracle/lau.py`).
"""
