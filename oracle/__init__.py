"""CPU oracle for the GMPNP hot path -- TEST INFRASTRUCTURE, not product code.

A NumPy/SciPy restatement of the discrete equations the reference scripts hand to
FEniCS (1D/MPNP_CO2ER_EDL.py:381-595, 737-742; 3D/MPNP_CO2ER_pore.py:503-769,
789-799) together with FFC's quadrature-degree conventions and dolfin's
NewtonSolver semantics.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; nothing under ``gmpnp_b200/`` does.

PARITY: the arithmetic of the reference lives in un-vendored third-party packages
(fenics-dolfin/ffc/ufl/fiat 2019.1.0, petsc 3.12.3, mumps 5.2.1, suitesparse 5.6.0 --
environment.yml:21-27,78,86,110) that can be neither imported nor built in this container, and
the reference has no tests.  1D PATH PINNED: the only result values it holds, the five
(field_OHP, eps_rel_OHP) pairs in 1D/Stern_CO2ER.py:66-68, are reproduced by this oracle to 13
digits by the literal 20000-step replay (V = -2.5) and to 9-10 digits at all five voltages by
extrapolation (state of the default non-dry run at t = 0.2 s; tests/golden/stern_pin_results.json,
tests/test_oracle_1d.py).  3D PATH UNPINNED (no reference output exists for it): it shares forms.py
with the 1D path and is checked by identities, Jacobian consistency and an independent second
restatement (oracle/rxn_diff3d.py) only.
"""
