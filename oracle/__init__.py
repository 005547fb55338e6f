"""CPU oracle for the ADEPT ``vlasov-1d`` time step -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference algorithm
(``/root/reference/adept/_vlasov1d`` and ``adept/driftdiffusion.py``).  It is the
checker that the CUDA path is compared against; it is never the thing measured or
shipped.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``adept_b200/`` imports it.

Pinning status (see DESIGN.md "Oracle"):
  * grids + initial distribution: pinned against the reference's golden
    ``*_array_config.yml`` files (tests/golden/*.npz, 14 significant figures);
  * operators: pinned by the reference's known-answer tests (characteristic shift
    < 1e-12, bit-exact integer-cell cubic shifts, Chang-Cooper delta identities,
    collision conservation tolerances, Landau damping rate vs the analytic root);
  * post-step arrays of the *JAX* implementation: PARITY UNPINNED -- jax cannot be
    imported in this image and the reference stores no post-step array.
"""
