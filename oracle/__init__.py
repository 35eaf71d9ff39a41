"""CPU oracle for the mind_the_gaps hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``mind_the_gaps_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

Parity status
-------------
The arithmetic of the reference's hot path lives in un-vendored third-party
packages that are NOT importable in the build container (``celerite>=0.4.2``,
``emcee==3.1.4``, ``stingray>=2.0.0rc1``, ``pyfftw``, ``numexpr``, ``astropy``;
pins from ``/root/reference/pyproject.toml:8-18``).  The reference's own tests pin
no log-likelihood, gradient or optimum value (``tests/gpmodelling_test.py`` checks
only ``spread_walkers``), so at the celerite boundary this oracle is
**"parity unpinned"**: it is anchored instead on

* an independent dense O(N^3) Cholesky evaluation of the explicitly built
  covariance matrix (``oracle.gp.loglike_dense``) -- the published celerite kernel
  ``k(tau) = sum_j a_j e^{-c_j tau} + sum_j e^{-c_j tau}(a_j cos d_j tau + b_j sin d_j tau)``
  (Foreman-Mackey et al. 2017, eq. 8/9), against which the O(N J^2) recursion
  restatement is verified;
* the analytic PSD identities the reference tests DO pin
  (``tests/models_test.py:14-102``, 6 decimals);
* the exact exposure-binning index sets of ``tests/simulator_test.py:192-253``;
* the ``stats_test.py`` formulas for the log-normal / uniform factories.

Modules
-------
``terms``      parameter -> celerite-coefficient maps and PSDs
``gp``         dense-Cholesky and semiseparable-recursion log-likelihoods, complex-step gradient
``simulator``  Timmer-Koenig / Emmanoulopoulos synthesis, binning, noise with explicit noise arrays
``sampler``    emcee-3.1.4 stretch move restatement (host numpy)
``cref``       ctypes loader for the C restatement ``celerite_ref.c`` (OpenMP CPU baseline)
"""
