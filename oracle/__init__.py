"""CPU oracle for the bayesianVARs Gibbs coefficient-draw path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU arm), never as a fallback for the CUDA path.

What it is: a NumPy / SciPy-LAPACK restatement of the reference's arithmetic
(`/root/reference/src/sample_coefficients.cpp`, `src/bvar_cpp.cpp`,
`src/irf.cpp`, `src/utilities_cpp.cpp`), each function citing the file:line it
follows.  It calls the same LAPACK routines RcppArmadillo dispatches to
(dsyrk/dgemm, dpotrf, dtrtri, dgesv) in the same order.

PARITY PINNING STATUS
---------------------
* The reference (R + Rcpp + RcppArmadillo) cannot be built or run in this image
  (no R, no Armadillo, no Rmath), and its test-suite holds NO golden vectors or
  known-answer values for this path - only statistical
  tests.  The linear-algebra oracle (``oracle.regression``, ``oracle.irf``,
  ``oracle.predict``) is therefore pinned by (i) the reference's own statistical
  acceptance tests re-expressed in ``tests/`` (flat-prior ~= OLS,
  Geweke joint-distribution test of ``sample_PHI_cholesky``) and (ii) algebraic
  identities (normal equations, collapsed-vs-Kronecker design).  Value-level
  parity against reference *output* is **unpinned**.
* GIG variates (GIGrvg >= 0.7), the univariate SV sweep (stochvol >= 3.0.3) and
  the factor-SV sweep (factorstochvol >= 1.1.0) live in third-party packages
  whose source is not under /root/reference: **parity unpinned**; the oracle
  restates the published algorithms (Hoermann & Leydold 2014; Kastner &
  Fruehwirth-Schnatter 2014; Kastner, Fruehwirth-Schnatter & Lopes 2017) and the
  tests are distributional.
"""
