"""CPU oracle for the lqg inverse-optimal-control likelihood path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
thing timed as the CPU baseline.  The product path (``lqg_b200``) never imports
it and fails loudly when the CUDA library is missing.

Parity status: **unpinned by the reference's own tests** -- the reference
(RothkopfLab/lqg, pure Python/JAX) has no numeric golden vectors
 and JAX/numpyro are not installable in this image, so the
reference itself cannot be run.  The restatement is instead pinned by
independent invariants (``tests/test_oracle_invariants.py``): brute-force joint
Gaussian conditioning, the structural Subjective==Bounded equivalence from the
reference's ``tests/lqg_test.py:69-93``, DARE fixed points, finite differences
and an independently hand-derived adjoint.

Modules
-------
``lqg_np``     float64 NumPy restatement (forward): gains, joint system,
               conditional moments, log-likelihood, simulate, model builders.
``lqg_torch``  float64 torch restatement, batched over parameter samples and
               differentiable (gradient oracle + CPU baseline).
``adjoint_np`` hand-derived reverse-mode equations in NumPy (the math the CUDA
               adjoint kernels implement), checked against ``lqg_torch`` autograd.
"""
