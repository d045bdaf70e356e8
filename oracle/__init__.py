"""CPU oracle for the ann3depth hot path -- TEST INFRASTRUCTURE ONLY.

This package restates the arithmetic of the reference's ``src/models.py`` (MSDN and
DCNF graphs) on the CPU with plain PyTorch tensors (float64 by default).  The reference's
arithmetic lives in the un-vendored ``tensorflow==1.3.0`` wheel (``requirements-cpu.txt:1``),
which cannot be installed in this environment, so every TF op is restated from its
published semantics and each function cites the ``src/models.py`` lines it follows.

PARITY UNPINNED: the reference ships no tests, golden vectors or checkpoints for this path
(SURVEY.md section 4 / 8c).  The oracle is pinned only by the hand-derivable known-answer
tests in ``tests/test_oracle_*.py`` and the self-generated regression vectors under
``tests/golden/`` (generator: ``tests/golden/make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The product path (``ann3depth_b200``) never does.
"""
