"""CPU oracle for the mclcar Monte Carlo likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mclcar_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do.

PARITY UNPINNED BY THE REFERENCE: shazhe/mclcar ships no tests, golden vectors
or fixtures for this path and R is not available to run it (SURVEY.md section
8c).  The oracle is pinned instead by closed-form known answers (exact CAR
likelihood ratio, analytic torus moments, finite-difference identities) --
see ``tests/test_oracle_*.py``.
"""
