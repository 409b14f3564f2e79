"""CPU oracle for the cppflow path-refinement hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch, fp32 or fp64) restatement of the reference's algorithm for
the hot path named in BASELINE.json: FK + geometric Jacobian, capsule distances, the
Levenberg-Marquardt residual/Jacobian assembly and solve, and the `dp_search` bottleneck
dynamic program.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it; the product (`cppflow_b200/`) never does.

Parity status
-------------
* In-tree reference logic (`search.py`, `optimization.py`, `optimization_utils.py`,
  `collision_detection.py`, `evaluation_utils.py`): PINNED.  `tests/golden/make_golden.py`
  imports the *real* reference modules from `/root/reference` (with `jrl`/`klampt`/`ikflow`/
  `matplotlib` stubbed, the `jrl.Robot` stub backed by this oracle's kinematics) and stores
  input/output vectors under `tests/golden/`; `tests/test_oracle_golden.py` checks the
  restatement against them, together with the known-answer vectors of the reference's own
  tests.
* The `jrl` side (robot tables, FK/Jacobian numerics, capsule tables and capsule distance
  algorithm): PARITY UNPINNED.  jrl 0.1.2 @ ef4c2f6 (pyproject.toml:12) is not installable
  offline; its published algorithm is restated from the public URDFs (see `robots.py`),
  anchored on the reference's call sites and the few vectors its tests hold
  (tests/search_test.py:35-42 limits, tests/planners_test.py:282-309 Panda FK point,
  tests/optimization_utils_test.py:377-402 torso -> z).
"""
