"""Shared synthetic-input generators for the tests: re-exported from oracle/workloads.py (bench.py's reference leg uses
the same generators without importing the tests package)."""
from oracle.workloads import (FETCH_CIRCLE_OBSTACLES, PANDA_1CUBE_OBSTACLES, OBSTACLES, cuboid_tensors, smooth_joint_path,  # noqa: F401
                              random_configs, synthetic_problem)
